// Micro-benchmark: cost of one tcgen05.mma (M=128, K=16, bf16) as a function of N, shared-memory layout
// (no-swizzle K-major with 16-byte row pitch vs SWIZZLE_128B), start-address alignment and accumulator reuse.
// One CTA, one issuing thread, zero-filled operands.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
struct Cfg { int N, swz, shift_rows, nacc, iters, lbo_rows, avary, bvary; };

__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = __reduce_or_sync(0xffffffffu, slot);
  if (threadIdx.x < 32) {  // warp-uniform issue loop, one elected lane per MMA (no ELECT/BROADCAST waterfall)
    const uint32_t sa = smem_u32(smem) + 1024 + c.shift_rows * (c.swz ? 128 : 16), sb = smem_u32(smem) + 96 * 1024;
    uint64_t ad, bd;
    if (c.swz) {
      ad = (uint64_t)((sa & 0x3ffff) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      bd = (uint64_t)((sb & 0x3ffff) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    } else {
      ad = (uint64_t)((sa & 0x3ffff) >> 4) | ((uint64_t)c.lbo_rows << 16) | ((uint64_t)8 << 32) | (1ull << 46);
      bd = (uint64_t)((sb & 0x3ffff) >> 4) | ((uint64_t)c.N << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    }
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
    const long long t0 = clock64();
    const uint32_t accmask = (uint32_t)c.nacc - 1u;  // nacc is a power of two
    for (int i = 0; i < c.iters; ++i) {
      uint32_t pred;
      asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
      // avary / bvary: rotate the A / B start address over 8 different tiles (defeats any operand reuse between MMAs)
      const uint64_t av = (uint64_t)(((uint32_t)i & 7u) * (uint32_t)c.avary), bv = (uint64_t)(((uint32_t)i & 7u) * (uint32_t)c.bvary);
      if (pred) umma(tm + ((uint32_t)i & accmask) * (uint32_t)c.N, ad + av, bd + bv, idesc, 1u);
    }
    if (threadIdx.x == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const Cfg cfgs[] = {
      // N, swz, shift, nacc, iters, lbo_rows, avary (16-B units), bvary
      {48, 0, 0, 1, 2000, 600, 0, 0},    {48, 0, 0, 1, 2000, 600, 128, 0},  {48, 0, 0, 1, 2000, 600, 0, 96},
      {48, 0, 0, 1, 2000, 600, 128, 96}, {48, 0, 0, 4, 2000, 600, 128, 96}, {48, 1, 0, 1, 2000, 0, 64, 0},
      {96, 0, 0, 1, 2000, 600, 0, 0},    {96, 0, 0, 1, 2000, 600, 128, 192}, {96, 0, 0, 2, 2000, 600, 128, 192},
      {192, 0, 0, 1, 2000, 600, 0, 0},   {192, 0, 0, 1, 2000, 600, 128, 96}, {192, 0, 0, 2, 2000, 600, 128, 96},
      {256, 0, 0, 1, 2000, 600, 0, 0},   {256, 0, 0, 1, 2000, 600, 128, 64}, {256, 1, 0, 1, 2000, 0, 64, 64},
      {256, 1, 0, 1, 2000, 0, 0, 0},     {16, 0, 0, 1, 2000, 600, 128, 0},
  };
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 200 * 1024>>>(c, d);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d %s shift=%d nacc=%d lbo_rows=%d avary=%d bvary=%d: issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %.0f) %s\n", c.N,
           c.swz ? "SW128 " : "noswz ", c.shift_rows, c.nacc, c.lbo_rows, c.avary, c.bvary, (double)h[0] / c.iters, (double)h[1] / c.iters,
           c.N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
