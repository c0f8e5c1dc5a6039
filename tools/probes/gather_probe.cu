// Micro-probe: how fast can a warp-per-4-pixels kernel pull 2x2 taps of 128-byte pixels through the
// memory system, with no arithmetic to speak of?  Variants: L1-allocating vs L1-bypassing loads,
// 1 or 2 rows per iteration, K refs.  Build: nvcc -arch=sm_100a -O3 -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ uint4 ld(const uint4* p) {
  uint4 v;
  if (MODE == 0) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// feat: [B][K][S][S][8 uint4]; out: [B][S][S][8 uint4]; dx,dy: integer shift of the tap origin
template <int MODE, int K, int ROWCHUNK>
__global__ void __launch_bounds__(256) probe(const uint4* __restrict__ feat, uint4* __restrict__ out, int S, int shift) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, j = lane & 7;
  const int tiles_x = S / 32, tiles_y = S / ROWCHUNK;
  int bid = blockIdx.x; const int tx = bid % tiles_x; bid /= tiles_x; const int ty = bid % tiles_y; const int b = bid / tiles_y;
  const int x = tx * 32 + warp * 4 + g;
  const size_t HW = (size_t)S * S;
  for (int y = ty * ROWCHUNK; y < (ty + 1) * ROWCHUNK; ++y) {
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      int x0 = min(max(x + shift * (k + 1), 0), S - 2), y0 = min(max(y + shift, 0), S - 2);
      const uint4* p = feat + (((size_t)b * K + k) * HW + (size_t)y0 * S + x0) * 8 + j;
      uint4 a = ld<MODE>(p), c = ld<MODE>(p + 8), d = ld<MODE>(p + (size_t)S * 8), e = ld<MODE>(p + (size_t)S * 8 + 8);
      acc.x ^= a.x ^ c.x ^ d.x ^ e.x; acc.y ^= a.y ^ c.y ^ d.y ^ e.y; acc.z ^= a.z ^ c.z ^ d.z ^ e.z; acc.w ^= a.w ^ c.w ^ d.w ^ e.w;
    }
    __stcs(out + ((size_t)b * HW + (size_t)y * S + x) * 8 + j, acc);
  }
}
template <int MODE, int K, int RC> float run(const uint4* f, uint4* o, int B, int S, int shift) {
  int grid = B * (S / 32) * (S / RC);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) probe<MODE, K, RC><<<grid, 256>>>(f, o, S, shift);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) probe<MODE, K, RC><<<grid, 256>>>(f, o, S, shift);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
  const int B = 120, K = 4, S = 256;
  size_t nf = (size_t)B * K * S * S * 8, no = (size_t)B * S * S * 8;
  uint4 *f, *o; cudaMalloc(&f, nf * 16); cudaMalloc(&o, no * 16); cudaMemset(f, 1, nf * 16);
  double gb = (nf + no) * 16 / 1e9;
  for (int shift = 0; shift <= 3; shift += 3) {
    float a = run<0, K, 32>(f, o, B, S, shift), b = run<1, K, 32>(f, o, B, S, shift), c = run<2, K, 32>(f, o, B, S, shift);
    float a8 = run<0, K, 8>(f, o, B, S, shift), a256 = run<0, K, 256>(f, o, B, S, shift);
    printf("shift %d: L1-alloc %.3f ms %.0f GB/s | no_allocate %.3f ms %.0f GB/s | cg %.3f ms %.0f GB/s | L1 rowchunk8 %.0f GB/s rowchunk256 %.0f GB/s\n",
           shift, a, gb / a * 1e3, b, gb / b * 1e3, c, gb / c * 1e3, gb / a8 * 1e3, gb / a256 * 1e3);
  }
  printf("cuda err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
