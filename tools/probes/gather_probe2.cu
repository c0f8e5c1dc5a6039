// Probe 2: grow the bare gather toward the real kernel to see which ingredient costs bandwidth.
//   V0 bare gather | V1 + flow loaded from memory, tap computed per lane (redundantly, no shuffles)
//   V2 = V1 + bf16 unpack + FFMA2 math + bf16 pack | V3 = V2 with lane-k prepare + shuffles
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pk(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
struct Tap { unsigned off; float nw, ne, sw, se; };
__device__ __forceinline__ Tap mk(float gx, float gy, int S) {
  float ix = fminf((float)(S - 1), fmaxf(((gx + 1.f) * S - 1.f) * 0.5f, 0.f)), iy = fminf((float)(S - 1), fmaxf(((gy + 1.f) * S - 1.f) * 0.5f, 0.f));
  float fx = fminf(floorf(ix), (float)(S - 2)), fy = fminf(floorf(iy), (float)(S - 2));
  float ax = fx + 1.f - ix, bx = ix - fx, ay = fy + 1.f - iy, by = iy - fy;
  Tap t; t.off = (unsigned)((int)fy * S + (int)fx); t.nw = ax * ay; t.ne = bx * ay; t.sw = ax * by; t.se = bx * by; return t;
}
__device__ __forceinline__ float2 ldg2s(const float2* p){float2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;}
__device__ __forceinline__ float ldg1s(const float* p){float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;}
template <int V, int K, int MINB>
__global__ void __launch_bounds__(256, MINB) probe(const uint4* __restrict__ feat, const float2* __restrict__ grid, uint4* __restrict__ out, int S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, j = lane & 7, gl = g * 8;
  const int tiles_x = S / 32, tiles_y = S / 32;
  int bid = blockIdx.x; const int tx = bid % tiles_x; bid /= tiles_x; const int ty = bid % tiles_y; const int b = bid / tiles_y;
  const int x = tx * 32 + warp * 4 + g;
  const unsigned HW = (unsigned)S * S;
  const char* fb = reinterpret_cast<const char*>(feat) + (size_t)b * K * HW * 128 + j * 16;
  const float2* gb = grid + (size_t)b * K * HW;
  for (int y = ty * 32; y < (ty + 1) * 32; ++y) {
    const unsigned pix = (unsigned)y * S + x;
    float2 acc[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    uint4 xacc = make_uint4(0, 0, 0, 0);
    Tap mine;
    bool any = true; float tm = 1.f;
    if (V >= 3) {
      const float2* gp = gb + ((unsigned)(j & 3) * HW + pix);
      float2 gxy = (V >= 5) ? ldg2s(gp) : *gp; mine = mk(gxy.x, gxy.y, S);
      if (V >= 4) {
        const float* lp = reinterpret_cast<const float*>(gb) + ((unsigned)(j & 3) * HW + pix);  // reuse the flow buffer as logits
        float lg = (V >= 5) ? ldg1s(lp) : *lp;
        float m = fmaxf(lg, __shfl_xor_sync(0xffffffffu, lg, 2)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        float e = expf(lg - m); float ss = e + __shfl_xor_sync(0xffffffffu, e, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        float aw = __fdividef(e, ss); mine.nw *= aw; mine.ne *= aw; mine.sw *= aw; mine.se *= aw;
        if (V == 6 || V == 7) any = __ballot_sync(0xffffffffu, aw != 0.f) != 0u;
        if (V == 6 || V == 8) tm = ldg1s(reinterpret_cast<const float*>(gb) + pix);
        if (V == 9) tm = ldg1s(reinterpret_cast<const float*>(gb) + pix + HW);
      }
    }
    if (any)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      Tap t;
      if (V == 0) { t.off = pix < HW - S - 1 ? pix : 0; t.nw = t.ne = t.sw = t.se = 0.25f; }
      else if (V >= 3) {
        t.off = __shfl_sync(0xffffffffu, mine.off, gl + k); t.nw = __shfl_sync(0xffffffffu, mine.nw, gl + k); t.ne = __shfl_sync(0xffffffffu, mine.ne, gl + k);
        t.sw = __shfl_sync(0xffffffffu, mine.sw, gl + k); t.se = __shfl_sync(0xffffffffu, mine.se, gl + k);
      } else { float2 gxy = gb[(unsigned)k * HW + pix]; t = mk(gxy.x, gxy.y, S); }
      const unsigned o0 = t.off + (unsigned)k * HW;
      const uint4* p0 = reinterpret_cast<const uint4*>(fb + (size_t)o0 * 128);
      const uint4* p1 = reinterpret_cast<const uint4*>(fb + (size_t)(o0 + S) * 128);
      uint4 q[4] = {__ldg(p0), __ldg(p0 + 8), __ldg(p1), __ldg(p1 + 8)};
      if (V <= 1) {
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) { xacc.x ^= q[tp].x; xacc.y ^= q[tp].y; xacc.z ^= q[tp].z; xacc.w ^= q[tp].w; }
      } else {
        const float wt[4] = {t.nw, t.ne, t.sw, t.se};
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {
          const float2 w2 = make_float2(wt[tp], wt[tp]);
          const uint32_t wd[4] = {q[tp].x, q[tp].y, q[tp].z, q[tp].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[c] = __ffma2_rn(make_float2(lo(wd[c]), hi(wd[c])), w2, acc[c]);
        }
      }
    }
    if (V >= 2) xacc = make_uint4(pk(acc[0].x * tm, acc[0].y * tm), pk(acc[1].x * tm, acc[1].y * tm), pk(acc[2].x * tm, acc[2].y * tm), pk(acc[3].x * tm, acc[3].y * tm));
    __stcs(out + ((size_t)b * HW + pix) * 8 + j, xacc);
  }
}
__global__ void fill_grid(float2* g, int S, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i >= n) return;
  int xx = i % S, yy = (i / S) % S; int k = (i / ((size_t)S * S)) % 4;
  g[i] = make_float2((2.f * xx + 1 - S) / S + 0.013f * (k + 1), (2.f * yy + 1 - S) / S + 0.009f * (k + 1));
}
template <int V, int K, int MINB> void run(const char* name, const uint4* f, const float2* gr, uint4* o, int B, int S, double gb) {
  int grid = B * (S / 32) * (S / 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) probe<V, K, MINB><<<grid, 256>>>(f, gr, o, S);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) probe<V, K, MINB><<<grid, 256>>>(f, gr, o, S);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  cudaFuncAttributes at; cudaFuncGetAttributes(&at, probe<V, K, MINB>);
  printf("%-34s minb %d regs %3d  %.3f ms  %.0f GB/s  (%s)\n", name, MINB, at.numRegs, ms, gb / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int B = 120, K = 4, S = 256;
  size_t nf = (size_t)B * K * S * S * 8, no = (size_t)B * S * S * 8, ng = (size_t)B * K * S * S;
  uint4 *f, *o; float2* gr; cudaMalloc(&f, nf * 16); cudaMalloc(&o, no * 16); cudaMalloc(&gr, ng * 8); cudaMemset(f, 1, nf * 16);
  fill_grid<<<(unsigned)((ng + 255) / 256), 256>>>(gr, S, ng);
  double gb0 = (nf + no) * 16 / 1e9, gb1 = gb0 + ng * 8 / 1e9;
  run<0, K, 4>("V0 bare gather", f, gr, o, B, S, gb0);
  run<0, K, 6>("V0 bare gather", f, gr, o, B, S, gb0);
  run<1, K, 4>("V1 +flow load, tap per lane", f, gr, o, B, S, gb1);
  run<1, K, 6>("V1 +flow load, tap per lane", f, gr, o, B, S, gb1);
  run<2, K, 4>("V2 +unpack/FFMA2/pack", f, gr, o, B, S, gb1);
  run<2, K, 5>("V2 +unpack/FFMA2/pack", f, gr, o, B, S, gb1);
  run<2, K, 6>("V2 +unpack/FFMA2/pack", f, gr, o, B, S, gb1);
  run<3, K, 4>("V3 lane-k prepare + shuffles", f, gr, o, B, S, gb1);
  run<3, K, 5>("V3 lane-k prepare + shuffles", f, gr, o, B, S, gb1);
  run<3, K, 6>("V3 lane-k prepare + shuffles", f, gr, o, B, S, gb1);
  run<4, K, 4>("V4 +softmax", f, gr, o, B, S, gb1);
  run<4, K, 6>("V4 +softmax", f, gr, o, B, S, gb1);
  run<5, K, 4>("V5 +streaming input loads", f, gr, o, B, S, gb1);
  run<5, K, 6>("V5 +streaming input loads", f, gr, o, B, S, gb1);
  run<6, K, 4>("V6 +any ballot +mask", f, gr, o, B, S, gb1);
  run<6, K, 5>("V6 +any ballot +mask", f, gr, o, B, S, gb1);
  run<6, K, 6>("V6 +any ballot +mask", f, gr, o, B, S, gb1);
  run<7, K, 4>("V7 any only", f, gr, o, B, S, gb1);
  run<7, K, 6>("V7 any only", f, gr, o, B, S, gb1);
  run<8, K, 4>("V8 mask only (same line as logit)", f, gr, o, B, S, gb1);
  run<8, K, 6>("V8 mask only (same line as logit)", f, gr, o, B, S, gb1);
  run<9, K, 4>("V9 mask only (own plane)", f, gr, o, B, S, gb1);
  run<9, K, 6>("V9 mask only (own plane)", f, gr, o, B, S, gb1);
  return 0;
}
