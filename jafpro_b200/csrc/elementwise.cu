// Streaming elementwise pieces of the path (rows a11, a12): visibility mask + confidence blend
// (src/flow_net.py:91,98) and the standalone softmax-over-K reduction (src/networks.py:1264-1286).
// Pure HBM-bound passes: every byte read once with L1-bypassing loads, written once with
// streaming stores; float4 vectors when the plane size allows it.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kMaxK = 32;

template <int VEC>
struct VecT;
template <>
struct VecT<1> {
  using type = float;
};
template <>
struct VecT<4> {
  using type = float4;
};

template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float* v) {
  if constexpr (VEC == 4) {
    float4 q;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                 : "l"(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    v[0] = ld_stream_f32(p);
  }
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float* v) {
  if constexpr (VEC == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  } else {
    __stcs(p, v[0]);
  }
}

// one thread per VEC consecutive pixels of one (b, c) plane
template <int VEC>
__global__ void __launch_bounds__(256)
k_mask_blend(const float* __restrict__ fake, const float* __restrict__ tsf, const float* __restrict__ mask,
             int mask_c, const float* __restrict__ conf, int B, int C, long HW, float* __restrict__ masked_out,
             float* __restrict__ pred) {
  const long nvec = HW / VEC;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * C * nvec) return;
  const long p = (i % nvec) * VEC;
  const int c = (int)((i / nvec) % C);
  const int b = (int)(i / (nvec * C));
  const long o = ((long)b * C + c) * HW + p;
  float t[VEC], m[VEC], w[VEC], f[VEC], r[VEC];
  ldv<VEC>(tsf + o, t);
  if (mask) {
    ldv<VEC>(mask + ((long)b * mask_c + (mask_c == 1 ? 0 : c)) * HW + p, m);
#pragma unroll
    for (int q = 0; q < VEC; ++q) t[q] = t[q] * m[q];  // flow_net.py:91
  }
  if (masked_out) stv<VEC>(masked_out + o, t);
  if (pred) {
    ldv<VEC>(conf + (long)b * HW + p, w);
    ldv<VEC>(fake + o, f);
#pragma unroll
    for (int q = 0; q < VEC; ++q) r[q] = __fadd_rn(__fmul_rn(f[q], w[q]), __fmul_rn(t[q], __fsub_rn(1.0f, w[q])));  // :98
    stv<VEC>(pred + o, r);
  }
}

// one thread per VEC consecutive pixels; loops over channels, K weights kept in registers
template <int VEC, int KT>
__global__ void __launch_bounds__(256)
k_softmax_fuse(const float* __restrict__ feat, const float* __restrict__ logits, int B, int C, long HW,
               float* __restrict__ out) {
  const long nvec = HW / VEC;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * nvec) return;
  const long p = (i % nvec) * VEC;
  const int b = (int)(i / nvec);
  float a[KT][VEC];
  float m[VEC], s[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { m[q] = -CUDART_INF_F; s[q] = 0.f; }
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    ldv<VEC>(logits + ((long)b * KT + k) * HW + p, a[k]);
#pragma unroll
    for (int q = 0; q < VEC; ++q) m[q] = fmaxf(m[q], a[k][q]);
  }
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      a[k][q] = expf(a[k][q] - m[q]);
      s[q] += a[k][q];
    }
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int q = 0; q < VEC; ++q) a[k][q] = a[k][q] / s[q];
  for (int c = 0; c < C; ++c) {
    float acc[VEC], v[VEC];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      ldv<VEC>(feat + ((long)b * KT * C + (long)k * C + c) * HW + p, v);
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        const float pr = __fmul_rn(v[q], a[k][q]);  // networks.py:1276-1280 (separate multiply)
        acc[q] = (k == 0) ? pr : __fadd_rn(acc[q], pr);  // :1282-1286 (sliced adds, left to right)
      }
    }
    stv<VEC>(out + ((long)b * C + c) * HW + p, acc);
  }
}

template <int VEC>
int launch_softmax_fuse(const float* feat, const float* logits, int B, int K, int C, long HW, float* out,
                        cudaStream_t st) {
  const long n = (long)B * (HW / VEC);
  const int grid = jaf::ceil_div(n, 256);
#define JAF_SF(KV) \
  case KV: k_softmax_fuse<VEC, KV><<<grid, 256, 0, st>>>(feat, logits, B, C, HW, out); return 1;
  switch (K) {
    JAF_SF(1) JAF_SF(2) JAF_SF(3) JAF_SF(4) JAF_SF(5) JAF_SF(6) JAF_SF(7) JAF_SF(8)
    default: return 0;
  }
#undef JAF_SF
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

int jaf_mask_blend(const float* fake, const float* tsf, const float* mask, int mask_c, const float* conf, int B,
                   int C, int H, int W, float* masked_out, float* pred, void* stream) {
  JAF_REQUIRE(tsf != nullptr, "tsf is required");
  JAF_REQUIRE(masked_out || pred, "no output requested");
  JAF_REQUIRE(!pred || (fake && conf), "pred needs fake and conf");
  JAF_REQUIRE(!mask || mask_c == 1 || mask_c == C, "mask_c must be 1 or C");
  JAF_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "bad sizes");
  if (B == 0) return JAF_OK;
  const long HW = (long)H * W;
  cudaStream_t st = jaf::as_stream(stream);
  const bool v4 = (HW % 4 == 0) && aligned16(tsf) && aligned16(fake) && aligned16(mask) && aligned16(conf) &&
                  aligned16(masked_out) && aligned16(pred);
  if (v4) {
    k_mask_blend<4><<<jaf::ceil_div((long)B * C * (HW / 4), 256), 256, 0, st>>>(fake, tsf, mask, mask_c, conf, B, C,
                                                                               HW, masked_out, pred);
  } else {
    k_mask_blend<1><<<jaf::ceil_div((long)B * C * HW, 256), 256, 0, st>>>(fake, tsf, mask, mask_c, conf, B, C, HW,
                                                                         masked_out, pred);
  }
  return jaf::finish_launch("k_mask_blend");
}

int jaf_softmax_fuse(const float* feat, const float* logits, int B, int K, int C, int H, int W, float* out,
                     void* stream) {
  JAF_REQUIRE(feat && logits && out, "null pointer");
  JAF_REQUIRE(B >= 0 && K >= 1 && K <= 8 && C > 0 && H > 0 && W > 0, "bad sizes (1 <= K <= 8)");
  if (B == 0) return JAF_OK;
  const long HW = (long)H * W;
  cudaStream_t st = jaf::as_stream(stream);
  const bool v4 = (HW % 4 == 0) && aligned16(feat) && aligned16(logits) && aligned16(out);
  const int ok = v4 ? launch_softmax_fuse<4>(feat, logits, B, K, C, HW, out, st)
                    : launch_softmax_fuse<1>(feat, logits, B, K, C, HW, out, st);
  JAF_REQUIRE(ok, "unsupported K");
  return jaf::finish_launch("k_softmax_fuse");
}

}  // extern "C"
