// ConvLSTM cell step, fp32 reference layout (row a13; src/convLSTM.py:41-56).
//
// CUDA-core direct convolution with the whole gate epilogue fused: cat(x, h) is never
// materialised (the input-channel loop just switches base pointer, :43), the 4*Ch
// pre-activation tensor never reaches HBM (:45-46), and the five activation / three
// multiply-add launches of :48-54 happen in registers.  This is the kernel for the
// reference's own cell sizes (Ch = 12..96 at 200^2..13^2, src/networks.py:1304-1313);
// wide cells go to the tcgen05 implicit GEMM in convlstm_tc.cu.
//
// CTA = 16x16 output pixels x 4 hidden channels (16 gate pre-activations per thread).
// Input channels are staged through shared memory 8 at a time with their halo; the weight
// slice for the CTA's 16 output channels is staged alongside and read as warp broadcasts.
#include "common.cuh"

namespace {

constexpr int TS = 16;    // tile side
constexpr int CH_T = 4;   // hidden channels per CTA
constexpr int CI_T = 8;   // input channels per smem stage
constexpr int KMAX = 7;   // largest supported kernel side

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int KS>  // KS > 0: compile-time square kernel; KS == 0: runtime kh x kw
__global__ void __launch_bounds__(TS * TS)
k_convlstm_f32(const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ c,
               const float* __restrict__ weight, const float* __restrict__ bias, int B, int Cin, int Ch, int H,
               int W, int kh_rt, int kw_rt, float* __restrict__ h_out, float* __restrict__ c_out) {
  const int kh = KS > 0 ? KS : kh_rt, kw = KS > 0 ? KS : kw_rt;
  const int ph = kh / 2, pw = kw / 2;
  const int Ct = Cin + Ch;
  const int tiles_x = (W + TS - 1) / TS;
  const int tx0 = (blockIdx.x % tiles_x) * TS, ty0 = (blockIdx.x / tiles_x) * TS;
  const int ch0 = blockIdx.y * CH_T;
  const int b = blockIdx.z;
  const int lx = threadIdx.x % TS, ly = threadIdx.x / TS;
  const int ox = tx0 + lx, oy = ty0 + ly;
  const int IW = TS + kw - 1, IH = TS + kh - 1;
  const long HW = (long)H * W;

  extern __shared__ float smem[];
  float* s_in = smem;                    // [CI_T][IH][IW]
  float* s_w = smem + CI_T * IH * IW;    // [4*CH_T][CI_T][kh*kw]

  float acc[4 * CH_T];
#pragma unroll
  for (int q = 0; q < 4 * CH_T; ++q) acc[q] = 0.f;

  for (int ci0 = 0; ci0 < Ct; ci0 += CI_T) {
    __syncthreads();
    // stage inputs (zero padding, Conv2d padding=k//2, :38)
    for (int idx = threadIdx.x; idx < CI_T * IH * IW; idx += TS * TS) {
      const int cc = idx / (IH * IW), rem = idx % (IH * IW);
      const int iy = ty0 + rem / IW - ph, ix = tx0 + rem % IW - pw;
      const int ci = ci0 + cc;
      float v = 0.f;
      if (ci < Ct && iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const float* pl = ci < Cin ? x + ((long)b * Cin + ci) * HW : h + ((long)b * Ch + (ci - Cin)) * HW;
        v = __ldg(pl + (long)iy * W + ix);
      }
      s_in[idx] = v;
    }
    // stage weights: out channel (gate g, hidden ch0+j) = g*Ch + ch0 + j  (:46 split order i,f,o,g)
    for (int idx = threadIdx.x; idx < 4 * CH_T * CI_T * kh * kw; idx += TS * TS) {
      const int q = idx / (CI_T * kh * kw), rem = idx % (CI_T * kh * kw);
      const int cc = rem / (kh * kw), t = rem % (kh * kw);
      const int g = q / CH_T, jch = ch0 + q % CH_T, ci = ci0 + cc;
      float v = 0.f;
      if (jch < Ch && ci < Ct) v = __ldg(weight + (((long)(g * Ch + jch) * Ct + ci) * kh * kw) + t);
      s_w[idx] = v;
    }
    __syncthreads();
    for (int cc = 0; cc < CI_T; ++cc) {
      const float* in = s_in + cc * IH * IW + ly * IW + lx;
      if (KS > 0) {
        float v[KS > 0 ? KS * KS : 1];
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) v[ky * KS + kx] = in[ky * IW + kx];
#pragma unroll
        for (int q = 0; q < 4 * CH_T; ++q) {
          const float* wq = s_w + (q * CI_T + cc) * KS * KS;
#pragma unroll
          for (int t = 0; t < KS * KS; ++t) acc[q] = fmaf(v[t], wq[t], acc[q]);
        }
      } else {
        for (int ky = 0; ky < kh; ++ky)
          for (int kx = 0; kx < kw; ++kx) {
            const float v = in[ky * IW + kx];
#pragma unroll
            for (int q = 0; q < 4 * CH_T; ++q) acc[q] = fmaf(v, s_w[(q * CI_T + cc) * kh * kw + ky * kw + kx], acc[q]);
          }
      }
    }
  }

  if (ox < W && oy < H) {
#pragma unroll
    for (int j = 0; j < CH_T; ++j) {
      const int ch = ch0 + j;
      if (ch >= Ch) break;
      float gi = acc[0 * CH_T + j], gf = acc[1 * CH_T + j], go = acc[2 * CH_T + j], gg = acc[3 * CH_T + j];
      if (bias) {
        gi += __ldg(bias + 0 * Ch + ch);
        gf += __ldg(bias + 1 * Ch + ch);
        go += __ldg(bias + 2 * Ch + ch);
        gg += __ldg(bias + 3 * Ch + ch);
      }
      const long o = ((long)b * Ch + ch) * HW + (long)oy * W + ox;
      const float cn = sigmoidf_acc(gf) * __ldg(c + o) + sigmoidf_acc(gi) * tanhf(gg);  // :53
      c_out[o] = cn;
      h_out[o] = sigmoidf_acc(go) * tanhf(cn);  // :54
    }
  }
}

}  // namespace

extern "C" int jaf_convlstm_step_f32(const float* x, const float* h, const float* c, const float* weight,
                                     const float* bias, int B, int Cin, int Ch, int H, int W, int kh, int kw,
                                     float* h_out, float* c_out, void* stream) {
  JAF_REQUIRE(x && h && c && weight && h_out && c_out, "null pointer");
  JAF_REQUIRE(B >= 0 && Cin > 0 && Ch > 0 && H > 0 && W > 0, "bad sizes");
  JAF_REQUIRE(kh >= 1 && kw >= 1 && kh <= KMAX && kw <= KMAX && (kh & 1) && (kw & 1), "kernel must be odd and <= 7");
  JAF_REQUIRE(B <= 65535 && (Ch + CH_T - 1) / CH_T <= 65535, "batch / channel count too large for one launch");
  if (B == 0) return JAF_OK;
  const int IW = TS + kw - 1, IH = TS + kh - 1;
  const size_t smem = sizeof(float) * ((size_t)CI_T * IH * IW + (size_t)4 * CH_T * CI_T * kh * kw);
  const dim3 grid((unsigned)(((W + TS - 1) / TS) * ((H + TS - 1) / TS)), (unsigned)((Ch + CH_T - 1) / CH_T),
                  (unsigned)B);
  cudaStream_t st = jaf::as_stream(stream);
  if (kh == 3 && kw == 3) {
    k_convlstm_f32<3><<<grid, TS * TS, smem, st>>>(x, h, c, weight, bias, B, Cin, Ch, H, W, kh, kw, h_out, c_out);
  } else if (kh == 5 && kw == 5) {
    k_convlstm_f32<5><<<grid, TS * TS, smem, st>>>(x, h, c, weight, bias, B, Cin, Ch, H, W, kh, kw, h_out, c_out);
  } else {
    JAF_CUDA(cudaFuncSetAttribute(k_convlstm_f32<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_convlstm_f32<0><<<grid, TS * TS, smem, st>>>(x, h, c, weight, bias, B, Cin, Ch, H, W, kh, kw, h_out, c_out);
  }
  return jaf::finish_launch("k_convlstm_f32");
}
