// ConvLSTM cell step, fp32 reference layout (row a13; src/convLSTM.py:41-56).
//
// CUDA-core direct convolution with the whole gate epilogue fused: cat(x, h) is never
// materialised (the input-channel loop just switches base pointer, :43), the 4*Ch
// pre-activation tensor never reaches HBM (:45-46), and the five activation / three
// multiply-add launches of :48-54 happen in registers.  This is the kernel for the
// reference's own cell sizes (Ch = 12..96 at 200^2..13^2, src/networks.py:1304-1313);
// wide cells go to the tcgen05 implicit GEMM in convlstm_tc.cu.
//
// CTA = 128 threads = a 32x16 pixel tile x 4 hidden channels (16 gate pre-activations).  Each thread
// owns 4 horizontally adjacent pixels x 16 output channels = 64 fp32 accumulators, so one 128-bit
// shared-memory read of four weights feeds 16 FMAs.  Input channels are staged through shared memory
// 8 at a time with their halo (row pitch = 1 mod 4 words: conflict-free for the 4-pixel stride);
// the weight slice is staged as [channel][tap][16 outputs].
#include "common.cuh"

namespace {

constexpr int TSX = 32, TSY = 16;  // tile
constexpr int PXT = 4;             // pixels per thread (along x)
constexpr int NTHR = (TSX / PXT) * TSY;
constexpr int CH_T = 4;            // hidden channels per CTA
constexpr int NQ = 4 * CH_T;       // gate pre-activations per pixel per CTA
constexpr int CI_T = 8;            // input channels per smem stage
constexpr int KMAX = 7;            // largest supported kernel side

__host__ __device__ inline int in_pitch(int kw) { return ((TSX + kw - 1 + 3) / 4) * 4 + 1; }

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int KS>  // KS > 0: compile-time square kernel; KS == 0: runtime kh x kw
__global__ void __launch_bounds__(NTHR)
k_convlstm_f32(const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ c,
               const float* __restrict__ weight, const float* __restrict__ bias, int B, int Cin, int Ch, int H,
               int W, int kh_rt, int kw_rt, float* __restrict__ h_out, float* __restrict__ c_out) {
  const int kh = KS > 0 ? KS : kh_rt, kw = KS > 0 ? KS : kw_rt;
  const int ph = kh / 2, pw = kw / 2;
  const int Ct = Cin + Ch;
  const int tiles_x = (W + TSX - 1) / TSX;
  const int tx0 = (blockIdx.x % tiles_x) * TSX, ty0 = (blockIdx.x / tiles_x) * TSY;
  const int ch0 = blockIdx.y * CH_T;
  const int b = blockIdx.z;
  const int lx = threadIdx.x % (TSX / PXT), ly = threadIdx.x / (TSX / PXT);
  const int IH = TSY + kh - 1, IP = in_pitch(kw), IWv = TSX + kw - 1;
  const long HW = (long)H * W;

  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                        // [CI_T][kh*kw][NQ]   (16-byte aligned rows)
  float* s_in = smem + CI_T * kh * kw * NQ;  // [CI_T][IH][IP]

  float acc[PXT][NQ];
#pragma unroll
  for (int p = 0; p < PXT; ++p)
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[p][q] = 0.f;

  for (int ci0 = 0; ci0 < Ct; ci0 += CI_T) {
    __syncthreads();
    // stage inputs (zero padding, Conv2d padding=k//2, :38); cat(x, h) is just a base-pointer switch (:43)
    for (int idx = threadIdx.x; idx < CI_T * IH * IWv; idx += NTHR) {
      const int cc = idx / (IH * IWv), rem = idx % (IH * IWv);
      const int ry = rem / IWv, rx = rem % IWv;
      const int iy = ty0 + ry - ph, ix = tx0 + rx - pw;
      const int ci = ci0 + cc;
      float v = 0.f;
      if (ci < Ct && iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const float* pl = ci < Cin ? x + ((long)b * Cin + ci) * HW : h + ((long)b * Ch + (ci - Cin)) * HW;
        v = __ldg(pl + (long)iy * W + ix);
      }
      s_in[(cc * IH + ry) * IP + rx] = v;
    }
    // stage weights: out channel (gate g, hidden ch0+j) = g*Ch + ch0 + j  (:46 split order i,f,o,g)
    for (int idx = threadIdx.x; idx < CI_T * kh * kw * NQ; idx += NTHR) {
      const int q = idx % NQ, rem = idx / NQ;
      const int t = rem % (kh * kw), cc = rem / (kh * kw);
      const int g = q / CH_T, jch = ch0 + q % CH_T, ci = ci0 + cc;
      float v = 0.f;
      if (jch < Ch && ci < Ct) v = __ldg(weight + (((long)(g * Ch + jch) * Ct + ci) * kh * kw) + t);
      s_w[idx] = v;
    }
    __syncthreads();
    for (int cc = 0; cc < CI_T; ++cc) {
      const float* in = s_in + (cc * IH + ly) * IP + lx * PXT;
      const float4* wv = reinterpret_cast<const float4*>(s_w + cc * kh * kw * NQ);
      if (KS > 0) {
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          float v[PXT + (KS > 0 ? KS : 1) - 1];
#pragma unroll
          for (int i = 0; i < PXT + KS - 1; ++i) v[i] = in[ky * IP + i];
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const float4* w4 = wv + (ky * KS + kx) * (NQ / 4);
#pragma unroll
            for (int q4 = 0; q4 < NQ / 4; ++q4) {
              const float4 wq = w4[q4];
#pragma unroll
              for (int p = 0; p < PXT; ++p) {
                acc[p][4 * q4 + 0] = fmaf(v[p + kx], wq.x, acc[p][4 * q4 + 0]);
                acc[p][4 * q4 + 1] = fmaf(v[p + kx], wq.y, acc[p][4 * q4 + 1]);
                acc[p][4 * q4 + 2] = fmaf(v[p + kx], wq.z, acc[p][4 * q4 + 2]);
                acc[p][4 * q4 + 3] = fmaf(v[p + kx], wq.w, acc[p][4 * q4 + 3]);
              }
            }
          }
        }
      } else {
        for (int ky = 0; ky < kh; ++ky)
          for (int kx = 0; kx < kw; ++kx) {
            const float4* w4 = wv + (ky * kw + kx) * (NQ / 4);
            float v[PXT];
#pragma unroll
            for (int p = 0; p < PXT; ++p) v[p] = in[ky * IP + kx + p];
#pragma unroll
            for (int q4 = 0; q4 < NQ / 4; ++q4) {
              const float4 wq = w4[q4];
#pragma unroll
              for (int p = 0; p < PXT; ++p) {
                acc[p][4 * q4 + 0] = fmaf(v[p], wq.x, acc[p][4 * q4 + 0]);
                acc[p][4 * q4 + 1] = fmaf(v[p], wq.y, acc[p][4 * q4 + 1]);
                acc[p][4 * q4 + 2] = fmaf(v[p], wq.z, acc[p][4 * q4 + 2]);
                acc[p][4 * q4 + 3] = fmaf(v[p], wq.w, acc[p][4 * q4 + 3]);
              }
            }
          }
      }
    }
  }

  const int oy = ty0 + ly;
  if (oy < H) {
#pragma unroll
    for (int j = 0; j < CH_T; ++j) {
      const int ch = ch0 + j;
      if (ch >= Ch) break;
      float bi = 0.f, bf = 0.f, bo = 0.f, bg = 0.f;
      if (bias) {
        bi = __ldg(bias + 0 * Ch + ch);
        bf = __ldg(bias + 1 * Ch + ch);
        bo = __ldg(bias + 2 * Ch + ch);
        bg = __ldg(bias + 3 * Ch + ch);
      }
#pragma unroll
      for (int p = 0; p < PXT; ++p) {
        const int ox = tx0 + lx * PXT + p;
        if (ox < W) {
          const long o = ((long)b * Ch + ch) * HW + (long)oy * W + ox;
          const float gi = acc[p][0 * CH_T + j] + bi, gf = acc[p][1 * CH_T + j] + bf;
          const float go = acc[p][2 * CH_T + j] + bo, gg = acc[p][3 * CH_T + j] + bg;
          const float cn = sigmoidf_acc(gf) * __ldg(c + o) + sigmoidf_acc(gi) * tanhf(gg);  // :53
          c_out[o] = cn;
          h_out[o] = sigmoidf_acc(go) * tanhf(cn);  // :54
        }
      }
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
  const int IH = TSY + kh - 1, IP = in_pitch(kw);
  const size_t smem = sizeof(float) * ((size_t)CI_T * kh * kw * NQ + (size_t)CI_T * IH * IP);
  const dim3 grid((unsigned)(((W + TSX - 1) / TSX) * ((H + TSY - 1) / TSY)), (unsigned)((Ch + CH_T - 1) / CH_T),
                  (unsigned)B);
  cudaStream_t st = jaf::as_stream(stream);
  if (kh == 3 && kw == 3) {
    k_convlstm_f32<3><<<grid, NTHR, smem, st>>>(x, h, c, weight, bias, B, Cin, Ch, H, W, kh, kw, h_out, c_out);
  } else if (kh == 5 && kw == 5) {
    k_convlstm_f32<5><<<grid, NTHR, smem, st>>>(x, h, c, weight, bias, B, Cin, Ch, H, W, kh, kw, h_out, c_out);
  } else {
    JAF_CUDA(cudaFuncSetAttribute(k_convlstm_f32<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_convlstm_f32<0><<<grid, NTHR, smem, st>>>(x, h, c, weight, bias, B, Cin, Ch, H, W, kh, kw, h_out, c_out);
  }
  return jaf::finish_launch("k_convlstm_f32");
}
