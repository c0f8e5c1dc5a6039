// Fused bilinear backward warp of K references + visibility / softmax-weighted fusion (row F;
// rows a10-a12 of SURVEY.md §8a).  See include/jafpro_b200.h for the contract.
//
// Two kernels:
//   k_warp_fuse_nhwc  — the hot kernel.  Features are channels-last bf16, so one bilinear tap of one
//       reference is C*2 contiguous bytes (128 B = one cache line at C = 64).  A group of LPP = C/8
//       lanes owns one output pixel; each lane moves 16 bytes (8 channels) per tap with one 128-bit
//       load, so a warp-wide load instruction fetches whole lines.  Lane k of a group builds
//       reference k's sample position, bilinear weights and softmax term; the group exchanges them
//       with warp shuffles and the reduction over K happens in registers.  A CTA sweeps a 8*PPW-pixel
//       wide strip downwards, so the lower tap row of one step is the upper row of the next and is
//       served by L1; flows, logits, masks are read once with L1-bypassing loads and outputs are
//       written once with streaming stores.  RGB (planar f32, the reference layout) rides along on
//       the first three lanes of each group.
//   k_warp_fuse_generic — one thread per pixel, any layout / dtype / K <= 16: the reference's NCHW
//       fp32 call sites (warp_image, crn_model feature warps), per-reference warped outputs, and
//       every shape the hot kernel does not cover.
//
// Arithmetic follows ATen's grid_sampler_2d (bilinear, border): unnormalise, clamp the
// unnormalised coordinate to [0, size-1], floor, four corner weights, nw/ne/sw/se accumulation
// (ATen/native/cuda/GridSampler.cuh:14-45 and the kernel body in GridSampler.cu).
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "raster_math.cuh"

namespace {

constexpr int kMaxKGeneric = 16;

struct Tap {
  int off;  // y0 * Ws + x0
  int dx;   // 0 or 1   (0 when x0 + 1 is outside: that tap has weight exactly 0)
  int dy;   // 0 or Ws
  float nw, ne, sw, se;
};

// Every operation is pinned (no compiler-chosen contraction): (c + 1) * size - 1 is the one place
// where nvcc fuses ATen's expression into an FMA, and the oracle restates exactly that.
__device__ __forceinline__ float unnormalize(float c, int size, int align_corners) {
  if (align_corners) return __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), 0.5f), (float)(size - 1));
  return __fmul_rn(__fmaf_rn(__fadd_rn(c, 1.f), (float)size, -1.f), 0.5f);
}

__device__ __forceinline__ Tap make_tap(float gx, float gy, int Ws, int Hs, int align_corners) {
  float ix = unnormalize(gx, Ws, align_corners);
  float iy = unnormalize(gy, Hs, align_corners);
  ix = fminf((float)(Ws - 1), fmaxf(ix, 0.f));  // clip_coordinates (NaN -> 0)
  iy = fminf((float)(Hs - 1), fmaxf(iy, 0.f));
  const float fx = floorf(ix), fy = floorf(iy);
  const float x1 = fx + 1.f, y1 = fy + 1.f;  // exact (small integers)
  Tap t;
  const int x0 = (int)fx, y0 = (int)fy;
  t.off = y0 * Ws + x0;
  t.dx = (x0 + 1 < Ws) ? 1 : 0;
  t.dy = (y0 + 1 < Hs) ? Ws : 0;
  const float ax = __fsub_rn(x1, ix), bx = __fsub_rn(ix, fx), ay = __fsub_rn(y1, iy), by = __fsub_rn(iy, fy);
  t.nw = __fmul_rn(ax, ay);
  t.ne = __fmul_rn(bx, ay);
  t.sw = __fmul_rn(ax, by);
  t.se = __fmul_rn(bx, by);
  return t;
}

struct WFArgs {
  int B, K, H, W, Hs, Ws, C, align_corners, mask_c;
  int rows_per_cta, tiles_x, tiles_y;
  int c_chunk;  // generic kernel: channels per blockIdx.y slice
  const float* rgb;
  const void* feat;
  const int* ref_index;
  const float* grid;
  const float* logits;
  const float* vis;
  const int* fim;
  const float* tgt_mask;
  const float* fake;
  const float* conf;
  float* out_rgb;
  void* out_feat;
  float* warped_rgb;
  // pose-driven flavour (jaf_warp_fuse_from_poses): the transfer flows are composed per tile from the target pose's
  // z-buffer keys and the poses, in shared memory, instead of being read from `grid`
  unsigned long long* zkeys;        // [B,S,S] (raster rows, i.e. not flipped)
  int leave_clean;                  // reset every consumed key to "empty" (the caller skips the next clear)
  const float* tgt_cam;             // [B,3]
  const float* tgt_verts;           // [B,V,3]
  const float* src_cam;             // [R,K,3]
  const float* src_verts;           // [R,K,V,3]
  const int* fidx;                  // [F,3]
  int V;
  float eye_z, near_, far_;
  int* fim_out;                     // [B,H,W] or nullptr
  float* T_out;                     // [B,K,H,W,2] or nullptr
};

// =====================================================================================
// Hot kernel: channels-last bf16 features (+ optional planar f32 RGB)
// =====================================================================================
// Sample position for the hot kernel: identical values to make_tap(), but the integer corner is
// clamped to [0, size-2] so that all four taps are always in bounds (when the clamped coordinate sits
// exactly on the last row/column the weights become (0, 1) instead of (1, skipped) — the same sum,
// bit for bit — and no per-tap border flags have to travel with the offset).
struct HotTap {
  int off;  // y0 * Ws + x0 with x0 <= Ws-2, y0 <= Hs-2
  float nw, ne, sw, se;
};
__device__ __forceinline__ HotTap make_hot_tap(float gx, float gy, int Ws, int Hs, int align_corners) {
  float ix = unnormalize(gx, Ws, align_corners);
  float iy = unnormalize(gy, Hs, align_corners);
  ix = fminf((float)(Ws - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(Hs - 1), fmaxf(iy, 0.f));
  const float fx = fminf(floorf(ix), (float)(Ws - 2)), fy = fminf(floorf(iy), (float)(Hs - 2));
  const float ax = __fsub_rn(fx + 1.f, ix), bx = __fsub_rn(ix, fx);
  const float ay = __fsub_rn(fy + 1.f, iy), by = __fsub_rn(iy, fy);
  HotTap t;
  t.off = (int)fy * Ws + (int)fx;
  t.nw = __fmul_rn(ax, ay);
  t.ne = __fmul_rn(bx, ay);
  t.sw = __fmul_rn(ax, by);
  t.se = __fmul_rn(bx, by);
  return t;
}

// Stand-alone flavour of rgb_pixel for the RGB-only kernel: constants hoisted, one reciprocal instead of K
// divisions, byte addressing.  (Inlined into the fused kernel it costs phase A 2 % through register
// allocation, measured, so the fused kernel keeps the plain version below.)
// One output pixel of the RGB planes (planar fp32, the reference layout): softmax in the reference's
// sequential order, ATen's tap order, acc += w_k * warped_k — bit-identical to k_warp_fuse_generic.
// XSHARE (whole warp converged, lanes = x-adjacent pixels; `live` = this lane owns a real pixel): the ne / se taps of a
// pixel are the nw / sw taps of its right-hand neighbour whenever the neighbour's corner is one column further (true for
// most pixels of a locally translation-like flow) — they come from the neighbour lane by shuffle instead of a second,
// line-straddling gather (same address, same bits).  Lanes whose neighbour samples elsewhere load the taps themselves.
template <int KT, bool SKIP, bool XSHARE = false>
__device__ __forceinline__ void rgb_pixel_lean(const WFArgs& a, const float* __restrict__ rgb_base,
                                          const float2* __restrict__ b_grid, const float* __restrict__ b_logit,
                                          const float* __restrict__ b_vis, const int* __restrict__ b_fim,
                                          const float* __restrict__ b_mask, const float* __restrict__ b_fake,
                                          const float* __restrict__ b_conf, float* __restrict__ b_orgb, unsigned pix,
                                          unsigned HW, unsigned HWs, unsigned Ws, bool live = true) {
  // every load that does not depend on another one is issued first: sample positions, logits, target mask
  float2 gxy0[KT];
  if constexpr (!SKIP && KT <= 4) {
#pragma unroll
    for (int k = 0; k < KT; ++k) gxy0[k] = __ldg(b_grid + ((unsigned)k * HW + pix));
  }
  constexpr bool kEarly = !SKIP && KT <= 4;
  float tm = 1.f;
  if (kEarly && b_mask) tm = __ldg(b_mask + pix);
  // softmax in the reference order: max, exp, running sum; one correctly-rounded reciprocal replaces the
  // K divisions (identical for K = 1, within 1 ulp otherwise)
  float aw[KT];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    aw[k] = b_logit ? __ldg(b_logit + ((unsigned)k * HW + pix)) : 0.f;
    m = fmaxf(m, aw[k]);
  }
  float ssum = 0.f;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    aw[k] = expf(aw[k] - m);
    ssum += aw[k];
  }
  const float inv = (KT == 1) ? 1.0f : __frcp_rn(ssum);
  const float vf = b_fim ? ((__ldg(b_fim + pix) != -1) ? 1.f : 0.f) : 1.f;
  // constants of the sample-position arithmetic, once per pixel instead of once per reference
  const int Hs = a.Hs;
  const float wsf = (float)Ws, hsf = (float)Hs, wm1 = (float)(Ws - 1), hm1 = (float)(Hs - 1);
  const float wm2 = (float)(Ws - 2), hm2 = (float)(Hs - 2);
  const bool ac = a.align_corners != 0;
  const char* __restrict__ rgb_bytes = reinterpret_cast<const char*>(rgb_base);
  float acc[3] = {0.f, 0.f, 0.f};
  if constexpr (!SKIP && KT <= 4) {  // beyond 4 references the taps no longer fit the register budget
    // No visibility input: nothing is skipped, so the whole pixel is software-pipelined by hand — all K sample
    // positions are loaded up front, all K taps are built, and the 12 gathers of reference k+1 are in flight while
    // reference k is reduced.  (Left to itself the compiler serialises load -> taps -> gathers -> FMAs per reference:
    // 2K exposed memory latencies per pixel, measured as 20 stall cycles per issued instruction.)
    const float2* gxy = gxy0;
    unsigned ok[KT];
    float tw[KT][4], wk[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      wk[k] = (KT == 1) ? aw[k] : (aw[k] * inv);
      float ix = ac ? __fmul_rn(__fmul_rn(__fadd_rn(gxy[k].x, 1.f), 0.5f), wm1) : __fmul_rn(__fmaf_rn(__fadd_rn(gxy[k].x, 1.f), wsf, -1.f), 0.5f);
      float iy = ac ? __fmul_rn(__fmul_rn(__fadd_rn(gxy[k].y, 1.f), 0.5f), hm1) : __fmul_rn(__fmaf_rn(__fadd_rn(gxy[k].y, 1.f), hsf, -1.f), 0.5f);
      ix = fminf(wm1, fmaxf(ix, 0.f));
      iy = fminf(hm1, fmaxf(iy, 0.f));
      const float fx = fminf(floorf(ix), wm2), fy = fminf(floorf(iy), hm2);
      const float ax = __fsub_rn(fx + 1.f, ix), bx = __fsub_rn(ix, fx), ay = __fsub_rn(fy + 1.f, iy), by = __fsub_rn(iy, fy);
      tw[k][0] = __fmul_rn(ax, ay); tw[k][1] = __fmul_rn(bx, ay); tw[k][2] = __fmul_rn(ax, by); tw[k][3] = __fmul_rn(bx, by);
      ok[k] = (unsigned)(k * 3) * HWs + (unsigned)((int)fy * (int)Ws + (int)fx);
    }
    float vbuf[2][12];
    bool coh[KT];  // XSHARE: the right-hand neighbour lane's nw / sw taps are this pixel's ne / se taps
    if constexpr (XSHARE) {
#pragma unroll
      for (int k = 0; k < KT; ++k)
        coh[k] = (__shfl_down_sync(0xffffffffu, ok[k], 1) == ok[k] + 1u) && ((threadIdx.x & 31u) != 31u);
    }
    auto gather = [&](int k, float* v) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* p0 = reinterpret_cast<const float*>(rgb_bytes + (size_t)(ok[k] + (unsigned)c * HWs) * 4u);
        const float* p1 = reinterpret_cast<const float*>(rgb_bytes + (size_t)(ok[k] + (unsigned)c * HWs + Ws) * 4u);
        v[4 * c + 0] = __ldg(p0);
        v[4 * c + 2] = __ldg(p1);
        if constexpr (XSHARE) {
          v[4 * c + 1] = 0.f;
          v[4 * c + 3] = 0.f;
          if (!coh[k]) {
            v[4 * c + 1] = __ldg(p0 + 1);
            v[4 * c + 3] = __ldg(p1 + 1);
          }
        } else {
          v[4 * c + 1] = __ldg(p0 + 1);
          v[4 * c + 3] = __ldg(p1 + 1);
        }
      }
    };
    gather(0, vbuf[0]);
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k + 1 < KT) gather(k + 1, vbuf[(k + 1) & 1]);
      const float* v = vbuf[k & 1];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float ne = v[4 * c + 1], se = v[4 * c + 3];
        if constexpr (XSHARE) {
          const float ne_s = __shfl_down_sync(0xffffffffu, v[4 * c + 0], 1), se_s = __shfl_down_sync(0xffffffffu, v[4 * c + 2], 1);
          if (coh[k]) {
            ne = ne_s;
            se = se_s;
          }
        }
        float sacc = fmaf(v[4 * c + 0], tw[k][0], 0.f);
        sacc = fmaf(ne, tw[k][1], sacc);
        sacc = fmaf(v[4 * c + 2], tw[k][2], sacc);
        sacc = fmaf(se, tw[k][3], sacc);
        acc[c] = fmaf(wk[k], sacc, acc[c]);
      }
    }
  } else {
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const float v = b_vis ? __ldg(b_vis + ((unsigned)k * HW + pix)) : vf;
    const float w = (KT == 1) ? aw[k] * v : (aw[k] * inv) * v;
    if (!SKIP || w != 0.f) {  // without a visibility input nothing is skipped: no branch, loads of all k overlap
      const float2 gxy = __ldg(b_grid + ((unsigned)k * HW + pix));
      // == make_hot_tap (same pinned operations)
      float ix = ac ? __fmul_rn(__fmul_rn(__fadd_rn(gxy.x, 1.f), 0.5f), wm1) : __fmul_rn(__fmaf_rn(__fadd_rn(gxy.x, 1.f), wsf, -1.f), 0.5f);
      float iy = ac ? __fmul_rn(__fmul_rn(__fadd_rn(gxy.y, 1.f), 0.5f), hm1) : __fmul_rn(__fmaf_rn(__fadd_rn(gxy.y, 1.f), hsf, -1.f), 0.5f);
      ix = fminf(wm1, fmaxf(ix, 0.f));
      iy = fminf(hm1, fmaxf(iy, 0.f));
      const float fx = fminf(floorf(ix), wm2), fy = fminf(floorf(iy), hm2);
      const float ax = __fsub_rn(fx + 1.f, ix), bx = __fsub_rn(ix, fx), ay = __fsub_rn(fy + 1.f, iy), by = __fsub_rn(iy, fy);
      const float nw = __fmul_rn(ax, ay), ne = __fmul_rn(bx, ay), sw = __fmul_rn(ax, by), se = __fmul_rn(bx, by);
      const unsigned ok = (unsigned)(k * 3) * HWs + (unsigned)((int)fy * (int)Ws + (int)fx);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* p0 = reinterpret_cast<const float*>(rgb_bytes + (size_t)(ok + (unsigned)c * HWs) * 4u);
        const float* p1 = reinterpret_cast<const float*>(rgb_bytes + (size_t)(ok + (unsigned)c * HWs + Ws) * 4u);
        float s = fmaf(__ldg(p0), nw, 0.f);
        s = fmaf(__ldg(p0 + 1), ne, s);
        s = fmaf(__ldg(p1), sw, s);
        s = fmaf(__ldg(p1 + 1), se, s);
        acc[c] = fmaf(w, s, acc[c]);
      }
    }
  }
  }
  if (!kEarly && b_mask) tm = __ldg(b_mask + pix);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float ov = acc[c];
    if (b_mask) ov *= (a.mask_c == 3) ? __ldg(b_mask + ((unsigned)c * HW + pix)) : tm;
    if (b_fake) {
      const float wc = __ldg(b_conf + pix);
      const float fkv = __ldg(b_fake + ((unsigned)c * HW + pix));
      ov = fkv * wc + ov * (1.0f - wc);  // src/flow_net.py:98
    }
    if (live) st_stream_f32(b_orgb + ((unsigned)c * HW + pix), ov);
  }
}

// One output pixel of the RGB planes (planar fp32, the reference layout): softmax in the reference's
// sequential order, ATen's tap order, acc += w_k * warped_k (the visibility-skipping flavour is bit-identical to
// k_warp_fuse_generic; the pipelined one replaces the K softmax divisions by one reciprocal, <= 1 ulp).
// s_grid != nullptr: the flows of this tile live in shared memory (s_grid[k * s_kstride + s_pix], composed from the
// poses by the kernel itself) and the pixel's visibility is known (s_vis 0 / 1) instead of coming from fim / vis.
// Measured (same box, profiles/r02_bench_ab.jsonl): sharing helps the stand-alone RGB kernel (451 k -> 477 k frames/s dense,
// 424 k -> 440 k hard) and costs the fused kernels 2-4 % (their phase B competes with phase A of the co-resident CTAs for
// issue slots, and the shuffles + selects add ~36 instructions per pixel) — so the fused kernels keep their own gathers.
constexpr bool kShareRgbTaps = false;
// XSHARE / live: see rgb_pixel_lean (only the !SKIP pipelined flavour shares taps; the caller keeps the warp converged).
template <int KT, bool SKIP, bool PIPE = (KT <= 4), bool XSHARE = false>  // PIPE: hand-pipelined flavour (needs ~2*KT + 45 registers)
__device__ __forceinline__ void rgb_pixel(const WFArgs& a, const float* __restrict__ rgb_base,
                                          const float2* __restrict__ b_grid, const float* __restrict__ b_logit,
                                          const float* __restrict__ b_vis, const int* __restrict__ b_fim,
                                          const float* __restrict__ b_mask, const float* __restrict__ b_fake,
                                          const float* __restrict__ b_conf, float* __restrict__ b_orgb, unsigned pix,
                                          unsigned HW, unsigned HWs, unsigned Ws, const float2* s_grid = nullptr,
                                          unsigned s_kstride = 0, unsigned s_pix = 0, int s_vis = 1, bool live = true) {
  auto grid_at = [&](int k) -> float2 {
    if (s_grid != nullptr) return s_grid[(unsigned)k * s_kstride + s_pix];
    return __ldg(b_grid + ((unsigned)k * HW + pix));
  };
  float2 gxy0[KT];
  if constexpr (!SKIP && PIPE) {
#pragma unroll
    for (int k = 0; k < KT; ++k) gxy0[k] = grid_at(k);
  }
  constexpr bool kEarly = !SKIP && PIPE;
  float tm = 1.f;
  if (kEarly && b_mask) tm = __ldg(b_mask + pix);
  if constexpr (SKIP) {
    // pixel-level visibility (fim): a background pixel contributes nothing whatever the logits are — write the empty
    // result (or the blend with it) without touching logits, flows or references
    if ((s_grid != nullptr && s_vis == 0) ||
        (s_grid == nullptr && b_fim != nullptr && b_vis == nullptr && __ldg(b_fim + pix) == -1)) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float ov = 0.f;
        if (b_mask) ov *= (a.mask_c == 3) ? __ldg(b_mask + ((unsigned)c * HW + pix)) : __ldg(b_mask + pix);
        if (b_fake) {
          const float wc = __ldg(b_conf + pix);
          const float fkv = __ldg(b_fake + ((unsigned)c * HW + pix));
          ov = fkv * wc + ov * (1.0f - wc);  // src/flow_net.py:98
        }
        st_stream_f32(b_orgb + ((unsigned)c * HW + pix), ov);
      }
      return;
    }
  }
  // softmax in the reference order: max, exp, running sum, divide
  float aw[KT];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    aw[k] = b_logit ? __ldg(b_logit + ((unsigned)k * HW + pix)) : 0.f;
    m = fmaxf(m, aw[k]);
  }
  float ssum = 0.f;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    aw[k] = expf(aw[k] - m);
    ssum += aw[k];
  }
  const float vf = (s_grid != nullptr) ? (float)s_vis : (b_fim ? ((__ldg(b_fim + pix) != -1) ? 1.f : 0.f) : 1.f);
  float acc[3] = {0.f, 0.f, 0.f};
  if constexpr (!SKIP && PIPE) {  // beyond 4 references the taps no longer fit the register budget
    // hand-pipelined like rgb_pixel_lean: all K sample positions first, then the gathers of reference k+1 in flight
    // while reference k is reduced (same operations in the same order: bit-identical to the generic kernel)
    const float2* gxy = gxy0;
    // taps are built one reference ahead of their use (two live sets instead of K: no register spills at 64)
    float vbuf[2][12];
    auto tap_of = [&](int k) {
      HotTap t = make_hot_tap(gxy[k].x, gxy[k].y, (int)Ws, a.Hs, a.align_corners);
      t.off += k * 3 * (int)HWs;
      return t;
    };
    bool coh[2] = {false, false};  // XSHARE: the right-hand neighbour lane's nw / sw taps are this pixel's ne / se taps
    auto gather = [&](const HotTap& t, float* v, bool& co) {
      if constexpr (XSHARE) co = (__shfl_down_sync(0xffffffffu, t.off, 1) == t.off + 1) && ((threadIdx.x & 31u) != 31u);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* p0 = rgb_base + ((unsigned)t.off + (unsigned)c * HWs);
        v[4 * c + 0] = __ldg(p0);
        v[4 * c + 2] = __ldg(p0 + Ws);
        if constexpr (XSHARE) {
          v[4 * c + 1] = 0.f;
          v[4 * c + 3] = 0.f;
          if (!co) {
            v[4 * c + 1] = __ldg(p0 + 1);
            v[4 * c + 3] = __ldg(p0 + Ws + 1);
          }
        } else {
          v[4 * c + 1] = __ldg(p0 + 1);
          v[4 * c + 3] = __ldg(p0 + Ws + 1);
        }
      }
    };
    const float inv = (KT == 1) ? 1.0f : __frcp_rn(ssum);  // one correctly-rounded reciprocal instead of K divisions
    HotTap tn = tap_of(0);
    gather(tn, vbuf[0], coh[0]);
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const HotTap tc = tn;
      if (k + 1 < KT) {
        tn = tap_of(k + 1);
        gather(tn, vbuf[(k + 1) & 1], coh[(k + 1) & 1]);
      }
      const float* v = vbuf[k & 1];
      const float w = (KT == 1) ? aw[k] : (aw[k] * inv);  // vf == 1 here (no visibility input)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float ne = v[4 * c + 1], se = v[4 * c + 3];
        if constexpr (XSHARE) {
          const float ne_s = __shfl_down_sync(0xffffffffu, v[4 * c + 0], 1), se_s = __shfl_down_sync(0xffffffffu, v[4 * c + 2], 1);
          if (coh[k & 1]) {
            ne = ne_s;
            se = se_s;
          }
        }
        float sacc = fmaf(v[4 * c + 0], tc.nw, 0.f);
        sacc = fmaf(ne, tc.ne, sacc);
        sacc = fmaf(v[4 * c + 2], tc.sw, sacc);
        sacc = fmaf(se, tc.se, sacc);
        acc[c] = fmaf(w, sacc, acc[c]);
      }
    }
  } else {
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const float v = b_vis ? __ldg(b_vis + ((unsigned)k * HW + pix)) : vf;
    const float w = (aw[k] / ssum) * v;
    if (!SKIP || w != 0.f) {  // without a visibility input nothing is skipped: no branch, loads of all k overlap
      const float2 gxy = grid_at(k);
      const HotTap t = make_hot_tap(gxy.x, gxy.y, (int)Ws, a.Hs, a.align_corners);
      const unsigned ok = (unsigned)(k * 3) * HWs + (unsigned)t.off;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* p0 = rgb_base + (ok + (unsigned)c * HWs);
        const float* p1 = rgb_base + (ok + (unsigned)c * HWs + Ws);
        float s = fmaf(__ldg(p0), t.nw, 0.f);
        s = fmaf(__ldg(p0 + 1), t.ne, s);
        s = fmaf(__ldg(p1), t.sw, s);
        s = fmaf(__ldg(p1 + 1), t.se, s);
        acc[c] = fmaf(w, s, acc[c]);
      }
    }
  }
  }
  if (!kEarly && b_mask) tm = __ldg(b_mask + pix);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float ov = acc[c];
    if (b_mask) ov *= (a.mask_c == 3) ? __ldg(b_mask + ((unsigned)c * HW + pix)) : tm;
    if (b_fake) {
      const float wc = __ldg(b_conf + pix);
      const float fkv = __ldg(b_fake + ((unsigned)c * HW + pix));
      ov = fkv * wc + ov * (1.0f - wc);  // src/flow_net.py:98
    }
    if (live) st_stream_f32(b_orgb + ((unsigned)c * HW + pix), ov);
  }
}

// A tile without a single visible pixel: every output is the empty result (0, or the confidence blend with 0).
// Vectorised fill by the whole CTA, no row loop, no flow / logit / reference read.
template <int LPP, int TW>
__device__ __forceinline__ void fill_empty_tile(const WFArgs& a, int b, int tx, int y_begin, int y_end, unsigned W, unsigned HW) {
  const unsigned rows = (unsigned)(y_end - y_begin);
  const unsigned wpx = min((unsigned)TW, W - (unsigned)tx * TW);  // tile width inside the frame
  if (a.out_feat != nullptr) {
    const unsigned per_row = wpx * LPP;                            // uint4 per tile row
    uint4* o_base = reinterpret_cast<uint4*>(a.out_feat) + ((size_t)b * HW + (size_t)y_begin * W + (size_t)tx * TW) * LPP;
    if (TW * LPP == 256 && wpx == (unsigned)TW) {
      // a full-width tile row is exactly 256 x 16 bytes: one store per thread and row, no index arithmetic
      uint4* o_t = o_base + threadIdx.x;
      for (unsigned ry = 0; ry < rows; ++ry, o_t += (size_t)W * LPP) st_stream_u128(o_t, make_uint4(0u, 0u, 0u, 0u));
    } else
    for (unsigned i = threadIdx.x; i < rows * per_row; i += 256) {
      const unsigned ry = i / per_row, rx = i - ry * per_row;
      st_stream_u128(o_base + (size_t)ry * W * LPP + rx, make_uint4(0u, 0u, 0u, 0u));
    }
  }
  if (a.rgb != nullptr && a.out_rgb != nullptr) {
    const bool blend = a.fake != nullptr && a.conf != nullptr;
    for (unsigned i = threadIdx.x; i < 3u * rows * wpx; i += 256) {
      const unsigned c = i / (rows * wpx), q = i - c * rows * wpx;
      const unsigned pix = (unsigned)(y_begin + q / wpx) * W + (unsigned)tx * TW + q % wpx;
      float ov = 0.f;
      if (blend) {
        const float wc = __ldg(a.conf + (size_t)b * HW + pix);
        ov = __ldg(a.fake + ((size_t)b * 3 + c) * HW + pix) * wc + ov * (1.0f - wc);  // src/flow_net.py:98
      }
      st_stream_f32(a.out_rgb + ((size_t)b * 3 + c) * HW + pix, ov);
    }
  }
}

// Two phases per CTA tile (a strip of TW = 8 * 32/LPP pixel columns x rows_per_cta rows):
//
// Phase A, features.  A group of LPP = C/8 lanes owns one pixel column; lane j owns channels 8j..8j+7 and
//   PREPARES reference kk = j % KT: flow sample, clamped corner, bilinear weights pre-multiplied by
//   softmax * visibility.  Offsets and weights travel by warp shuffle; each lane gathers its 16 bytes of
//   every tap with 128-bit loads (a warp-wide load = whole 128-byte lines) and reduces over K in registers
//   with packed fp32x2 FMAs.  The CTA sweeps downwards, so the lower tap row of one step is the upper
//   row of the next (L1 hits).
// Phase B, RGB.  The same threads re-walk the tile one thread per pixel (coalesced planar fp32 reads:
//   one or two lines per warp-wide load), re-reading the tile's flow / logit lines from L2.  Softmax in
//   the reference's sequential order, ATen's accumulation order: bit-identical to the generic kernel.
//
// POSES (pose-driven flavour, SURVEY §7 step 4: "the [B,H,W,2] grid never round-trips HBM").  A phase 0 precedes the two
// phases: one thread per tile pixel reads the target pose's z-buffer key, recomputes the winning face's barycentric
// weights with the rasteriser's pinned arithmetic (rows a5/a6) and composes the transfer flow into every reference
// pose (row a9, src/nmr.py:617-659 with src/cal_flow.py:30-31) — the work of k_raster_resolve<.,COMPOSE>, but the K flows
// of the tile land in SHARED memory, where phases A and B read them.  The flow tensor T [B,K,H,W,2] and the face-index
// map are written to HBM only when the caller passes pointers (bit-identical to jaf_cal_flow_multi's).  A tile without
// a single covered pixel (80 % of the tiles of a DanceVideo frame) is finished by a vectorised zero fill.
template <int LPP, int KT, int MINB, bool SKIP, int ROWS_REQ = 1, bool POSES = false>
__global__ void __launch_bounds__(256, MINB)
k_warp_fuse_nhwc(const WFArgs a) {
  static_assert(KT <= LPP, "one lane of the pixel group per reference");
  static_assert(!POSES || SKIP, "the pose-driven flavour carries pixel-level visibility");
  // ROWS = 2: the group's spare lanes prepare the NEXT row as well, so one pass of sample-position /
  // softmax arithmetic serves two output rows
  constexpr int ROWS = (ROWS_REQ >= 2 && LPP >= 2 * KT) ? 2 : 1;
  constexpr int PPW = 32 / LPP;  // pixel columns per warp
  constexpr int TW = 8 * PPW;    // strip width of the CTA (8 warps)
  constexpr bool KPOW2 = (KT & (KT - 1)) == 0;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned PIXB = LPP * 16;  // bytes of one channels-last pixel
  int bid = blockIdx.x;
  const int tx = bid % a.tiles_x;
  bid /= a.tiles_x;
  const int ty = bid % a.tiles_y;
  const int b = bid / a.tiles_y;
  const int y_begin = ty * a.rows_per_cta;
  const int y_end = min(a.H, y_begin + a.rows_per_cta);
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  // CTA-uniform 64-bit bases + 32-bit per-lane offsets: one IMAD.WIDE per address
  const float* __restrict__ b_logit = a.logits ? a.logits + bK : nullptr;
  const float* __restrict__ b_vis = a.vis ? a.vis + bK : nullptr;
  const int* __restrict__ b_fim = (!a.vis && a.fim) ? a.fim + (size_t)b * HW : nullptr;
  const float2* __restrict__ b_grid = POSES ? nullptr : reinterpret_cast<const float2*>(a.grid) + bK;
  const float* __restrict__ b_mask = a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr;

  // =========================== phase 0 (POSES): flows of the tile from the poses ===========================
  extern __shared__ __align__(16) unsigned char wf_smem[];
  const unsigned tile_px = (unsigned)TW * (unsigned)a.rows_per_cta;
  // dynamic shared memory of the SKIP flavours: [K flows per tile pixel (POSES only)] [coverage flag per pixel]
  // [compacted list of covered pixels: (tile-local index, face)]
  constexpr size_t kFlowBytes = POSES ? (size_t)KT * sizeof(float2) : 0;
  float2* __restrict__ s_T = reinterpret_cast<float2*>(wf_smem);                       // [KT][tile_px]
  unsigned char* __restrict__ s_vis = wf_smem + kFlowBytes * tile_px;                  // [tile_px]
  int2* __restrict__ s_list = reinterpret_cast<int2*>(wf_smem + ((kFlowBytes * tile_px + tile_px + 15) & ~(size_t)15));
  __shared__ unsigned s_cnt;
  if constexpr (POSES) {
    using namespace jaf_raster;
    const int S = a.H;  // the target raster IS the output frame (H == W == raster size, checked by the host)
    // [p, face] of the covered pixels of the tile are compacted into s_list: the expensive part below runs on full warps
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    // ---- step 1: z-buffer keys -> coverage.  (fim is written here: it needs nothing else)
    int any_vis = 0;
    for (unsigned p = threadIdx.x; p < tile_px; p += 256) {
      const int x = tx * TW + (int)(p % TW), y = y_begin + (int)(p / TW);
      int fn = -1;
      if (x < (int)W && y < y_end) {
        const int yi = S - 1 - y;  // NR/rasterize.py:334-338: output row y is raster row S-1-y
        unsigned long long* kp = a.zkeys + ((size_t)b * S + yi) * S + x;
        const unsigned long long key = *kp;
        if (key != kEmptyKey) {
          fn = (int)(unsigned int)(key & 0xffffffffull);
          s_list[atomicAdd(&s_cnt, 1u)] = make_int2((int)p, fn);
          if (a.leave_clean) *kp = kEmptyKey;  // only covered pixels (~12 %) need the write
        }
        if (a.fim_out != nullptr) a.fim_out[(size_t)b * HW + (unsigned)y * W + (unsigned)x] = fn;
      }
      s_vis[p] = fn >= 0 ? 1 : 0;
      any_vis |= fn >= 0;
    }
    if (__syncthreads_or(any_vis) == 0) {  // nothing of the body in this tile
      if (a.T_out != nullptr) {            // the optional flow output still gets its sentinel (src/nmr.py:627)
        for (unsigned p = threadIdx.x; p < tile_px; p += 256) {
          const int x = tx * TW + (int)(p % TW), y = y_begin + (int)(p / TW);
          if (x < (int)W && y < y_end) {
#pragma unroll
            for (int ks = 0; ks < KT; ++ks)
              reinterpret_cast<float2*>(a.T_out)[((size_t)b * KT + ks) * HW + (unsigned)y * W + (unsigned)x] = make_float2(-2.0f, -2.0f);
          }
        }
      }
      fill_empty_tile<LPP, TW>(a, b, tx, y_begin, y_end, W, HW);
      return;
    }
    // ---- step 2: uncovered pixels of a tile that does hold body pixels: the sentinel flow
    for (unsigned p = threadIdx.x; p < tile_px; p += 256) {
      if (s_vis[p]) continue;
      const int x = tx * TW + (int)(p % TW), y = y_begin + (int)(p / TW);
      const bool inside = x < (int)W && y < y_end;
#pragma unroll
      for (int ks = 0; ks < KT; ++ks) {
        s_T[(unsigned)ks * tile_px + p] = make_float2(-2.0f, -2.0f);
        if (inside && a.T_out != nullptr)
          reinterpret_cast<float2*>(a.T_out)[((size_t)b * KT + ks) * HW + (unsigned)y * W + (unsigned)x] = make_float2(-2.0f, -2.0f);
      }
    }
    // ---- step 3: covered pixels: barycentric weights of the winning face (rows a5/a6), K composed flows (row a9)
    const unsigned cnt = s_cnt;
    for (unsigned i = threadIdx.x; i < cnt; i += 256) {
      const int2 e = s_list[i];
      const unsigned p = (unsigned)e.x;
      const int fn = e.y;
      const int x = tx * TW + (int)(p % TW), y = y_begin + (int)(p / TW);
      const int yi = S - 1 - y;
      float f[9], inv[9], px[3], py[3], zp, w[3];
      load_face_projected(a.tgt_cam, a.tgt_verts, a.fidx, b, fn, a.V, a.eye_z, f);
      face_setup(f, S, inv, px, py);
      pixel_test(f, inv, x, yi, S, a.near_, a.far_, w, &zp);
      int vi[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) vi[k] = __ldg(a.fidx + fn * 3 + k);
#pragma unroll
      for (int ks = 0; ks < KT; ++ks) {
        const size_t sb = r * KT + ks;
        const float sc = __ldg(a.src_cam + sb * 3 + 0), ctx = __ldg(a.src_cam + sb * 3 + 1), cty = __ldg(a.src_cam + sb * 3 + 2);
        float ax[3], ay[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float* pv = a.src_verts + (sb * a.V + vi[k]) * 3;
          ax[k] = __fmul_rn(sc, __fadd_rn(__ldg(pv), ctx));
          ay[k] = __fmul_rn(sc, __fadd_rn(__ldg(pv + 1), cty));  // -(-(s*(Y+ty))): raster flip then cal_flow.py:31
        }
        const float ftx = __fadd_rn(__fadd_rn(__fmul_rn(ax[0], w[0]), __fmul_rn(ax[1], w[1])), __fmul_rn(ax[2], w[2]));
        const float fty = __fadd_rn(__fadd_rn(__fmul_rn(ay[0], w[0]), __fmul_rn(ay[1], w[1])), __fmul_rn(ay[2], w[2]));
        s_T[(unsigned)ks * tile_px + p] = make_float2(ftx, fty);
        if (a.T_out != nullptr)
          reinterpret_cast<float2*>(a.T_out)[((size_t)b * KT + ks) * HW + (unsigned)y * W + (unsigned)x] = make_float2(ftx, fty);
      }
    }
    __syncthreads();
  } else if constexpr (SKIP) {
    // pixel-level visibility from a face-index map: the same whole-tile early-out (80 % of the tiles of a DanceVideo
    // frame hold no body pixel); the map's lines are read again, from L2, by the row loop of the other tiles
    if (b_fim != nullptr) {  // uniform
      if (threadIdx.x == 0) s_cnt = 0;
      __syncthreads();
      int any_vis = 0;
      for (unsigned p = threadIdx.x; p < tile_px; p += 256) {
        const int x = tx * TW + (int)(p % TW), y = y_begin + (int)(p / TW);
        int fn = -1;
        if (x < (int)W && y < y_end) fn = __ldg(b_fim + (unsigned)y * W + (unsigned)x);
        if (fn != -1) s_list[atomicAdd(&s_cnt, 1u)] = make_int2((int)p, fn);
        s_vis[p] = fn != -1 ? 1 : 0;
        any_vis |= fn != -1;
      }
      if (__syncthreads_or(any_vis) == 0) {
        fill_empty_tile<LPP, TW>(a, b, tx, y_begin, y_end, W, HW);
        return;
      }
    }
  }

  // =========================== list-driven phases (pixel-level visibility) ===========================
  // With pixel-level visibility (pose-driven, or a face-index map) the tile's covered pixels are known as a compacted list:
  // the feature and RGB phases walk that list instead of the rows, so an uncovered pixel inside a partly covered tile costs
  // one zero store instead of K x 4 weight-0 gathers and their FMAs (the body covers 40-50 % of the tiles it touches).
  // Per pixel the operations and their order are those of the row loop below: identical results.
  if constexpr (SKIP) {
    if (POSES || b_fim != nullptr) {  // uniform
      const unsigned cnt = s_cnt;
      constexpr int SL = (LPP >= 2 * KT) ? 2 : 1;        // list entries a lane group prepares per step (spare lanes: the 2nd)
      constexpr int PER_WARP = (32 / LPP) * SL;
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int g = lane / LPP, j = lane % LPP, gl = g * LPP;
      const int kk = j % KT;
      const int slot = (SL == 2) ? min(j / KT, 1) : 0;
      const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + j * 16;
      uint4* __restrict__ o_lane = reinterpret_cast<uint4*>(a.out_feat) + (size_t)b * HW * LPP + j;
      const uint64_t keep = l2_policy_evict_last();
      const bool have_feat = a.feat != nullptr && a.out_feat != nullptr;  // RGB-only calls skip the feature part (uniform)
#pragma unroll 1
      for (unsigned base = (unsigned)warp * PER_WARP; have_feat && base < cnt; base += 8u * PER_WARP) {  // warp-uniform
        const unsigned i = base + (unsigned)(g * SL + slot);
        const bool pin = i < cnt;
        const unsigned p = pin ? (unsigned)s_list[i].x : 0u;
        const unsigned pix = (unsigned)(y_begin + (int)(p / TW)) * W + (unsigned)(tx * TW) + p % TW;
        float lg = 0.f, v = 1.f;
        float2 gxy = make_float2(0.f, 0.f);
        if (pin) {
          if constexpr (POSES) gxy = s_T[(unsigned)kk * tile_px + p];
          else gxy = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + ((unsigned)kk * HW + pix)), keep);
          if (b_logit) lg = ld_stream_keep_f32(b_logit + ((unsigned)kk * HW + pix), keep);
          if (b_mask) v *= ld_stream_f32(b_mask + pix);  // fused*mask == sum_k (alpha_k vis_k mask) warped_k
        }
        const int kb = gl + slot * KT;
        float m = lg, ssum;
        if constexpr (KPOW2) {
#pragma unroll
          for (int sft = KT / 2; sft > 0; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, sft));
        } else {
#pragma unroll
          for (int k = 0; k < KT; ++k) m = fmaxf(m, __shfl_sync(FULL, lg, (kb + k) & 31));
        }
        const float e = expf(lg - m);
        if constexpr (KPOW2) {
          ssum = e;
#pragma unroll
          for (int sft = 1; sft < KT; sft <<= 1) ssum += __shfl_xor_sync(FULL, ssum, sft);
        } else {
          ssum = 0.f;
#pragma unroll
          for (int k = 0; k < KT; ++k) ssum += __shfl_sync(FULL, e, (kb + k) & 31);
        }
        const float aw = pin ? __fdividef(e, ssum) * v : 0.f;  // alpha_k * vis_k * mask
        HotTap t = make_hot_tap(gxy.x, gxy.y, (int)Ws, a.Hs, a.align_corners);
        t.nw *= aw;
        t.ne *= aw;
        t.sw *= aw;
        t.se *= aw;
        const unsigned off = (aw != 0.f) ? (unsigned)t.off : 0u;
#pragma unroll
        for (int rs = 0; rs < SL; ++rs) {
          const unsigned i_rs = base + (unsigned)(g * SL + rs);
          const unsigned pix_rs = __shfl_sync(FULL, pix, gl + rs * KT);
          float2 acc[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[c] = make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const int src = gl + rs * KT + k;
            const unsigned o0 = __shfl_sync(FULL, off, src) + (unsigned)k * HWs;
            const uint4* p0 = reinterpret_cast<const uint4*>(f_lane + (size_t)o0 * PIXB);
            const uint4* p1 = reinterpret_cast<const uint4*>(f_lane + (size_t)(o0 + Ws) * PIXB);
            uint4 q[4];
            q[0] = ld_gather_u128(p0);
            q[1] = ld_gather_u128(p0 + LPP);
            q[2] = ld_gather_u128(p1);
            q[3] = ld_gather_u128(p1 + LPP);
            float wt[4];
            wt[0] = __shfl_sync(FULL, t.nw, src);
            wt[1] = __shfl_sync(FULL, t.ne, src);
            wt[2] = __shfl_sync(FULL, t.sw, src);
            wt[3] = __shfl_sync(FULL, t.se, src);
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
              const float2 w2 = make_float2(wt[tp], wt[tp]);
              const uint32_t wd[4] = {q[tp].x, q[tp].y, q[tp].z, q[tp].w};
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acc[c]);
            }
          }
          if (i_rs < cnt) {
            uint4 o;
            o.x = pack_bf16x2(acc[0].x, acc[0].y);
            o.y = pack_bf16x2(acc[1].x, acc[1].y);
            o.z = pack_bf16x2(acc[2].x, acc[2].y);
            o.w = pack_bf16x2(acc[3].x, acc[3].y);
            st_stream_u128(o_lane + (size_t)pix_rs * LPP, o);
          }
        }
      }
      // uncovered pixels of the tile: the empty feature vector
      if (have_feat) {
        uint4* __restrict__ o_b = reinterpret_cast<uint4*>(a.out_feat) + (size_t)b * HW * LPP;
        for (unsigned q = threadIdx.x; q < tile_px * LPP; q += 256) {
          const unsigned p = q / LPP, l = q % LPP;
          const int x = tx * TW + (int)(p % TW), y = y_begin + (int)(p / TW);
          if (x < (int)W && y < y_end && !s_vis[p])
            st_stream_u128(o_b + ((size_t)y * W + (unsigned)x) * LPP + l, make_uint4(0u, 0u, 0u, 0u));
        }
      }
      // RGB planes: covered pixels from the list, the others empty (or the blend with the empty frame)
      if (a.rgb != nullptr && a.out_rgb != nullptr) {
        const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
        const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
        const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
        float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
        for (unsigned i = threadIdx.x; i < cnt; i += 256) {
          const unsigned p = (unsigned)s_list[i].x;
          const unsigned pix = (unsigned)(y_begin + (int)(p / TW)) * W + (unsigned)(tx * TW) + p % TW;
          if constexpr (POSES)
            rgb_pixel<KT, SKIP>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, pix, HW, HWs, Ws, s_T,
                                tile_px, p, 1);
          else
            rgb_pixel<KT, SKIP>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, pix, HW, HWs, Ws);
        }
        for (unsigned p = threadIdx.x; p < tile_px; p += 256) {
          const int x = tx * TW + (int)(p % TW), y = y_begin + (int)(p / TW);
          if (x >= (int)W || y >= y_end || s_vis[p]) continue;
          const unsigned pix = (unsigned)y * W + (unsigned)x;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float ov = 0.f;
            if (b_mask) ov *= (a.mask_c == 3) ? __ldg(b_mask + ((unsigned)c * HW + pix)) : __ldg(b_mask + pix);
            if (b_fake) {
              const float wc = __ldg(b_conf + pix);
              ov = __ldg(b_fake + ((unsigned)c * HW + pix)) * wc + ov * (1.0f - wc);  // src/flow_net.py:98
            }
            st_stream_f32(b_orgb + ((unsigned)c * HW + pix), ov);
          }
        }
      }
      return;
    }
  }

  // =========================== phase A: features ===========================
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPP, j = lane % LPP, gl = g * LPP;
    const int kk = j % KT;
    const int rr = (ROWS == 2) ? min(j / KT, 1) : 0;  // row slot this lane prepares (surplus lanes replicate)
    const int x = tx * TW + warp * PPW + g;
    const bool xin = x < (int)W;
    const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + j * 16;
    uint4* __restrict__ o_lane = reinterpret_cast<uint4*>(a.out_feat) + (size_t)b * HW * LPP + j;
    const unsigned lane_in = (unsigned)kk * HW + (unsigned)rr * W;  // this lane's (b, kk) plane, row slot rr

    // flow / logit lines are read again by phase B ~100 us later: ask L2 to keep them (evict_last)
    const uint64_t keep = l2_policy_evict_last();
    unsigned pix = (unsigned)y_begin * W + (unsigned)x;
#pragma unroll 1
    for (int y = y_begin; y < y_end; y += ROWS, pix += ROWS * W) {
      // ---- prepare: (row slot rr, reference kk)
      const bool pin = xin && (y + rr < y_end);
      float lg = 0.f, v = 1.f;
      float2 gxy = make_float2(0.f, 0.f);
      // tile-local pixel index of (row slot rr, this lane's column): where phase 0 left the flows and visibility
      const unsigned lp = (unsigned)(y + rr - y_begin) * TW + (unsigned)(warp * PPW + g);
      if constexpr (SKIP) {
        if (POSES || b_fim != nullptr) {  // uniform.  Pixel-level visibility: look at the face-index map first
          if (pin) v = POSES ? (float)s_vis[lp] : ((ld_stream_s32(b_fim + (pix + (unsigned)rr * W)) != -1) ? 1.f : 0.f);
          if (__ballot_sync(FULL, pin && v != 0.f) == 0u) {
            // every pixel this warp owns in these rows is background: empty output, no flow / logit / reference read
            if (xin) {
#pragma unroll
              for (int rs = 0; rs < ROWS; ++rs)
                if (y + rs < y_end) st_stream_u128(o_lane + (size_t)(pix + (unsigned)rs * W) * LPP, make_uint4(0u, 0u, 0u, 0u));
            }
            continue;
          }
        }
      }
      if (pin) {
        if constexpr (POSES) gxy = s_T[(unsigned)kk * tile_px + lp];
        else gxy = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + (lane_in + pix)), keep);
        if (b_logit) lg = ld_stream_keep_f32(b_logit + (lane_in + pix), keep);
        if (b_vis) v = ld_stream_f32(b_vis + (lane_in + pix));
        if (!SKIP && b_fim) v = (ld_stream_s32(b_fim + (pix + (unsigned)rr * W)) != -1) ? 1.f : 0.f;
        if (b_mask) v *= ld_stream_f32(b_mask + (pix + (unsigned)rr * W));  // fused*mask == sum_k (alpha_k vis_k mask) warped_k
      }
      // softmax over the KT lanes of this row slot (they are an aligned block when KT is a power of two)
      const int kb = gl + rr * KT;
      float m = lg, ssum;
      if constexpr (KPOW2) {
#pragma unroll
        for (int s = KT / 2; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
      } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) m = fmaxf(m, __shfl_sync(FULL, lg, (kb + k) & 31));
      }
      const float e = expf(lg - m);
      if constexpr (KPOW2) {
        ssum = e;
#pragma unroll
        for (int s = 1; s < KT; s <<= 1) ssum += __shfl_xor_sync(FULL, ssum, s);
      } else {
        ssum = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) ssum += __shfl_sync(FULL, e, (kb + k) & 31);
      }
      const float aw = pin ? __fdividef(e, ssum) * v : 0.f;  // alpha_k * vis_k * mask
      HotTap t = make_hot_tap(gxy.x, gxy.y, (int)Ws, a.Hs, a.align_corners);
      t.nw *= aw;
      t.ne *= aw;
      t.sw *= aw;
      t.se *= aw;
      // invisible references still issue their (weight-0) loads, from pixel 0 of the reference: no
      // divergent branch in the gather loop and no new cache lines
      const unsigned off = (aw != 0.f) ? (unsigned)t.off : 0u;
      // SKIP (host-selected when a visibility input exists): warp-uniform early-out when nothing is visible
      const bool any = SKIP ? (__ballot_sync(FULL, aw != 0.f) != 0u) : true;

#pragma unroll
      for (int rs = 0; rs < ROWS; ++rs) {
        if (ROWS == 2 && rs == 1 && y + 1 >= y_end) break;  // uniform
        float2 acc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] = make_float2(0.f, 0.f);
        if (any) {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const int src = gl + rs * KT + k;
            const unsigned o0 = __shfl_sync(FULL, off, src) + (unsigned)k * HWs;
            const uint4* p0 = reinterpret_cast<const uint4*>(f_lane + (size_t)o0 * PIXB);
            const uint4* p1 = reinterpret_cast<const uint4*>(f_lane + (size_t)(o0 + Ws) * PIXB);
            uint4 q[4];
            q[0] = ld_gather_u128(p0);
            q[1] = ld_gather_u128(p0 + LPP);
            q[2] = ld_gather_u128(p1);
            q[3] = ld_gather_u128(p1 + LPP);
            float wt[4];
            wt[0] = __shfl_sync(FULL, t.nw, src);
            wt[1] = __shfl_sync(FULL, t.ne, src);
            wt[2] = __shfl_sync(FULL, t.sw, src);
            wt[3] = __shfl_sync(FULL, t.se, src);
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
              const float2 w2 = make_float2(wt[tp], wt[tp]);
              const uint32_t wd[4] = {q[tp].x, q[tp].y, q[tp].z, q[tp].w};
#pragma unroll
              for (int c = 0; c < 4; ++c)  // packed fp32x2 FMA: two channels per instruction
                acc[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acc[c]);
            }
          }
        }
        if (xin) {
          uint4 o;
          o.x = pack_bf16x2(acc[0].x, acc[0].y);
          o.y = pack_bf16x2(acc[1].x, acc[1].y);
          o.z = pack_bf16x2(acc[2].x, acc[2].y);
          o.w = pack_bf16x2(acc[3].x, acc[3].y);
          st_stream_u128(o_lane + (size_t)(pix + (unsigned)rs * W) * LPP, o);
        }
      }
    }
  }

  // =========================== phase B: RGB ===========================
  if (a.rgb != nullptr && a.out_rgb != nullptr) {
    const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
    const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
    const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
    float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
    const int npx = TW * (y_end - y_begin);
    for (int p = threadIdx.x; p < npx; p += 256) {
      const int x = tx * TW + p % TW, y = y_begin + p / TW;
      if constexpr (POSES) {
        if (x >= (int)W) continue;
        rgb_pixel<KT, SKIP>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws,
                            s_T, tile_px, (unsigned)p, (int)s_vis[p]);
      } else if constexpr (SKIP) {
        if (x >= (int)W) continue;
        rgb_pixel<KT, SKIP>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws);
      } else {  // TW >= 32: a warp is one row segment and stays converged (tap sharing by shuffle)
        const int xc = min(x, (int)W - 1);
        // (narrower strips end the loop with partial warps: no sharing there)
        rgb_pixel<KT, SKIP, (KT <= 4), (kShareRgbTaps && TW % 32 == 0)>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb,
                                                       (unsigned)y * W + (unsigned)xc, HW, HWs, Ws, nullptr, 0, 0, 1, x < (int)W);
      }
    }
  }
}

// =====================================================================================
// Wide-lane flavour of the hot kernel for C = 64: a group of FOUR lanes owns a pixel column, each lane 16 channels
// (32 bytes) per tap, moved with one 256-bit load (LDG.E.256, new on sm_100).  A warp covers 8 columns, the CTA 64.
// Same arithmetic, order and results as k_warp_fuse_nhwc<8, K>; what changes is the amount of data behind each load
// instruction: twice the bytes per scoreboard slot (a warp can only track a handful of outstanding loads) and half the
// per-lane overhead (shuffles, addresses, sample-position arithmetic) per byte.
// =====================================================================================
struct U256 {
  uint4 lo, hi;
};
__device__ __forceinline__ U256 ld_gather_u256(const void* p) {
  U256 v;
  asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v.lo.x), "=r"(v.lo.y), "=r"(v.lo.z), "=r"(v.lo.w), "=r"(v.hi.x), "=r"(v.hi.y), "=r"(v.hi.z), "=r"(v.hi.w)
      : "l"(p));
  return v;
}

// The same gather issued only by the lanes whose `take` is set (the others get unspecified registers).
__device__ __forceinline__ U256 ld_gather_u256_if(const void* p, bool take) {
  U256 v;
  asm("{\n\t.reg .pred pp;\n\tsetp.ne.u32 pp, %9, 0;\n\t@pp ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t}"
      : "=r"(v.lo.x), "=r"(v.lo.y), "=r"(v.lo.z), "=r"(v.lo.w), "=r"(v.hi.x), "=r"(v.hi.y), "=r"(v.hi.z), "=r"(v.hi.w)
      : "l"(p), "r"((uint32_t)take));
  return v;
}
// coh ? (the same word of the pixel group one to the right: lane + 4) : own
__device__ __forceinline__ uint32_t right_word(uint32_t left, uint32_t own, bool coh) {
  const uint32_t t = __shfl_down_sync(0xffffffffu, left, 4);
  return coh ? t : own;
}

// RGBM ("RGB merged"): the RGB planes are produced inside the feature row loop instead of a second pass over the
// tile.  In the reduction over the references every lane of the pixel group already receives reference k's corner
// offset and its four bilinear weights (pre-multiplied by softmax * visibility * mask) by shuffle; lane c < 3 of the
// group uses them once more for plane c of the planar fp32 RGB reference: 4 scalar taps per reference, issued with the
// same batch of loads as the feature taps, one running FMA chain in ATen's tap order, one plane store per row.
// Nothing is re-read (the two-pass flavour reads flows, logits and masks twice and rebuilds softmax and taps per
// pixel), no CTA barrier, one extra accumulator register.  The weights carry the approximate softmax division and
// the mask before the sum instead of after it: <= 1e-5 of the generic kernel instead of bit-identical (the north
// star's fp32 tolerance is 1e-4).  Per-channel target masks (mask_c == 3) keep the two-pass flavour.
template <int KT, int MINB, bool SKIP, bool RGBM>
__global__ void __launch_bounds__(256, MINB)
k_warp_fuse_nhwc_wide(const WFArgs a) {
  static_assert(KT <= 4, "one lane of the 4-lane pixel group per reference");
  constexpr int LPP = 4, PPW = 8, TW = 64;
  constexpr bool KPOW2 = (KT & (KT - 1)) == 0;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned PIXB = 128;  // bytes of one channels-last pixel (64 bf16)
  int bid = blockIdx.x;
  const int tx = bid % a.tiles_x;
  bid /= a.tiles_x;
  const int ty = bid % a.tiles_y;
  const int b = bid / a.tiles_y;
  const int y_begin = ty * a.rows_per_cta;
  const int y_end = min(a.H, y_begin + a.rows_per_cta);
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  const float* __restrict__ b_logit = a.logits ? a.logits + bK : nullptr;
  const float* __restrict__ b_vis = a.vis ? a.vis + bK : nullptr;
  const int* __restrict__ b_fim = (!a.vis && a.fim) ? a.fim + (size_t)b * HW : nullptr;
  const float2* __restrict__ b_grid = reinterpret_cast<const float2*>(a.grid) + bK;
  const float* __restrict__ b_mask = a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr;

  // =========================== phase A: features ===========================
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPP, j = lane % LPP, gl = g * LPP;
    const int kk = j % KT;
    const int x = tx * TW + warp * PPW + g;
    const bool xin = x < (int)W;
    const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + j * 32;
    char* __restrict__ o_lane = reinterpret_cast<char*>(a.out_feat) + (size_t)b * HW * PIXB + j * 32;
    const unsigned lane_in = (unsigned)kk * HW;
    const uint64_t keep = l2_policy_evict_last();
    unsigned pix = (unsigned)y_begin * W + (unsigned)x;
    // merged RGB: CTA-uniform 64-bit base + 32-bit per-lane offsets (lane 3 shadows plane 0 and stores nothing)
    const float* __restrict__ rgb_cta = RGBM ? a.rgb + r * KT * 3 * (size_t)HWs : nullptr;
    const unsigned rgb_c = (j < 3 ? (unsigned)j : 0u) * HWs;
#if defined(JAF_PROBE_TAPS) && JAF_PROBE_TAPS == 5
    uint32_t probe_sink = 0;
#endif
#pragma unroll 1
    for (int y = y_begin; y < y_end; ++y, pix += W) {
      float lg = 0.f, v = 1.f;
      float2 gxy = make_float2(0.f, 0.f);
      if (xin) {
        gxy = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + (lane_in + pix)), keep);
        if (b_logit) lg = ld_stream_keep_f32(b_logit + (lane_in + pix), keep);
        if (b_vis) v = ld_stream_f32(b_vis + (lane_in + pix));
        if (b_fim) v = (ld_stream_s32(b_fim + pix) != -1) ? 1.f : 0.f;
        if (b_mask) v *= ld_stream_f32(b_mask + pix);
      }
      float m = lg, ssum;
      if constexpr (KPOW2) {
#pragma unroll
        for (int s = KT / 2; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
      } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) m = fmaxf(m, __shfl_sync(FULL, lg, gl + k));
      }
      const float e = expf(lg - m);
      if constexpr (KPOW2) {
        ssum = e;
#pragma unroll
        for (int s = 1; s < KT; s <<= 1) ssum += __shfl_xor_sync(FULL, ssum, s);
      } else {
        ssum = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) ssum += __shfl_sync(FULL, e, gl + k);
      }
      const float aw = xin ? __fdividef(e, ssum) * v : 0.f;
      HotTap t = make_hot_tap(gxy.x, gxy.y, (int)Ws, a.Hs, a.align_corners);
      t.nw *= aw;
      t.ne *= aw;
      t.sw *= aw;
      t.se *= aw;
      const unsigned off = (aw != 0.f) ? (unsigned)t.off : 0u;
      const bool any = SKIP ? (__ballot_sync(FULL, aw != 0.f) != 0u) : true;

      float2 acc[8];
      float racc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
      if (any) {
#if defined(JAF_WF_XTAP_UNROLL1)
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int k = 0; k < KT; ++k) {
          const int src = gl + k;
          const unsigned osh = __shfl_sync(FULL, off, src);
          const unsigned o0 = osh + (unsigned)k * HWs;
          const char* p0 = f_lane + (size_t)o0 * PIXB;
          const char* p1 = f_lane + (size_t)(o0 + Ws) * PIXB;
          U256 q[4];
          q[0] = ld_gather_u256(p0);
          q[2] = ld_gather_u256(p1);
#if defined(JAF_PROBE_TAPS) && JAF_PROBE_TAPS == 2   // what-if probes (WRONG results): upper bound of any tap-sharing scheme
          q[1] = q[0];
          q[3] = q[2];
#elif defined(JAF_PROBE_TAPS) && JAF_PROBE_TAPS == 3
          q[1] = ld_gather_u256(p0 + PIXB);
          q[3] = q[2];
#elif defined(JAF_PROBE_TAPS) && JAF_PROBE_TAPS == 4   // the se gather dropped, its unpack kept (runtime-zero xor defeats CSE)
          q[1] = ld_gather_u256(p0 + PIXB);
          {
            const uint32_t z = blockIdx.x >> 31;
            q[3].lo = make_uint4(q[2].lo.x ^ z, q[2].lo.y ^ z, q[2].lo.z ^ z, q[2].lo.w ^ z);
            q[3].hi = make_uint4(q[2].hi.x ^ z, q[2].hi.y ^ z, q[2].hi.z ^ z, q[2].hi.w ^ z);
          }
#elif defined(JAF_WF_XTAP)
          // the ne / se taps of this pixel are the nw / sw taps of the pixel group to the right whenever that group's
          // corner is one column further: take them from its registers (8 shuffles per tap) instead of gathering the
          // same line a second time; groups whose neighbour samples elsewhere (and the last group of the warp) gather
          const unsigned osh_r = __shfl_down_sync(FULL, osh, 4);
          const bool coh = lane < 28 && osh_r == osh + 1u;
          q[1] = ld_gather_u256_if(p0 + PIXB, !coh);
          q[3] = ld_gather_u256_if(p1 + PIXB, !coh);
#elif defined(JAF_PROBE_TAPS) && JAF_PROBE_TAPS == 5   // the se gather kept (folded into a sink), its unpack shared with sw
          q[1] = ld_gather_u256(p0 + PIXB);
          {
            const U256 t3 = ld_gather_u256(p1 + PIXB);
            probe_sink ^= t3.lo.x ^ t3.lo.y ^ t3.lo.z ^ t3.lo.w ^ t3.hi.x ^ t3.hi.y ^ t3.hi.z ^ t3.hi.w;
            q[3] = q[2];
          }
#else
          q[1] = ld_gather_u256(p0 + PIXB);
          q[3] = ld_gather_u256(p1 + PIXB);
#endif
          float rv[4];
          if constexpr (RGBM) {
            const float* r0 = rgb_cta + (osh + (unsigned)(3 * k) * HWs + rgb_c);
            rv[0] = __ldg(r0);
            rv[1] = __ldg(r0 + 1);
            rv[2] = __ldg(r0 + Ws);
            rv[3] = __ldg(r0 + Ws + 1);
          }
          float wt[4];
          wt[0] = __shfl_sync(FULL, t.nw, src);
          wt[1] = __shfl_sync(FULL, t.ne, src);
          wt[2] = __shfl_sync(FULL, t.sw, src);
          wt[3] = __shfl_sync(FULL, t.se, src);
          if constexpr (RGBM) {
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) racc = fmaf(rv[tp], wt[tp], racc);
          }
#pragma unroll
          for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
            const float2 w2 = make_float2(wt[tp], wt[tp]);
#if defined(JAF_WF_XTAP)
            // odd taps (ne, se): the words come from the right-hand group's even tap when its corner is adjacent
            const int te = tp & ~1;
            const uint32_t wd[8] = {
                (tp & 1) ? right_word(q[te].lo.x, q[tp].lo.x, coh) : q[tp].lo.x, (tp & 1) ? right_word(q[te].lo.y, q[tp].lo.y, coh) : q[tp].lo.y,
                (tp & 1) ? right_word(q[te].lo.z, q[tp].lo.z, coh) : q[tp].lo.z, (tp & 1) ? right_word(q[te].lo.w, q[tp].lo.w, coh) : q[tp].lo.w,
                (tp & 1) ? right_word(q[te].hi.x, q[tp].hi.x, coh) : q[tp].hi.x, (tp & 1) ? right_word(q[te].hi.y, q[tp].hi.y, coh) : q[tp].hi.y,
                (tp & 1) ? right_word(q[te].hi.z, q[tp].hi.z, coh) : q[tp].hi.z, (tp & 1) ? right_word(q[te].hi.w, q[tp].hi.w, coh) : q[tp].hi.w};
#else
            const uint32_t wd[8] = {q[tp].lo.x, q[tp].lo.y, q[tp].lo.z, q[tp].lo.w,
                                    q[tp].hi.x, q[tp].hi.y, q[tp].hi.z, q[tp].hi.w};
#endif
#pragma unroll
            for (int c = 0; c < 8; ++c)
              acc[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acc[c]);
          }
        }
      }
      if (xin) {
        uint4 o0v, o1v;
        o0v.x = pack_bf16x2(acc[0].x, acc[0].y); o0v.y = pack_bf16x2(acc[1].x, acc[1].y);
        o0v.z = pack_bf16x2(acc[2].x, acc[2].y); o0v.w = pack_bf16x2(acc[3].x, acc[3].y);
        o1v.x = pack_bf16x2(acc[4].x, acc[4].y); o1v.y = pack_bf16x2(acc[5].x, acc[5].y);
        o1v.z = pack_bf16x2(acc[6].x, acc[6].y); o1v.w = pack_bf16x2(acc[7].x, acc[7].y);
        uint4* op = reinterpret_cast<uint4*>(o_lane + (size_t)pix * PIXB);
        st_stream_u128(op, o0v);
        st_stream_u128(op + 1, o1v);
      }
      if constexpr (RGBM) {
        if (xin && j < 3) {
          const unsigned po = (unsigned)j * HW + pix;
          float ov = racc;
          if (a.fake != nullptr && a.conf != nullptr) {  // uniform
            const float wc = __ldg(a.conf + ((size_t)b * HW + pix));
            const float fkv = __ldg(a.fake + ((size_t)b * 3 * HW + po));
            ov = fkv * wc + ov * (1.0f - wc);  // src/flow_net.py:98
          }
          st_stream_f32(a.out_rgb + ((size_t)b * 3 * HW + po), ov);
        }
      }
    }
#if defined(JAF_PROBE_TAPS) && JAF_PROBE_TAPS == 5
    if (probe_sink == 0x9e3779b9u && a.out_rgb != nullptr) a.out_rgb[0] = 0.f;  // keeps the probe's loads alive
#endif
  }

  // =========================== phase B: RGB (two-pass flavour) ===========================
  if (!RGBM && a.rgb != nullptr && a.out_rgb != nullptr) {
    const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
    const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
    const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
    float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
    const int npx = TW * (y_end - y_begin);
    for (int p = threadIdx.x; p < npx; p += 256) {
      // TW is a multiple of 32: a warp is 32 x-adjacent pixels of one row and stays converged (tap sharing by shuffle);
      // lanes beyond the right edge repeat the last column and store nothing
      const int x = tx * TW + p % TW, y = y_begin + p / TW;
      if constexpr (SKIP) {
        if (x >= (int)W) continue;
        rgb_pixel<KT, SKIP>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws);
      } else {
        const int xc = min(x, (int)W - 1);
        rgb_pixel<KT, SKIP, (KT <= 4), kShareRgbTaps>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb,
                                             (unsigned)y * W + (unsigned)xc, HW, HWs, Ws, nullptr, 0, 0, 1, x < (int)W);
      }
    }
  }
}

// The same kernel for 5..8 references: every lane of the 4-lane group prepares two references (k = j and j + 4).
// (Kept apart from the K <= 4 kernel: folding both into one template costs the headline shape 3.5 %, measured.)
template <int KT, int MINB, bool SKIP>
__global__ void __launch_bounds__(256, MINB)
k_warp_fuse_nhwc_wide2(const WFArgs a) {
  static_assert(KT <= 8, "each lane of the 4-lane pixel group prepares at most two references");
  constexpr int LPP = 4, PPW = 8, TW = 64;
  constexpr int NP = (KT + LPP - 1) / LPP;  // references prepared per lane: lane j takes k = j, j + 4
  constexpr int KL = KT < LPP ? KT : LPP;   // lanes of a group that prepare slot 0
  constexpr bool KPOW2 = (KL & (KL - 1)) == 0 && KT % KL == 0;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned PIXB = 128;  // bytes of one channels-last pixel (64 bf16)
  int bid = blockIdx.x;
  const int tx = bid % a.tiles_x;
  bid /= a.tiles_x;
  const int ty = bid % a.tiles_y;
  const int b = bid / a.tiles_y;
  const int y_begin = ty * a.rows_per_cta;
  const int y_end = min(a.H, y_begin + a.rows_per_cta);
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  const float* __restrict__ b_logit = a.logits ? a.logits + bK : nullptr;
  const float* __restrict__ b_vis = a.vis ? a.vis + bK : nullptr;
  const int* __restrict__ b_fim = (!a.vis && a.fim) ? a.fim + (size_t)b * HW : nullptr;
  const float2* __restrict__ b_grid = reinterpret_cast<const float2*>(a.grid) + bK;
  const float* __restrict__ b_mask = a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr;

  // =========================== phase A: features ===========================
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPP, j = lane % LPP, gl = g * LPP;
    const int x = tx * TW + warp * PPW + g;
    const bool xin = x < (int)W;
    const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + j * 32;
    char* __restrict__ o_lane = reinterpret_cast<char*>(a.out_feat) + (size_t)b * HW * PIXB + j * 32;
    const uint64_t keep = l2_policy_evict_last();
    unsigned pix = (unsigned)y_begin * W + (unsigned)x;
#pragma unroll 1
    for (int y = y_begin; y < y_end; ++y, pix += W) {
      // ---- prepare: slot n of lane j is reference kk = j % KL + n * LPP (surplus lanes / slots replicate or idle)
      float lg[NP], vv[NP];
      float2 gxy[NP];
      bool act[NP];
      float vm = 1.f;
      if (xin) {
        if (b_fim) vm = (ld_stream_s32(b_fim + pix) != -1) ? 1.f : 0.f;
        if (b_mask) vm *= ld_stream_f32(b_mask + pix);  // fused*mask == sum_k (alpha_k vis_k mask) warped_k
      }
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const int kk = j % KL + n * LPP;
        act[n] = xin && kk < KT;
        lg[n] = 0.f;
        vv[n] = vm;
        gxy[n] = make_float2(0.f, 0.f);
        if (act[n]) {
          const unsigned li = (unsigned)kk * HW + pix;
          gxy[n] = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + li), keep);
          if (b_logit) lg[n] = ld_stream_keep_f32(b_logit + li, keep);
          if (b_vis) vv[n] = ld_stream_f32(b_vis + li) * (b_mask ? vm : 1.f);
        }
      }
      // softmax over the K references: own slots first, then the KL preparing lanes of the group
      float m = -CUDART_INF_F;
#pragma unroll
      for (int n = 0; n < NP; ++n)
        if (j % KL + n * LPP < KT) m = fmaxf(m, lg[n]);
      if constexpr (KPOW2) {
#pragma unroll
        for (int s = KL / 2; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
      } else {
        float mm = m;
#pragma unroll
        for (int k = 0; k < KL; ++k) mm = fmaxf(mm, __shfl_sync(FULL, m, gl + k));
        m = mm;
      }
      float e[NP], esum = 0.f;
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        e[n] = expf(lg[n] - m);
        if (j % KL + n * LPP < KT) esum += e[n];
      }
      float ssum;
      if constexpr (KPOW2) {
        ssum = esum;
#pragma unroll
        for (int s = 1; s < KL; s <<= 1) ssum += __shfl_xor_sync(FULL, ssum, s);
      } else {
        ssum = 0.f;
#pragma unroll
        for (int k = 0; k < KL; ++k) ssum += __shfl_sync(FULL, esum, gl + k);
      }
      HotTap t[NP];
      unsigned off[NP];
      bool vis_any = false;
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const float aw = act[n] ? __fdividef(e[n], ssum) * vv[n] : 0.f;  // alpha_k * vis_k * mask
        t[n] = make_hot_tap(gxy[n].x, gxy[n].y, (int)Ws, a.Hs, a.align_corners);
        t[n].nw *= aw;
        t[n].ne *= aw;
        t[n].sw *= aw;
        t[n].se *= aw;
        off[n] = (aw != 0.f) ? (unsigned)t[n].off : 0u;
        vis_any = vis_any || aw != 0.f;
      }
      const bool any = SKIP ? (__ballot_sync(FULL, vis_any) != 0u) : true;

      float2 acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
      if (any) {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const int src = gl + k % LPP, n = k / LPP;
          const unsigned o0 = __shfl_sync(FULL, off[n], src) + (unsigned)k * HWs;
          const char* p0 = f_lane + (size_t)o0 * PIXB;
          const char* p1 = f_lane + (size_t)(o0 + Ws) * PIXB;
          U256 q[4];
          q[0] = ld_gather_u256(p0);
          q[1] = ld_gather_u256(p0 + PIXB);
          q[2] = ld_gather_u256(p1);
          q[3] = ld_gather_u256(p1 + PIXB);
          float wt[4];
          wt[0] = __shfl_sync(FULL, t[n].nw, src);
          wt[1] = __shfl_sync(FULL, t[n].ne, src);
          wt[2] = __shfl_sync(FULL, t[n].sw, src);
          wt[3] = __shfl_sync(FULL, t[n].se, src);
#pragma unroll
          for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
            const float2 w2 = make_float2(wt[tp], wt[tp]);
            const uint32_t wd[8] = {q[tp].lo.x, q[tp].lo.y, q[tp].lo.z, q[tp].lo.w,
                                    q[tp].hi.x, q[tp].hi.y, q[tp].hi.z, q[tp].hi.w};
#pragma unroll
            for (int c = 0; c < 8; ++c)
              acc[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acc[c]);
          }
        }
      }
      if (xin) {
        uint4 o0v, o1v;
        o0v.x = pack_bf16x2(acc[0].x, acc[0].y); o0v.y = pack_bf16x2(acc[1].x, acc[1].y);
        o0v.z = pack_bf16x2(acc[2].x, acc[2].y); o0v.w = pack_bf16x2(acc[3].x, acc[3].y);
        o1v.x = pack_bf16x2(acc[4].x, acc[4].y); o1v.y = pack_bf16x2(acc[5].x, acc[5].y);
        o1v.z = pack_bf16x2(acc[6].x, acc[6].y); o1v.w = pack_bf16x2(acc[7].x, acc[7].y);
        uint4* op = reinterpret_cast<uint4*>(o_lane + (size_t)pix * PIXB);
        st_stream_u128(op, o0v);
        st_stream_u128(op + 1, o1v);
      }
    }
  }

  // =========================== phase B: RGB ===========================
  if (a.rgb != nullptr && a.out_rgb != nullptr) {
    const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
    const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
    const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
    float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
    const int npx = TW * (y_end - y_begin);
    for (int p = threadIdx.x; p < npx; p += 256) {
      const int x = tx * TW + p % TW, y = y_begin + p / TW;  // converged warp, see k_warp_fuse_nhwc_wide
      if constexpr (SKIP) {
        if (x >= (int)W) continue;
        rgb_pixel<KT, SKIP, true>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws);
      } else {
        const int xc = min(x, (int)W - 1);
        rgb_pixel<KT, SKIP, true, kShareRgbTaps>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb,
                                        (unsigned)y * W + (unsigned)xc, HW, HWs, Ws, nullptr, 0, 0, 1, x < (int)W);
      }
    }
  }
}

// NO-SHUFFLE flavour of k_warp_fuse_nhwc_wide (A/B, JAF_WF_WIDE_NOSHFL=1; K <= 4, no visibility input): every lane of the
// 4-lane pixel group loads the flow samples / logits of ALL K references of its pixel (the four lanes read the same
// addresses: one line per load instruction, broadcast) and builds every reference's taps itself, instead of lane k
// preparing reference k and broadcasting corner + weights with 5 shuffles per reference.  Shuffles are LSU / L1TEX
// instructions — the resource that bounds this kernel — while the redundant arithmetic runs on
// pipes that have slack.  Same operations in the same order per channel: bit-identical to the shuffle flavour.
template <int KT, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_warp_fuse_nhwc_wide_ns(const WFArgs a) {
  static_assert(KT <= 4, "register budget");
  constexpr int LPP = 4, PPW = 8, TW = 64;
  constexpr unsigned PIXB = 128;  // bytes of one channels-last pixel (64 bf16)
  int bid = blockIdx.x;
  const int tx = bid % a.tiles_x;
  bid /= a.tiles_x;
  const int ty = bid % a.tiles_y;
  const int b = bid / a.tiles_y;
  const int y_begin = ty * a.rows_per_cta;
  const int y_end = min(a.H, y_begin + a.rows_per_cta);
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  const float* __restrict__ b_logit = a.logits ? a.logits + bK : nullptr;
  const float2* __restrict__ b_grid = reinterpret_cast<const float2*>(a.grid) + bK;
  const float* __restrict__ b_mask = a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr;

  // =========================== phase A: features ===========================
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPP, j = lane % LPP;
    const int x = tx * TW + warp * PPW + g;
    const bool xin = x < (int)W;
    const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + j * 32;
    char* __restrict__ o_lane = reinterpret_cast<char*>(a.out_feat) + (size_t)b * HW * PIXB + j * 32;
    const uint64_t keep = l2_policy_evict_last();
    unsigned pix = (unsigned)y_begin * W + (unsigned)x;
#pragma unroll 1
    for (int y = y_begin; y < y_end; ++y, pix += W) {
      float lg[KT];
      float2 gxy[KT];
      float v = 1.f;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        gxy[k] = make_float2(0.f, 0.f);
        lg[k] = 0.f;
        if (xin) {
          gxy[k] = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + ((unsigned)k * HW + pix)), keep);
          if (b_logit) lg[k] = ld_stream_keep_f32(b_logit + ((unsigned)k * HW + pix), keep);
        }
      }
      if (xin && b_mask) v = ld_stream_f32(b_mask + pix);
      // softmax over the K references, in the association of the shuffle flavour: (e0 + e1) + (e2 + e3) for K = 4
      float m = lg[0];
#pragma unroll
      for (int k = 1; k < KT; ++k) m = fmaxf(m, lg[k]);
#pragma unroll
      for (int k = 0; k < KT; ++k) lg[k] = expf(lg[k] - m);
      float ssum;
      if constexpr (KT == 4) ssum = (lg[0] + lg[1]) + (lg[2] + lg[3]);
      else if constexpr (KT == 3) ssum = ((0.f + lg[0]) + lg[1]) + lg[2];
      else if constexpr (KT == 2) ssum = lg[0] + lg[1];
      else ssum = lg[0];

      float2 acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const float aw = xin ? __fdividef(lg[k], ssum) * v : 0.f;
        HotTap t = make_hot_tap(gxy[k].x, gxy[k].y, (int)Ws, a.Hs, a.align_corners);
        const float wt[4] = {t.nw * aw, t.ne * aw, t.sw * aw, t.se * aw};
        const unsigned o0 = ((aw != 0.f) ? (unsigned)t.off : 0u) + (unsigned)k * HWs;
        const char* p0 = f_lane + (size_t)o0 * PIXB;
        const char* p1 = f_lane + (size_t)(o0 + Ws) * PIXB;
        U256 q[4];
        q[0] = ld_gather_u256(p0);
        q[1] = ld_gather_u256(p0 + PIXB);
        q[2] = ld_gather_u256(p1);
        q[3] = ld_gather_u256(p1 + PIXB);
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
          const float2 w2 = make_float2(wt[tp], wt[tp]);
          const uint32_t wd[8] = {q[tp].lo.x, q[tp].lo.y, q[tp].lo.z, q[tp].lo.w,
                                  q[tp].hi.x, q[tp].hi.y, q[tp].hi.z, q[tp].hi.w};
#pragma unroll
          for (int c = 0; c < 8; ++c)
            acc[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acc[c]);
        }
      }
      if (xin) {
        uint4 o0v, o1v;
        o0v.x = pack_bf16x2(acc[0].x, acc[0].y); o0v.y = pack_bf16x2(acc[1].x, acc[1].y);
        o0v.z = pack_bf16x2(acc[2].x, acc[2].y); o0v.w = pack_bf16x2(acc[3].x, acc[3].y);
        o1v.x = pack_bf16x2(acc[4].x, acc[4].y); o1v.y = pack_bf16x2(acc[5].x, acc[5].y);
        o1v.z = pack_bf16x2(acc[6].x, acc[6].y); o1v.w = pack_bf16x2(acc[7].x, acc[7].y);
        uint4* op = reinterpret_cast<uint4*>(o_lane + (size_t)pix * PIXB);
        st_stream_u128(op, o0v);
        st_stream_u128(op + 1, o1v);
      }
    }
  }

  // =========================== phase B: RGB ===========================
  if (a.rgb != nullptr && a.out_rgb != nullptr) {
    const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
    const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
    const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
    float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
    const int npx = TW * (y_end - y_begin);
    for (int p = threadIdx.x; p < npx; p += 256) {
      const int x = tx * TW + p % TW, y = y_begin + p / TW;
      if (x >= (int)W) continue;
      rgb_pixel<KT, false>(a, rgb_base, b_grid, b_logit, nullptr, nullptr, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws);
    }
  }
}

// The gather issued only by the lanes whose `take` is set (the others keep unspecified registers).
__device__ __forceinline__ uint4 ld_gather_u128_if(const void* p, bool take) {
  uint4 v;
  asm("{\n\t.reg .pred pp;\n\tsetp.ne.u32 pp, %5, 0;\n\t@pp ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];\n\t}"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(p), "r"((uint32_t)take));
  return v;
}

// PAIR flavour (A/B, JAF_WF_PAIR=1): an eight-lane group owns TWO x-adjacent pixels and 8 channels per lane (16 bytes per
// tap).  The ne / se taps of the left pixel ARE the nw / sw taps of the right pixel whenever the right pixel's corner is one
// column further (per reference; true for most pixels of a locally translation-like flow): the group then gathers 3 columns
// x 2 rows = 6 taps for its 8 tap uses and the shared column never leaves the registers (no shuffle: the what-if probes put
// a tap that is not gathered at +10 %, and data shuffles cost more L1TEX time than the gather they replace).
// Pairs whose corners are not adjacent gather the right pixel's own nw / sw (predicated) and select.  Lanes 0..3 of a
// group prepare references 0..3 of the left pixel, lanes 4..7 those of the right pixel; per channel the arithmetic and its
// order are those of k_warp_fuse_nhwc_wide (bit-identical results).
template <int KT, int MINB, bool SKIP>
__global__ void __launch_bounds__(256, MINB)
k_warp_fuse_nhwc_pair(const WFArgs a) {
  static_assert(KT <= 4, "four lanes of a half group, one reference per lane");
  constexpr int TW = 64;         // 8 warps x 4 groups x 2 pixels
  constexpr bool KPOW2 = (KT & (KT - 1)) == 0;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned PIXB = 128;  // bytes of one channels-last pixel (64 bf16)
  int bid = blockIdx.x;
  const int tx = bid % a.tiles_x;
  bid /= a.tiles_x;
  const int ty = bid % a.tiles_y;
  const int b = bid / a.tiles_y;
  const int y_begin = ty * a.rows_per_cta;
  const int y_end = min(a.H, y_begin + a.rows_per_cta);
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  const float* __restrict__ b_logit = a.logits ? a.logits + bK : nullptr;
  const float* __restrict__ b_vis = a.vis ? a.vis + bK : nullptr;
  const int* __restrict__ b_fim = (!a.vis && a.fim) ? a.fim + (size_t)b * HW : nullptr;
  const float2* __restrict__ b_grid = reinterpret_cast<const float2*>(a.grid) + bK;
  const float* __restrict__ b_mask = a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr;

  // =========================== phase A: features ===========================
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, j = lane & 7, gl = g * 8;
    const int half = j >> 2, kk = (j & 3) % KT;  // which pixel of the pair / which reference this lane prepares
    const int xa = tx * TW + warp * 8 + g * 2;   // left pixel of the pair
    const int xm = xa + half;                    // the pixel this lane prepares
    const bool xin_m = xm < (int)W, xin_a = xa < (int)W, xin_b = xa + 1 < (int)W;
    const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + j * 16;
    char* __restrict__ o_lane = reinterpret_cast<char*>(a.out_feat) + (size_t)b * HW * PIXB + j * 16;
    const unsigned lane_in = (unsigned)kk * HW;
    const uint64_t keep = l2_policy_evict_last();
    unsigned pix = (unsigned)y_begin * W + (unsigned)xa;
#pragma unroll 1
    for (int y = y_begin; y < y_end; ++y, pix += W) {
      float lg = 0.f, v = 1.f;
      float2 gxy = make_float2(0.f, 0.f);
      if (xin_m) {
        const unsigned pm = pix + (unsigned)half;
        gxy = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + (lane_in + pm)), keep);
        if (b_logit) lg = ld_stream_keep_f32(b_logit + (lane_in + pm), keep);
        if (b_vis) v = ld_stream_f32(b_vis + (lane_in + pm));
        if (b_fim) v = (ld_stream_s32(b_fim + pm) != -1) ? 1.f : 0.f;
        if (b_mask) v *= ld_stream_f32(b_mask + pm);
      }
      float m = lg, ssum;
      if constexpr (KPOW2) {
#pragma unroll
        for (int s2 = KT / 2; s2 > 0; s2 >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s2));
      } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) m = fmaxf(m, __shfl_sync(FULL, lg, gl + half * 4 + k));
      }
      const float e = expf(lg - m);
      if constexpr (KPOW2) {
        ssum = e;
#pragma unroll
        for (int s2 = 1; s2 < KT; s2 <<= 1) ssum += __shfl_xor_sync(FULL, ssum, s2);
      } else {
        ssum = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) ssum += __shfl_sync(FULL, e, gl + half * 4 + k);
      }
      const float aw = xin_m ? __fdividef(e, ssum) * v : 0.f;
      HotTap t = make_hot_tap(gxy.x, gxy.y, (int)Ws, a.Hs, a.align_corners);
      t.nw *= aw;
      t.ne *= aw;
      t.sw *= aw;
      t.se *= aw;
      const unsigned off = (aw != 0.f) ? (unsigned)t.off : 0u;
      const bool any = SKIP ? (__ballot_sync(FULL, aw != 0.f) != 0u) : true;

      float2 acca[4], accb[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) acca[c] = accb[c] = make_float2(0.f, 0.f);
      if (any) {
#pragma unroll 1  // unrolled, the fallback registers of several references are live at once and spill (148 B at 64 registers)
        for (int k = 0; k < KT; ++k) {
          const unsigned oa = __shfl_sync(FULL, off, gl + k) + (unsigned)k * HWs;
          const unsigned ob = __shfl_sync(FULL, off, gl + 4 + k) + (unsigned)k * HWs;
          const bool coh = ob == oa + 1u;  // uniform over the group
          const char* pa = f_lane + (size_t)oa * PIXB;
          const char* pb = f_lane + (size_t)ob * PIXB;
          uint4 qa[4], qb[4];
          qa[0] = ld_gather_u128(reinterpret_cast<const uint4*>(pa));
          qa[1] = ld_gather_u128(reinterpret_cast<const uint4*>(pa + PIXB));
          qa[2] = ld_gather_u128(reinterpret_cast<const uint4*>(pa + (size_t)Ws * PIXB));
          qa[3] = ld_gather_u128(reinterpret_cast<const uint4*>(pa + (size_t)Ws * PIXB + PIXB));
          qb[1] = ld_gather_u128(reinterpret_cast<const uint4*>(pb + PIXB));
          qb[3] = ld_gather_u128(reinterpret_cast<const uint4*>(pb + (size_t)Ws * PIXB + PIXB));
          qb[0] = ld_gather_u128_if(pb, !coh);
          qb[2] = ld_gather_u128_if(pb + (size_t)Ws * PIXB, !coh);
          float wa[4], wb[4];
          wa[0] = __shfl_sync(FULL, t.nw, gl + k);
          wa[1] = __shfl_sync(FULL, t.ne, gl + k);
          wa[2] = __shfl_sync(FULL, t.sw, gl + k);
          wa[3] = __shfl_sync(FULL, t.se, gl + k);
#pragma unroll
          for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
            const float2 w2 = make_float2(wa[tp], wa[tp]);
            const uint32_t wd[4] = {qa[tp].x, qa[tp].y, qa[tp].z, qa[tp].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) acca[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acca[c]);
          }
          // the right pixel: its nw / sw taps are the left pixel's ne / se when the corners are adjacent
          qb[0].x = coh ? qa[1].x : qb[0].x; qb[0].y = coh ? qa[1].y : qb[0].y;
          qb[0].z = coh ? qa[1].z : qb[0].z; qb[0].w = coh ? qa[1].w : qb[0].w;
          qb[2].x = coh ? qa[3].x : qb[2].x; qb[2].y = coh ? qa[3].y : qb[2].y;
          qb[2].z = coh ? qa[3].z : qb[2].z; qb[2].w = coh ? qa[3].w : qb[2].w;
          wb[0] = __shfl_sync(FULL, t.nw, gl + 4 + k);
          wb[1] = __shfl_sync(FULL, t.ne, gl + 4 + k);
          wb[2] = __shfl_sync(FULL, t.sw, gl + 4 + k);
          wb[3] = __shfl_sync(FULL, t.se, gl + 4 + k);
#pragma unroll
          for (int tp = 0; tp < 4; ++tp) {
            const float2 w2 = make_float2(wb[tp], wb[tp]);
            const uint32_t wd[4] = {qb[tp].x, qb[tp].y, qb[tp].z, qb[tp].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) accb[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, accb[c]);
          }
        }
      }
      if (xin_a) {
        uint4 o;
        o.x = pack_bf16x2(acca[0].x, acca[0].y); o.y = pack_bf16x2(acca[1].x, acca[1].y);
        o.z = pack_bf16x2(acca[2].x, acca[2].y); o.w = pack_bf16x2(acca[3].x, acca[3].y);
        st_stream_u128(reinterpret_cast<uint4*>(o_lane + (size_t)pix * PIXB), o);
      }
      if (xin_b) {
        uint4 o;
        o.x = pack_bf16x2(accb[0].x, accb[0].y); o.y = pack_bf16x2(accb[1].x, accb[1].y);
        o.z = pack_bf16x2(accb[2].x, accb[2].y); o.w = pack_bf16x2(accb[3].x, accb[3].y);
        st_stream_u128(reinterpret_cast<uint4*>(o_lane + (size_t)(pix + 1u) * PIXB), o);
      }
    }
  }

  // =========================== phase B: RGB ===========================
  if (a.rgb != nullptr && a.out_rgb != nullptr) {
    const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
    const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
    const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
    float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
    const int npx = TW * (y_end - y_begin);
    for (int p = threadIdx.x; p < npx; p += 256) {
      const int x = tx * TW + p % TW, y = y_begin + p / TW;
      if (x >= (int)W) continue;
      rgb_pixel<KT, SKIP>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws);
    }
  }
}

// 5..8 references with the K <= 4 kernel's register budget: an EIGHT-lane group owns a pixel column.  Lane j prepares
// reference j (one reference per lane, like the K <= 4 kernel).  Lanes j and j + 4 own the same 16 channels (32 bytes
// per tap, one 256-bit load) and split the REFERENCES: the lower half of the group reduces references 0..3, the upper
// half 4..K-1, each exactly like a 4-lane group of k_warp_fuse_nhwc_wide; the two partial sums are added across the
// halves with 16 shuffles per row (x + y == y + x: both halves hold the same bits) and each half stores 8 of the 16
// channels.  A lane holds one reference's taps at a time: 64 registers, 4 CTAs/SM, instead of the 80 registers / 3
// CTAs of k_warp_fuse_nhwc_wide2 — and unlike the rounds flavour below nothing is serialised.  The sum over the
// references is associated as (0..3) + (4..K-1) instead of left to right: same value within fp32 rounding.
template <int KT, int MINB, bool SKIP>
__global__ void __launch_bounds__(256, MINB)
k_warp_fuse_nhwc_widesk(const WFArgs a) {
  static_assert(KT >= 5 && KT <= 8, "two halves of four lanes, one reference per lane");
  constexpr int LPP = 8, PPW = 4, TW = 32;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned PIXB = 128;  // bytes of one channels-last pixel (64 bf16)
  int bid = blockIdx.x;
  const int tx = bid % a.tiles_x;
  bid /= a.tiles_x;
  const int ty = bid % a.tiles_y;
  const int b = bid / a.tiles_y;
  const int y_begin = ty * a.rows_per_cta;
  const int y_end = min(a.H, y_begin + a.rows_per_cta);
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  const float* __restrict__ b_logit = a.logits ? a.logits + bK : nullptr;
  const float* __restrict__ b_vis = a.vis ? a.vis + bK : nullptr;
  const int* __restrict__ b_fim = (!a.vis && a.fim) ? a.fim + (size_t)b * HW : nullptr;
  const float2* __restrict__ b_grid = reinterpret_cast<const float2*>(a.grid) + bK;
  const float* __restrict__ b_mask = a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr;

  // =========================== phase A: features ===========================
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPP, j = lane % LPP, gl = g * LPP;
    const int half = j >> 2, jc = j & 3;  // which references this lane reduces; which 16 channels it owns
    const bool mine = j < KT;             // lane j prepares reference j
    const int x = tx * TW + warp * PPW + g;
    const bool xin = x < (int)W;
    const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + jc * 32;
    char* __restrict__ o_lane = reinterpret_cast<char*>(a.out_feat) + (size_t)b * HW * PIXB + jc * 32 + half * 16;
    const unsigned lane_in = (unsigned)(mine ? j : 0) * HW;
    const uint64_t keep = l2_policy_evict_last();
    unsigned pix = (unsigned)y_begin * W + (unsigned)x;
#pragma unroll 1
    for (int y = y_begin; y < y_end; ++y, pix += W) {
      float lg = 0.f, v = 1.f;
      float2 gxy = make_float2(0.f, 0.f);
      if (xin) {
        if (mine) {
          gxy = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + (lane_in + pix)), keep);
          if (b_logit) lg = ld_stream_keep_f32(b_logit + (lane_in + pix), keep);
          if (b_vis) v = ld_stream_f32(b_vis + (lane_in + pix));
        }
        if (b_fim) v = (ld_stream_s32(b_fim + pix) != -1) ? 1.f : 0.f;
        if (b_mask) v *= ld_stream_f32(b_mask + pix);
      }
      float m = mine ? lg : -CUDART_INF_F;
#pragma unroll
      for (int s = 4; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
      const float e = mine ? expf(lg - m) : 0.f;
      float ssum = e;
#pragma unroll
      for (int s = 1; s < 8; s <<= 1) ssum += __shfl_xor_sync(FULL, ssum, s);
      const float aw = (xin && mine) ? __fdividef(e, ssum) * v : 0.f;
      HotTap t = make_hot_tap(gxy.x, gxy.y, (int)Ws, a.Hs, a.align_corners);
      t.nw *= aw;
      t.ne *= aw;
      t.sw *= aw;
      t.se *= aw;
      const unsigned off = (aw != 0.f) ? (unsigned)t.off : 0u;
      const bool any = SKIP ? (__ballot_sync(FULL, aw != 0.f) != 0u) : true;

      float2 acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
      if (any) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int kidx = half * 4 + k;  // the reference this half reduces in step k
          const int src = gl + kidx;
          const unsigned osh = __shfl_sync(FULL, off, src);
          float wt[4];
          wt[0] = __shfl_sync(FULL, t.nw, src);
          wt[1] = __shfl_sync(FULL, t.ne, src);
          wt[2] = __shfl_sync(FULL, t.sw, src);
          wt[3] = __shfl_sync(FULL, t.se, src);
          if (KT == 8 || kidx < KT) {  // the upper half has fewer than four references when K < 8
            const unsigned o0 = osh + (unsigned)kidx * HWs;
            const char* p0 = f_lane + (size_t)o0 * PIXB;
            const char* p1 = f_lane + (size_t)(o0 + Ws) * PIXB;
            U256 q[4];
            q[0] = ld_gather_u256(p0);
            q[1] = ld_gather_u256(p0 + PIXB);
            q[2] = ld_gather_u256(p1);
            q[3] = ld_gather_u256(p1 + PIXB);
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
              const float2 w2 = make_float2(wt[tp], wt[tp]);
              const uint32_t wd[8] = {q[tp].lo.x, q[tp].lo.y, q[tp].lo.z, q[tp].lo.w,
                                      q[tp].hi.x, q[tp].hi.y, q[tp].hi.z, q[tp].hi.w};
#pragma unroll
              for (int c = 0; c < 8; ++c)
                acc[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acc[c]);
            }
          }
        }
        // (references 0..3) + (references 4..K-1): both halves end up with the same bits
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc[c].x += __shfl_xor_sync(FULL, acc[c].x, 4);
          acc[c].y += __shfl_xor_sync(FULL, acc[c].y, 4);
        }
      }
      if (xin) {  // lower half: channels 0..7 of its 16, upper half: channels 8..15
        const float2 s0 = half ? acc[4] : acc[0], s1 = half ? acc[5] : acc[1];
        const float2 s2 = half ? acc[6] : acc[2], s3 = half ? acc[7] : acc[3];
        uint4 ov;
        ov.x = pack_bf16x2(s0.x, s0.y);
        ov.y = pack_bf16x2(s1.x, s1.y);
        ov.z = pack_bf16x2(s2.x, s2.y);
        ov.w = pack_bf16x2(s3.x, s3.y);
        st_stream_u128(reinterpret_cast<uint4*>(o_lane + (size_t)pix * PIXB), ov);
      }
    }
  }

  // =========================== phase B: RGB ===========================
  if (a.rgb != nullptr && a.out_rgb != nullptr) {
    const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
    const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
    const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
    float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
    const int npx = TW * (y_end - y_begin);
    for (int p = threadIdx.x; p < npx; p += 256) {
      const int x = tx * TW + p % TW, y = y_begin + p / TW;
      if (x >= (int)W) continue;
      rgb_pixel<KT, SKIP, true>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws);
    }
  }
}

// (A/B flavour, JAF_WF_WIDE8_ROUNDS=1; measured slower than k_warp_fuse_nhwc_wide2, see wf_tune().)
// 5..8 references in ROUNDS of four: the softmax terms of all K references are formed first (two logits per lane), but
// the sample positions / bilinear weights of references 4..7 are built only after references 0..3 have been reduced
// (their flow sample is prefetched one round ahead).  A lane then holds ONE reference's taps at a time, like the K <= 4
// kernel, instead of two: 64 registers / 4 CTAs per SM instead of 80 / 3 for k_warp_fuse_nhwc_wide2.  Same arithmetic.
template <int KT, int MINB, bool SKIP>
__global__ void __launch_bounds__(256, MINB)
k_warp_fuse_nhwc_wide2r(const WFArgs a) {
  static_assert(KT <= 8, "each lane of the 4-lane pixel group prepares at most two references");
  constexpr int LPP = 4, PPW = 8, TW = 64;
  constexpr int NP = (KT + LPP - 1) / LPP;  // references prepared per lane: lane j takes k = j, j + 4
  constexpr int KL = KT < LPP ? KT : LPP;   // lanes of a group that prepare slot 0
  constexpr bool KPOW2 = (KL & (KL - 1)) == 0 && KT % KL == 0;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned PIXB = 128;  // bytes of one channels-last pixel (64 bf16)
  int bid = blockIdx.x;
  const int tx = bid % a.tiles_x;
  bid /= a.tiles_x;
  const int ty = bid % a.tiles_y;
  const int b = bid / a.tiles_y;
  const int y_begin = ty * a.rows_per_cta;
  const int y_end = min(a.H, y_begin + a.rows_per_cta);
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  const float* __restrict__ b_logit = a.logits ? a.logits + bK : nullptr;
  const float* __restrict__ b_vis = a.vis ? a.vis + bK : nullptr;
  const int* __restrict__ b_fim = (!a.vis && a.fim) ? a.fim + (size_t)b * HW : nullptr;
  const float2* __restrict__ b_grid = reinterpret_cast<const float2*>(a.grid) + bK;
  const float* __restrict__ b_mask = a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr;

  // =========================== phase A: features ===========================
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPP, j = lane % LPP, gl = g * LPP;
    const int x = tx * TW + warp * PPW + g;
    const bool xin = x < (int)W;
    const char* __restrict__ f_lane = reinterpret_cast<const char*>(a.feat) + r * KT * (size_t)HWs * PIXB + j * 32;
    char* __restrict__ o_lane = reinterpret_cast<char*>(a.out_feat) + (size_t)b * HW * PIXB + j * 32;
    const uint64_t keep = l2_policy_evict_last();
    unsigned pix = (unsigned)y_begin * W + (unsigned)x;
#pragma unroll 1
    for (int y = y_begin; y < y_end; ++y, pix += W) {
      // ---- softmax terms of all K references (slot n of lane j is reference kk = j % KL + n * LPP)
      float lg[NP];
      float vm = 1.f;
      if (xin) {
        if (b_fim) vm = (ld_stream_s32(b_fim + pix) != -1) ? 1.f : 0.f;
        if (b_mask) vm *= ld_stream_f32(b_mask + pix);  // fused*mask == sum_k (alpha_k vis_k mask) warped_k
      }
      float m = -CUDART_INF_F;
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const int kk = j % KL + n * LPP;
        lg[n] = 0.f;
        if (xin && kk < KT) {
          if (b_logit) lg[n] = ld_stream_keep_f32(b_logit + ((unsigned)kk * HW + pix), keep);
          m = fmaxf(m, lg[n]);
        }
      }
      if constexpr (KPOW2) {
#pragma unroll
        for (int s = KL / 2; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
      } else {
        float mm = m;
#pragma unroll
        for (int k = 0; k < KL; ++k) mm = fmaxf(mm, __shfl_sync(FULL, m, gl + k));
        m = mm;
      }
      float e[NP], esum = 0.f;
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        e[n] = expf(lg[n] - m);
        if (j % KL + n * LPP < KT) esum += e[n];
      }
      float ssum;
      if constexpr (KPOW2) {
        ssum = esum;
#pragma unroll
        for (int s = 1; s < KL; s <<= 1) ssum += __shfl_xor_sync(FULL, ssum, s);
      } else {
        ssum = 0.f;
#pragma unroll
        for (int k = 0; k < KL; ++k) ssum += __shfl_sync(FULL, esum, gl + k);
      }

      float2 acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
      // ---- rounds of four references; the next round's flow sample is in flight while this round is reduced
      float2 gnext = make_float2(0.f, 0.f);
      if (xin && j % KL < KT) gnext = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + ((unsigned)(j % KL) * HW + pix)), keep);
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const int kk = j % KL + n * LPP;
        const bool act = xin && kk < KT;
        const float2 gxy = gnext;
        if (n + 1 < NP) {
          const int kn = j % KL + (n + 1) * LPP;
          gnext = make_float2(0.f, 0.f);
          if (xin && kn < KT) gnext = ld_stream_keep_f32x2(reinterpret_cast<const float*>(b_grid + ((unsigned)kn * HW + pix)), keep);
        }
        float vv = vm;
        if (act && b_vis) vv = ld_stream_f32(b_vis + ((unsigned)kk * HW + pix)) * (b_mask ? vm : 1.f);
        const float aw = act ? __fdividef(e[n], ssum) * vv : 0.f;  // alpha_k * vis_k * mask
        HotTap t = make_hot_tap(gxy.x, gxy.y, (int)Ws, a.Hs, a.align_corners);
        t.nw *= aw;
        t.ne *= aw;
        t.sw *= aw;
        t.se *= aw;
        const unsigned off = (aw != 0.f) ? (unsigned)t.off : 0u;
        const bool any = SKIP ? (__ballot_sync(FULL, aw != 0.f) != 0u) : true;
        if (any) {
#pragma unroll
          for (int kq = 0; kq < LPP; ++kq) {
            const int k = n * LPP + kq;
            if (k >= KT) break;
            const int src = gl + kq;
            const unsigned o0 = __shfl_sync(FULL, off, src) + (unsigned)k * HWs;
            const char* p0 = f_lane + (size_t)o0 * PIXB;
            const char* p1 = f_lane + (size_t)(o0 + Ws) * PIXB;
            U256 q[4];
            q[0] = ld_gather_u256(p0);
            q[1] = ld_gather_u256(p0 + PIXB);
            q[2] = ld_gather_u256(p1);
            q[3] = ld_gather_u256(p1 + PIXB);
            float wt[4];
            wt[0] = __shfl_sync(FULL, t.nw, src);
            wt[1] = __shfl_sync(FULL, t.ne, src);
            wt[2] = __shfl_sync(FULL, t.sw, src);
            wt[3] = __shfl_sync(FULL, t.se, src);
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) {  // nw, ne, sw, se: ATen's accumulation order
              const float2 w2 = make_float2(wt[tp], wt[tp]);
              const uint32_t wd[8] = {q[tp].lo.x, q[tp].lo.y, q[tp].lo.z, q[tp].lo.w,
                                      q[tp].hi.x, q[tp].hi.y, q[tp].hi.z, q[tp].hi.w};
#pragma unroll
              for (int c = 0; c < 8; ++c)
                acc[c] = __ffma2_rn(make_float2(bf16_lo(wd[c]), bf16_hi(wd[c])), w2, acc[c]);
            }
          }
        }
      }
      if (xin) {
        uint4 o0v, o1v;
        o0v.x = pack_bf16x2(acc[0].x, acc[0].y); o0v.y = pack_bf16x2(acc[1].x, acc[1].y);
        o0v.z = pack_bf16x2(acc[2].x, acc[2].y); o0v.w = pack_bf16x2(acc[3].x, acc[3].y);
        o1v.x = pack_bf16x2(acc[4].x, acc[4].y); o1v.y = pack_bf16x2(acc[5].x, acc[5].y);
        o1v.z = pack_bf16x2(acc[6].x, acc[6].y); o1v.w = pack_bf16x2(acc[7].x, acc[7].y);
        uint4* op = reinterpret_cast<uint4*>(o_lane + (size_t)pix * PIXB);
        st_stream_u128(op, o0v);
        st_stream_u128(op + 1, o1v);
      }
    }
  }

  // =========================== phase B: RGB ===========================
  if (a.rgb != nullptr && a.out_rgb != nullptr) {
    const float* __restrict__ rgb_base = a.rgb + r * KT * 3 * (size_t)HWs;
    const float* __restrict__ b_fake = (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr;
    const float* __restrict__ b_conf = (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr;
    float* __restrict__ b_orgb = a.out_rgb + (size_t)b * 3 * HW;
    const int npx = TW * (y_end - y_begin);
    for (int p = threadIdx.x; p < npx; p += 256) {
      const int x = tx * TW + p % TW, y = y_begin + p / TW;  // converged warp, see k_warp_fuse_nhwc_wide
      if constexpr (SKIP) {
        if (x >= (int)W) continue;
        rgb_pixel<KT, SKIP, true>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb, (unsigned)y * W + (unsigned)x, HW, HWs, Ws);
      } else {
        const int xc = min(x, (int)W - 1);
        rgb_pixel<KT, SKIP, true, kShareRgbTaps>(a, rgb_base, b_grid, b_logit, b_vis, b_fim, b_mask, b_fake, b_conf, b_orgb,
                                        (unsigned)y * W + (unsigned)xc, HW, HWs, Ws, nullptr, 0, 0, 1, x < (int)W);
      }
    }
  }
}

// RGB-only calls (no feature tensor): one thread per pixel, the same per-pixel code as phase B
template <int KT, bool SKIP>
__global__ void __launch_bounds__(256)
k_warp_fuse_rgb(const WFArgs a) {
  const unsigned W = (unsigned)a.W, Ws = (unsigned)a.Ws;
  const unsigned HW = (unsigned)a.H * W, HWs = (unsigned)a.Hs * Ws;
  // a CTA is a 32 x 8 pixel tile: the lower tap row of one pixel row is the upper tap row of the next (L1 hits)
  const int tile = blockIdx.x % (a.tiles_x * a.tiles_y);
  const int b = blockIdx.x / (a.tiles_x * a.tiles_y);
  const unsigned x = (unsigned)(tile % a.tiles_x) * 32u + (threadIdx.x & 31u);
  const unsigned y = (unsigned)(tile / a.tiles_x) * 8u + (threadIdx.x >> 5);
  if (y >= (unsigned)a.H) return;  // warp-uniform
  bool live = x < W;
  if constexpr (SKIP) {
    if (!live) return;
  }
  // !SKIP: the warp stays converged (tap sharing by shuffle); lanes beyond the right edge repeat the last column
  const unsigned pix = y * W + (live ? x : W - 1u);
  const size_t r = a.ref_index ? (size_t)a.ref_index[b] : (size_t)b;
  const size_t bK = (size_t)b * KT * HW;
  rgb_pixel_lean<KT, SKIP, !SKIP>(a, a.rgb + r * KT * 3 * (size_t)HWs, reinterpret_cast<const float2*>(a.grid) + bK,
                      a.logits ? a.logits + bK : nullptr, a.vis ? a.vis + bK : nullptr,
                      (!a.vis && a.fim) ? a.fim + (size_t)b * HW : nullptr,
                      a.tgt_mask ? a.tgt_mask + (size_t)b * a.mask_c * HW : nullptr,
                      (a.fake && a.conf) ? a.fake + (size_t)b * 3 * HW : nullptr,
                      (a.fake && a.conf) ? a.conf + (size_t)b * HW : nullptr, a.out_rgb + (size_t)b * 3 * HW, pix, HW,
                      HWs, Ws, live);
}

template <int KT>
void launch_rgb_k(WFArgs a, int, cudaStream_t st) {
  a.tiles_x = (a.W + 31) / 32;
  a.tiles_y = (a.H + 7) / 8;
  const int grid = a.tiles_x * a.tiles_y * a.B;
  if (a.vis != nullptr || a.fim != nullptr) k_warp_fuse_rgb<KT, true><<<grid, 256, 0, st>>>(a);
  else k_warp_fuse_rgb<KT, false><<<grid, 256, 0, st>>>(a);
}

bool launch_rgb(const WFArgs& a, cudaStream_t st) {
  if (a.K > 8 || a.Ws < 2 || a.Hs < 2 || (long)a.H * a.W >= (1L << 29) || a.warped_rgb != nullptr || !a.out_rgb) return false;
  const int grid = jaf::ceil_div((long)a.B * a.H * a.W, 256);
  switch (a.K) {
    case 1: launch_rgb_k<1>(a, grid, st); break;
    case 2: launch_rgb_k<2>(a, grid, st); break;
    case 3: launch_rgb_k<3>(a, grid, st); break;
    case 4: launch_rgb_k<4>(a, grid, st); break;
    case 5: launch_rgb_k<5>(a, grid, st); break;
    case 6: launch_rgb_k<6>(a, grid, st); break;
    case 7: launch_rgb_k<7>(a, grid, st); break;
    case 8: launch_rgb_k<8>(a, grid, st); break;
    default: return false;
  }
  return true;
}

// =====================================================================================
// Generic kernel: one thread per output pixel
// =====================================================================================
template <typename T>
__device__ __forceinline__ float load_as_f32(const T* p);
template <>
__device__ __forceinline__ float load_as_f32<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
template <typename T>
__device__ __forceinline__ void store_from_f32(T* p, float v);
template <>
__device__ __forceinline__ void store_from_f32<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_from_f32<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// IS_RGB: the tensor is a.rgb / a.out_rgb (planar f32, C = 3, per-channel mask, blend, warped output);
// otherwise a.feat / a.out_feat in layout NHWC ? [.,H,W,C] : [.,C,H,W].
// KT > 0: compile-time K (taps and weights stay in registers); KT == 0: runtime K <= 16.
// blockIdx.y slices the channels so that small images with many channels still fill the GPU.
template <typename T, bool NHWC, bool IS_RGB, int KT>
__global__ void __launch_bounds__(256)
k_warp_fuse_generic(const WFArgs a) {
  const long HW = (long)a.H * a.W, HWs = (long)a.Hs * a.Ws;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)a.B * HW) return;
  const int b = (int)(i / HW);
  const long pix = i % HW;
  const long r = a.ref_index ? a.ref_index[b] : b;
  const int K = KT > 0 ? KT : a.K;
  const int C = IS_RGB ? 3 : a.C;
  const int c_begin = IS_RGB ? 0 : (int)blockIdx.y * a.c_chunk;
  const int c_end = IS_RGB ? 3 : min(C, c_begin + a.c_chunk);
  const T* __restrict__ src = IS_RGB ? reinterpret_cast<const T*>(a.rgb) : reinterpret_cast<const T*>(a.feat);
  T* __restrict__ dst = IS_RGB ? reinterpret_cast<T*>(a.out_rgb) : reinterpret_cast<T*>(a.out_feat);

  float aw[KT > 0 ? KT : kMaxKGeneric];
  Tap tp[KT > 0 ? KT : kMaxKGeneric];
  // softmax (reference order: max, exp, running sum, divide)
  if (a.logits) {
    float m = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < K; ++k) m = fmaxf(m, ld_stream_f32(a.logits + ((long)b * K + k) * HW + pix));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      aw[k] = expf(ld_stream_f32(a.logits + ((long)b * K + k) * HW + pix) - m);
      s += aw[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) aw[k] = aw[k] / s;
  } else {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) s += 1.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) aw[k] = 1.0f / s;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float v = 1.f;
    if (a.vis)
      v = ld_stream_f32(a.vis + ((long)b * K + k) * HW + pix);
    else if (a.fim)
      v = (ld_stream_s32(a.fim + (long)b * HW + pix) != -1) ? 1.f : 0.f;
    aw[k] *= v;
    const float2 gxy = ld_stream_f32x2(a.grid + (((long)b * K + k) * HW + pix) * 2);
    tp[k] = make_tap(gxy.x, gxy.y, a.Ws, a.Hs, a.align_corners);
  }
  const float tm1 = a.tgt_mask ? ld_stream_f32(a.tgt_mask + ((long)b * a.mask_c) * HW + pix) : 1.f;
  for (int c = c_begin; c < c_end; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const Tap& t = tp[k];
      const bool need = (aw[k] != 0.f) || (IS_RGB && a.warped_rgb);
      float s = 0.f;
      if (need) {
        const T* base = src + ((long)r * K + k) * C * HWs;
        float v00, v01, v10, v11;
        if (NHWC) {
          const T* p = base + (long)t.off * C + c;
          v00 = load_as_f32(p);
          v01 = load_as_f32(p + (long)t.dx * C);
          v10 = load_as_f32(p + (long)t.dy * C);
          v11 = load_as_f32(p + (long)(t.dy + t.dx) * C);
        } else {
          const T* p = base + (long)c * HWs + t.off;
          v00 = load_as_f32(p);
          v01 = load_as_f32(p + t.dx);
          v10 = load_as_f32(p + t.dy);
          v11 = load_as_f32(p + t.dy + t.dx);
        }
        s = fmaf(v00, t.nw, 0.f);
        s = fmaf(v01, t.ne, s);
        s = fmaf(v10, t.sw, s);
        s = fmaf(v11, t.se, s);
      }
      if (IS_RGB && a.warped_rgb) a.warped_rgb[(((long)b * K + k) * 3 + c) * HW + pix] = s;
      acc = fmaf(aw[k], s, acc);
    }
    if (a.tgt_mask) {
      if (IS_RGB && a.mask_c == 3)
        acc *= ld_stream_f32(a.tgt_mask + ((long)b * 3 + c) * HW + pix);
      else
        acc *= tm1;
    }
    if (IS_RGB && a.fake && a.conf) {
      const float wc = ld_stream_f32(a.conf + (long)b * HW + pix);
      const float fk = ld_stream_f32(a.fake + ((long)b * 3 + c) * HW + pix);
      acc = fk * wc + acc * (1.0f - wc);
    }
    if (dst) {
      const long o = NHWC ? ((long)b * HW + pix) * C + c : ((long)b * C + c) * HW + pix;
      store_from_f32(dst + o, acc);
    }
  }
}

// =====================================================================================
// the bidirectional multi-scale feature warps of SpatioTempoCRN.forward
// (src/crn_model.py:457-566).  Per pyramid level the reference runs
//   flow_s = F.interpolate(flow, size, mode='nearest');  a = grid_sample(prev_pool, (grid + flow_s).permute(0,2,3,1), border)
//                                                        b = grid_sample(pool,      (grid - flow_s).permute(0,2,3,1), border)
// i.e. interpolate + 2 adds + 2 permutes + 2 grid_samples; here one thread per (pixel, channel slice) reads the
// base grid and the full-resolution flow once (nearest index floor(dst * in/out), ATen's
// nearest_neighbor_compute_source_index) and writes both warps.
// =====================================================================================
struct PairArgs {
  const float* feat_fwd;  // warped with grid + flow
  const float* feat_bwd;  // warped with grid - flow
  const float* base_grid; // [B,2,h,w]  channel 0 = x, 1 = y
  const float* flow;      // [B,2,H,W]
  float* out_fwd;
  float* out_bwd;
  int B, C, h, w, H, W, align_corners, c_chunk;
  float scale_y, scale_x;
};

__global__ void __launch_bounds__(256)
k_flow_warp_pair(const PairArgs a) {
  const long hw = (long)a.h * a.w, HWf = (long)a.H * a.W;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)a.B * hw) return;
  const int b = (int)(i / hw);
  const int pix = (int)(i - (long)b * hw);
  const int y = pix / a.w, x = pix - y * a.w;
  // nearest source index: min(floor(dst * scale), in - 1) with scale = in / out in fp32
  const int ys = min((int)floorf(__fmul_rn((float)y, a.scale_y)), a.H - 1);
  const int xs = min((int)floorf(__fmul_rn((float)x, a.scale_x)), a.W - 1);
  const float fx = ld_stream_f32(a.flow + ((long)b * 2 + 0) * HWf + (long)ys * a.W + xs);
  const float fy = ld_stream_f32(a.flow + ((long)b * 2 + 1) * HWf + (long)ys * a.W + xs);
  const float gx = ld_stream_f32(a.base_grid + ((long)b * 2 + 0) * hw + pix);
  const float gy = ld_stream_f32(a.base_grid + ((long)b * 2 + 1) * hw + pix);
  const Tap tf = make_tap(__fadd_rn(gx, fx), __fadd_rn(gy, fy), a.w, a.h, a.align_corners);
  const Tap tb = make_tap(__fsub_rn(gx, fx), __fsub_rn(gy, fy), a.w, a.h, a.align_corners);
  const int c0 = (int)blockIdx.y * a.c_chunk, c1 = min(a.C, c0 + a.c_chunk);
  for (int c = c0; c < c1; ++c) {
    const long plane = ((long)b * a.C + c) * hw;
    if (a.feat_fwd) {
      const float* p = a.feat_fwd + plane + tf.off;
      float s = fmaf(__ldg(p), tf.nw, 0.f);
      s = fmaf(__ldg(p + tf.dx), tf.ne, s);
      s = fmaf(__ldg(p + tf.dy), tf.sw, s);
      s = fmaf(__ldg(p + tf.dy + tf.dx), tf.se, s);
      st_stream_f32(a.out_fwd + plane + pix, s);
    }
    if (a.feat_bwd) {
      const float* p = a.feat_bwd + plane + tb.off;
      float s = fmaf(__ldg(p), tb.nw, 0.f);
      s = fmaf(__ldg(p + tb.dx), tb.ne, s);
      s = fmaf(__ldg(p + tb.dy), tb.sw, s);
      s = fmaf(__ldg(p + tb.dy + tb.dx), tb.se, s);
      st_stream_f32(a.out_bwd + plane + pix, s);
    }
  }
}

int wf_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Tuning knobs (environment, read once per process; jaf_tuning_info() reports the effective values).  The defaults are
// the measured best on B200 for the 240-frame 256^2 K=4 C=64 workload.
struct WFTune {
  int minb_dense;   // JAF_WF_MINB: CTAs/SM of the 8-lane kernel, C=64 K=4, dense (register cap = 65536 / (256 * MINB))
  int minb_skip;    // JAF_WF_MINB_SKIP: the same for the visibility-skipping variant
  int rows;         // JAF_WF_ROWS: rows per iteration of the 8-lane kernel
  int rows_per_cta; // JAF_WF_ROWS_PER_CTA: rows of an 8-lane tile (0 = 32 dense / 16)
  int wide;         // JAF_WF_WIDE: 1 wide-lane kernel for C=64 without a visibility input, 2 also with one, 0 never
  int wide_minb;    // JAF_WF_WIDE_MINB: CTAs/SM of the wide kernel, K <= 4
  int wide_minb8;   // JAF_WF_WIDE_MINB8: the same for K = 5..8
  int wide_rows;    // JAF_WF_WIDE_ROWS_PER_CTA: rows of a wide tile (64 columns)
  int rgb_merge;    // JAF_WF_RGB_MERGE: RGB planes inside the feature row loop (1) or as a second pass (0)
  int minb_poses;   // JAF_WF_MINB_POSES: CTAs/SM of the pose-driven kernel, K <= 4 (4 or 5)
  int noshfl;       // JAF_WF_WIDE_NOSHFL: K <= 4 wide kernel in which every lane prepares all K references (no shuffles)
  int pair;         // JAF_WF_PAIR: K <= 4, C = 64 on 8-lane groups that own a PAIR of adjacent pixels (shared tap column in registers)
  int wide8_splitk; // JAF_WF_WIDE8_SPLITK: K = 5..8 on 8-lane groups whose halves split the references (64 registers, 4 CTAs/SM)
  int splitk_rows;  // JAF_WF_SPLITK_ROWS: rows of a split-K tile (32 columns)
  int wide8_rounds; // JAF_WF_WIDE8_ROUNDS: K = 5..8 in rounds of four references (64 registers, 4 CTAs/SM) instead of
                    // two references per lane (80 registers, 3 CTAs/SM)
};
const WFTune& wf_tune() {
  static const WFTune t = [] {
    WFTune v;
    v.minb_dense = wf_env("JAF_WF_MINB", 4);
    if (v.minb_dense < 3 || v.minb_dense > 6) v.minb_dense = 4;
    // visibility-skipping flavours, list-driven phases (round 2): 4 CTAs/SM (64 registers, no spills) beat 5 (48 registers,
    // spills in the list loop): SMPL flows 291.6 k vs 281.4 k frames/s, from poses 208.5 k vs 201.9 k
    v.minb_skip = wf_env("JAF_WF_MINB_SKIP", 4);
    if (v.minb_skip < 4 || v.minb_skip > 6) v.minb_skip = 4;
    v.rows = wf_env("JAF_WF_ROWS", 2);
    v.rows_per_cta = wf_env("JAF_WF_ROWS_PER_CTA", 0);
    v.wide = wf_env("JAF_WF_WIDE", 1);
    v.wide_minb = wf_env("JAF_WF_WIDE_MINB", 4);
    if (v.wide_minb < 3 || v.wide_minb > 5) v.wide_minb = 4;
    v.wide_minb8 = wf_env("JAF_WF_WIDE_MINB8", 3) == 4 ? 4 : 3;
    // 64 x 8 tiles: measured (profiles/r02_bench_ab.jsonl) dense flows 0.710 of peak at 8 rows vs 0.715-0.726 at 16,
    // piecewise-affine "hard" flows 0.712 vs 0.680, random permutation 0.369 vs 0.285: the small tile is the robust one
    v.wide_rows = wf_env("JAF_WF_WIDE_ROWS_PER_CTA", 8);
    if (v.wide_rows < 1) v.wide_rows = 8;
    // measured on B200 (profiles/r02_bench_ab.jsonl): merged 93.8 k frames/s dense / 83.7 k hard, two-pass 94.8 k / 89.9 k:
    // the scalar RGB taps issued from the 4-lane groups cost more L1 wavefronts and issue slots than the second pass
    v.rgb_merge = wf_env("JAF_WF_RGB_MERGE", 0);
    v.minb_poses = wf_env("JAF_WF_MINB_POSES", 4) == 5 ? 5 : 4;
    // measured (profiles/r02_bench_ab.jsonl, 512^2 K=8): rounds of four 9.6 k frames/s (0.525) vs two references per lane
    // 11.1 k (0.605): the second round's sample positions serialise behind the first round's reduction, which costs more
    // than the fourth resident CTA brings
    v.wide8_rounds = wf_env("JAF_WF_WIDE8_ROUNDS", 0);
    // measured (profiles/r02_bench_ab.jsonl, 512^2 K=8, same box): split-K 11.0 k frames/s (0.600) dense / 10.4 k (0.565) hard
    // vs two references per lane 11.2 k (0.612) / 10.6 k (0.580) — the fourth resident CTA buys nothing at this shape
    // (SM clock 1.7-1.8 GHz under sw_power_cap for the whole 21 ms launch), so the flavour stays an A/B knob
    v.wide8_splitk = wf_env("JAF_WF_WIDE8_SPLITK", 0);
    v.pair = wf_env("JAF_WF_PAIR", 0);
    v.noshfl = wf_env("JAF_WF_WIDE_NOSHFL", 0);
    v.splitk_rows = wf_env("JAF_WF_SPLITK_ROWS", 16);
    if (v.splitk_rows < 1) v.splitk_rows = 16;
    return v;
  }();
  return t;
}
int wf_minb(bool skip) { return skip ? wf_tune().minb_skip : wf_tune().minb_dense; }

template <int LPP, int KV>
bool launch_nhwc_kc(const WFArgs& a, int grid, cudaStream_t st) {
  if constexpr (KV <= LPP) {
    const bool skip = a.vis != nullptr || a.fim != nullptr;
    // coverage flags + compacted list of covered pixels of a tile (the SKIP flavours; see the kernel's smem layout)
    const size_t tile_px = (size_t)(8 * (32 / LPP)) * a.rows_per_cta;
    const size_t skip_smem = ((tile_px + 15) & ~(size_t)15) + tile_px * sizeof(int2);
    if constexpr (LPP == 8 && KV == 4) {  // the headline shape carries the occupancy variants
      const int mb = wf_minb(skip);
      const int rows = wf_tune().rows;
#define JAF_V(MB, R) if (mb == MB && rows == R) { \
        jaf::note_kernel("k_warp_fuse_nhwc<LPP=8,K=4,MINB=%d,SKIP=%d,ROWS=%d>", MB, (int)skip, R); \
        if (skip) k_warp_fuse_nhwc<8, 4, MB, true, R><<<grid, 256, skip_smem, st>>>(a); else k_warp_fuse_nhwc<8, 4, MB, false, R><<<grid, 256, 0, st>>>(a); return true; }
      JAF_V(4, 1) JAF_V(5, 1) JAF_V(6, 1) JAF_V(3, 2) JAF_V(4, 2) JAF_V(5, 2) JAF_V(6, 2)
#undef JAF_V
    }
    jaf::note_kernel("k_warp_fuse_nhwc<LPP=%d,K=%d,MINB=6,SKIP=%d,ROWS=1>", LPP, KV, (int)skip);
    if (skip) k_warp_fuse_nhwc<LPP, KV, 6, true><<<grid, 256, skip_smem, st>>>(a);
    else k_warp_fuse_nhwc<LPP, KV, 6, false><<<grid, 256, 0, st>>>(a);
    return true;
  } else {
    return false;
  }
}

template <int LPP>
bool launch_nhwc_k(const WFArgs& a, int grid, cudaStream_t st) {
  switch (a.K) {
    case 1: return launch_nhwc_kc<LPP, 1>(a, grid, st);
    case 2: return launch_nhwc_kc<LPP, 2>(a, grid, st);
    case 3: return launch_nhwc_kc<LPP, 3>(a, grid, st);
    case 4: return launch_nhwc_kc<LPP, 4>(a, grid, st);
    case 5: return launch_nhwc_kc<LPP, 5>(a, grid, st);
    case 6: return launch_nhwc_kc<LPP, 6>(a, grid, st);
    case 7: return launch_nhwc_kc<LPP, 7>(a, grid, st);
    case 8: return launch_nhwc_kc<LPP, 8>(a, grid, st);
    default: return false;
  }
}

template <typename T, bool NHWC, bool IS_RGB>
void launch_generic(const WFArgs& a, int gx, int gy, cudaStream_t st) {
  const dim3 grid((unsigned)gx, (unsigned)gy);
  switch (a.K) {
    case 1: k_warp_fuse_generic<T, NHWC, IS_RGB, 1><<<grid, 256, 0, st>>>(a); break;
    case 2: k_warp_fuse_generic<T, NHWC, IS_RGB, 2><<<grid, 256, 0, st>>>(a); break;
    case 3: k_warp_fuse_generic<T, NHWC, IS_RGB, 3><<<grid, 256, 0, st>>>(a); break;
    case 4: k_warp_fuse_generic<T, NHWC, IS_RGB, 4><<<grid, 256, 0, st>>>(a); break;
    default: k_warp_fuse_generic<T, NHWC, IS_RGB, 0><<<grid, 256, 0, st>>>(a); break;
  }
}

// Returns true when the hot kernel was launched.
bool launch_nhwc(WFArgs a, cudaStream_t st) {
  const int lpp = a.C / 8;
  if (a.C % 8 != 0 || !(lpp == 4 || lpp == 8 || lpp == 16 || lpp == 32) || a.K > lpp || a.K > 8) return false;
  if (a.Ws < 2 || a.Hs < 2 || (long)a.H * a.W >= (1L << 29)) return false;
  // C = 64 without a visibility input goes to the wide-lane kernel (256-bit gathers): 94.3 k frames/s against 90.9 k
  // for k_warp_fuse_nhwc<8,4> on the headline workload (4 CTAs/SM, 64 registers, 64 x 16 pixel tiles; 5 or 6 CTAs/SM
  // spill).  With visibility the narrow skipping variant stays ahead (216 k vs 199 k frames/s on SMPL flows).
  const WFTune& tn = wf_tune();
  const int wide_minb = tn.wide_minb, wide_minb8 = tn.wide_minb8;
  const bool wide_ok = tn.wide == 2 || (tn.wide == 1 && a.vis == nullptr && a.fim == nullptr);
  if (wide_ok && a.C == 64 && a.K <= 8 && (reinterpret_cast<uintptr_t>(a.feat) & 31u) == 0 &&
      (reinterpret_cast<uintptr_t>(a.out_feat) & 31u) == 0) {
    a.tiles_x = (a.W + 63) / 64;
    a.rows_per_cta = a.H < tn.wide_rows ? a.H : tn.wide_rows;
    a.tiles_y = (a.H + a.rows_per_cta - 1) / a.rows_per_cta;
    const long gridw = (long)a.tiles_x * a.tiles_y * a.B;
    if (gridw <= 0x7fffffffL) {
      const bool skip = a.vis != nullptr || a.fim != nullptr;
      const bool rgbm = tn.rgb_merge != 0 && a.rgb != nullptr && a.out_rgb != nullptr && a.mask_c == 1;
      if (tn.noshfl != 0 && a.K <= 4 && !skip) {
#define JAF_WN(KV) if (a.K == KV) { \
        jaf::note_kernel("k_warp_fuse_nhwc_wide_ns<K=%d,MINB=%d>", KV, tn.noshfl == 3 ? 3 : 4); \
        if (tn.noshfl == 3) k_warp_fuse_nhwc_wide_ns<KV, 3><<<(unsigned)gridw, 256, 0, st>>>(a); \
        else k_warp_fuse_nhwc_wide_ns<KV, 4><<<(unsigned)gridw, 256, 0, st>>>(a); \
        return true; }
        JAF_WN(1) JAF_WN(2) JAF_WN(3) JAF_WN(4)
#undef JAF_WN
      }
      if (tn.pair != 0 && a.K <= 4 && !skip) {
#define JAF_WP(KV) if (a.K == KV) { \
        jaf::note_kernel("k_warp_fuse_nhwc_pair<K=%d,MINB=%d,SKIP=0>", KV, tn.pair == 3 ? 3 : 4); \
        if (tn.pair == 3) k_warp_fuse_nhwc_pair<KV, 3, false><<<(unsigned)gridw, 256, 0, st>>>(a); \
        else k_warp_fuse_nhwc_pair<KV, 4, false><<<(unsigned)gridw, 256, 0, st>>>(a); \
        return true; }
        JAF_WP(1) JAF_WP(2) JAF_WP(3) JAF_WP(4)
#undef JAF_WP
      }
#define JAF_W(KV, MB) if (a.K == KV && wide_minb == MB) { \
        jaf::note_kernel("k_warp_fuse_nhwc_wide<K=%d,MINB=%d,SKIP=%d,RGBM=%d>", KV, MB, (int)skip, (int)rgbm); \
        if (skip) { if (rgbm) k_warp_fuse_nhwc_wide<KV, MB, true, true><<<(unsigned)gridw, 256, 0, st>>>(a); else k_warp_fuse_nhwc_wide<KV, MB, true, false><<<(unsigned)gridw, 256, 0, st>>>(a); } \
        else { if (rgbm) k_warp_fuse_nhwc_wide<KV, MB, false, true><<<(unsigned)gridw, 256, 0, st>>>(a); else k_warp_fuse_nhwc_wide<KV, MB, false, false><<<(unsigned)gridw, 256, 0, st>>>(a); } \
        return true; }
      JAF_W(4, 3) JAF_W(4, 4) JAF_W(4, 5)
      if (a.K <= 3) {
        const int wide_minb = 4;  // K < 4 carries one occupancy variant
        JAF_W(1, 4) JAF_W(2, 4) JAF_W(3, 4)
      }
      if (a.K >= 5 && tn.wide8_splitk != 0 && tn.wide8_rounds == 0) {
        WFArgs s = a;
        s.tiles_x = (s.W + 31) / 32;
        s.rows_per_cta = s.H < tn.splitk_rows ? s.H : tn.splitk_rows;
        s.tiles_y = (s.H + s.rows_per_cta - 1) / s.rows_per_cta;
        const long grids = (long)s.tiles_x * s.tiles_y * s.B;
        if (grids <= 0x7fffffffL) {
#define JAF_WSK(KV) if (s.K == KV) { \
          jaf::note_kernel("k_warp_fuse_nhwc_widesk<K=%d,MINB=4,SKIP=%d>", KV, (int)skip); \
          if (skip) k_warp_fuse_nhwc_widesk<KV, 4, true><<<(unsigned)grids, 256, 0, st>>>(s); else k_warp_fuse_nhwc_widesk<KV, 4, false><<<(unsigned)grids, 256, 0, st>>>(s); \
          return true; }
          JAF_WSK(5) JAF_WSK(6) JAF_WSK(7) JAF_WSK(8)
#undef JAF_WSK
        }
      }
#define JAF_W8R(KV) if (a.K == KV && tn.wide8_rounds != 0) { \
        jaf::note_kernel("k_warp_fuse_nhwc_wide2r<K=%d,MINB=4,SKIP=%d>", KV, (int)skip); \
        if (skip) k_warp_fuse_nhwc_wide2r<KV, 4, true><<<(unsigned)gridw, 256, 0, st>>>(a); else k_warp_fuse_nhwc_wide2r<KV, 4, false><<<(unsigned)gridw, 256, 0, st>>>(a); \
        return true; }
      JAF_W8R(5) JAF_W8R(6) JAF_W8R(7) JAF_W8R(8)
#undef JAF_W8R
#define JAF_W8(KV) if (a.K == KV) { \
        jaf::note_kernel("k_warp_fuse_nhwc_wide2<K=%d,MINB=%d,SKIP=%d>", KV, wide_minb8, (int)skip); \
        if (wide_minb8 == 4) { if (skip) k_warp_fuse_nhwc_wide2<KV, 4, true><<<(unsigned)gridw, 256, 0, st>>>(a); else k_warp_fuse_nhwc_wide2<KV, 4, false><<<(unsigned)gridw, 256, 0, st>>>(a); } \
        else { if (skip) k_warp_fuse_nhwc_wide2<KV, 3, true><<<(unsigned)gridw, 256, 0, st>>>(a); else k_warp_fuse_nhwc_wide2<KV, 3, false><<<(unsigned)gridw, 256, 0, st>>>(a); } \
        return true; }
      JAF_W8(5) JAF_W8(6) JAF_W8(7) JAF_W8(8)
#undef JAF_W8
#undef JAF_W
    }
  }
  const int ppw = 32 / lpp;
  const int tw = 8 * ppw;
  a.tiles_x = (a.W + tw - 1) / tw;
  const int rows_env = tn.rows_per_cta;
  const bool tuned_dense = lpp == 8 && a.K == 4 && a.vis == nullptr && a.fim == nullptr;
  const int rows_dflt = rows_env > 0 ? rows_env : (tuned_dense ? 32 : 16);
  a.rows_per_cta = a.H < rows_dflt ? a.H : rows_dflt;
  a.tiles_y = (a.H + a.rows_per_cta - 1) / a.rows_per_cta;
  const long grid = (long)a.tiles_x * a.tiles_y * a.B;
  if (grid > 0x7fffffffL) return false;
  switch (lpp) {
    case 4: return launch_nhwc_k<4>(a, (int)grid, st);
    case 8: return launch_nhwc_k<8>(a, (int)grid, st);
    case 16: return launch_nhwc_k<16>(a, (int)grid, st);
    case 32: return launch_nhwc_k<32>(a, (int)grid, st);
  }
  return false;
}

}  // namespace

extern "C" int jaf_warp_fuse(const JafWarpFuseParams* p) {
  JAF_REQUIRE(p != nullptr, "null params");
  JAF_REQUIRE(p->B >= 0 && p->K >= 1 && p->H > 0 && p->W > 0 && p->Hs > 0 && p->Ws > 0, "bad sizes");
  JAF_REQUIRE(p->grid != nullptr, "grid is required");
  JAF_REQUIRE((long)p->Hs * p->Ws < (1L << 29), "reference image too large");
  JAF_REQUIRE((reinterpret_cast<uintptr_t>(p->grid) & 7u) == 0, "grid must be 8-byte aligned");
  const bool want_rgb = p->rgb && (p->out_rgb || p->warped_rgb);
  const bool want_feat = p->feat && p->out_feat && p->C > 0;
  JAF_REQUIRE(want_rgb || want_feat, "nothing to do: need rgb+out_rgb and/or feat+out_feat");
  JAF_REQUIRE(!p->tgt_mask || p->mask_c == 1 || p->mask_c == 3, "mask_c must be 1 or 3");
  JAF_REQUIRE(p->feat_layout == JAF_LAYOUT_PLANAR || p->feat_layout == JAF_LAYOUT_NHWC, "bad feat_layout");
  JAF_REQUIRE(p->feat_dtype == JAF_DTYPE_F32 || p->feat_dtype == JAF_DTYPE_BF16, "bad feat_dtype");
  if (p->B == 0) return JAF_OK;
  cudaStream_t st = jaf::as_stream(p->stream);

  WFArgs a = {};
  a.B = p->B; a.K = p->K; a.H = p->H; a.W = p->W; a.Hs = p->Hs; a.Ws = p->Ws; a.C = p->C;
  a.align_corners = p->align_corners; a.mask_c = p->tgt_mask ? p->mask_c : 1;
  a.rows_per_cta = 0; a.tiles_x = 0; a.tiles_y = 0;
  a.rgb = p->rgb; a.feat = p->feat; a.ref_index = p->ref_index; a.grid = p->grid; a.logits = p->logits;
  a.vis = p->vis; a.fim = p->fim; a.tgt_mask = p->tgt_mask; a.fake = p->fake; a.conf = p->conf;
  a.out_rgb = p->out_rgb; a.out_feat = p->out_feat; a.warped_rgb = p->warped_rgb;

  int launches = 0;
  bool rgb_done = !want_rgb, feat_done = !want_feat;
  // hot path: channels-last bf16 features, 16-byte aligned, RGB fused in unless per-reference
  // warps are requested
  if (want_feat && p->feat_layout == JAF_LAYOUT_NHWC && p->feat_dtype == JAF_DTYPE_BF16 &&
      (reinterpret_cast<uintptr_t>(p->feat) & 15u) == 0 && (reinterpret_cast<uintptr_t>(p->out_feat) & 15u) == 0) {
    WFArgs h = a;
    const bool fuse_rgb = want_rgb && !p->warped_rgb && p->out_rgb;
    if (!fuse_rgb) {
      h.rgb = nullptr;
      h.out_rgb = nullptr;
    }
    h.warped_rgb = nullptr;
    if (launch_nhwc(h, st)) {
      ++launches;
      feat_done = true;
      if (fuse_rgb) rgb_done = true;
    }
  }
  const long npix = (long)p->B * p->H * p->W;
  const int grid = jaf::ceil_div(npix, 256);
  if (!rgb_done || !feat_done) JAF_REQUIRE(p->K <= kMaxKGeneric, "K > 16 is not supported by the generic kernel");
  if (!rgb_done) {
    if (launch_rgb(a, st)) {
      if (feat_done) jaf::note_kernel("k_warp_fuse_rgb<K=%d,SKIP=%d>", a.K, (int)(a.vis != nullptr || a.fim != nullptr));
    } else {
      launch_generic<float, false, true>(a, grid, 1, st);
      if (feat_done) jaf::note_kernel("k_warp_fuse_generic<f32,planar,rgb,K=%d>", a.K);
    }
    ++launches;
  }
  if (!feat_done) {
    WFArgs f = a;
    f.warped_rgb = nullptr;
    // slice the channels over blockIdx.y until ~600k threads are in flight
    long want = (600000 + npix - 1) / npix;
    if (want < 1) want = 1;
    if (want > p->C) want = p->C;
    if (want > 65535) want = 65535;
    f.c_chunk = (int)((p->C + want - 1) / want);
    const int ny = (p->C + f.c_chunk - 1) / f.c_chunk;
    const bool nhwc = p->feat_layout == JAF_LAYOUT_NHWC;
    if (p->feat_dtype == JAF_DTYPE_F32) {
      if (nhwc) launch_generic<float, true, false>(f, grid, ny, st);
      else      launch_generic<float, false, false>(f, grid, ny, st);
    } else {
      if (nhwc) launch_generic<__nv_bfloat16, true, false>(f, grid, ny, st);
      else      launch_generic<__nv_bfloat16, false, false>(f, grid, ny, st);
    }
    jaf::note_kernel("k_warp_fuse_generic<%s,%s,K=%d>", p->feat_dtype == JAF_DTYPE_F32 ? "f32" : "bf16",
                     nhwc ? "nhwc" : "planar", a.K);
    ++launches;
  }
  return jaf::finish_launch("jaf_warp_fuse", launches);
}

namespace {

// The pose-driven flavour exists for the headline layout: channels-last bf16 features with C = 64, K <= 8.
bool poses_supported(int C, int K, int feat_layout, int feat_dtype) {
  if (K < 1 || K > 8) return false;
  if (C == 0) return true;  // RGB planes only (the reference's per-frame warp_image call, test/conv_pro_test.py:255-278)
  return C == 64 && feat_layout == JAF_LAYOUT_NHWC && feat_dtype == JAF_DTYPE_BF16;
}

template <int KV>
void launch_poses_k(const WFArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  constexpr int R = KV <= 4 ? 2 : 1;
  // JAF_WF_MINB_POSES: 5 CTAs/SM (48 registers, the row loop spills a few loop invariants) or 4 (64, no spills)
  const int mb = (KV <= 4 && wf_tune().minb_poses == 5) ? 5 : 4;  // default 4 (wf_tune)
  jaf::note_kernel("k_warp_fuse_nhwc<LPP=8,K=%d,MINB=%d,SKIP=1,ROWS=%d,POSES=1>", KV, mb, R);
  if constexpr (KV <= 4) {
    if (mb == 5) {
      k_warp_fuse_nhwc<8, KV, 5, true, R, true><<<grid, 256, smem, st>>>(a);
      return;
    }
  }
  k_warp_fuse_nhwc<8, KV, 4, true, R, true><<<grid, 256, smem, st>>>(a);
}

}  // namespace

extern "C" int jaf_warp_fuse_from_poses_supported(int C, int K, int feat_layout, int feat_dtype) {
  return poses_supported(C, K, feat_layout, feat_dtype) ? 1 : 0;
}

extern "C" int jaf_warp_fuse_from_poses(const JafWarpFuseParams* p, const JafPoseFlowParams* q) {
  JAF_REQUIRE(p != nullptr && q != nullptr, "null params");
  JAF_REQUIRE(p->B >= 0 && p->K >= 1 && p->H > 0 && p->W == p->H && p->Hs > 1 && p->Ws > 1,
              "bad sizes (the output frame is the target raster: H == W)");
  JAF_REQUIRE(q->tgt_cam && q->tgt_verts && q->src_cam && q->src_verts && q->faces_idx && q->workspace, "null pose input");
  JAF_REQUIRE(q->V > 0 && q->F >= 0, "bad mesh sizes");
  const bool want_feat = p->feat && p->out_feat && p->C > 0;
  JAF_REQUIRE(want_feat || (p->rgb && p->out_rgb), "nothing to do: need rgb+out_rgb and/or feat+out_feat");
  JAF_REQUIRE(!p->tgt_mask || p->mask_c == 1 || p->mask_c == 3, "mask_c must be 1 or 3");
  JAF_REQUIRE((long)p->Hs * p->Ws < (1L << 29) && (long)p->H * p->W < (1L << 29), "image too large");
  JAF_REQUIRE(!p->warped_rgb, "per-reference warps are not available in the pose-driven flavour");
  JAF_REQUIRE(!q->T || (reinterpret_cast<uintptr_t>(q->T) & 7u) == 0, "T must be 8-byte aligned");
  JAF_REQUIRE(!want_feat || ((reinterpret_cast<uintptr_t>(p->feat) & 15u) == 0 && (reinterpret_cast<uintptr_t>(p->out_feat) & 15u) == 0),
              "features must be 16-byte aligned");
  if (!poses_supported(want_feat ? p->C : 0, p->K, p->feat_layout, p->feat_dtype)) {
    jaf::set_error("jaf_warp_fuse_from_poses: this build fuses C = 64 channels-last bf16 features with K <= 8; use "
                   "jaf_cal_flow_multi + jaf_warp_fuse for other shapes");
    return JAF_ERR_UNSUPPORTED;
  }
  if (p->B == 0) return JAF_OK;
  cudaStream_t st = jaf::as_stream(p->stream);
  int launches = 0;
  const int st1 = jaf::raster_keys_from_poses(q->tgt_cam, q->tgt_verts, q->faces_idx, p->B, q->V, q->F, p->H, q->eye_z,
                                              q->near_, q->far_, q->workspace, st, &launches,
                                              (q->flags & JAF_POSES_KEYS_CLEAN) != 0);
  if (st1 != JAF_OK) return st1;

  WFArgs a = {};
  a.B = p->B; a.K = p->K; a.H = p->H; a.W = p->W; a.Hs = p->Hs; a.Ws = p->Ws; a.C = p->C;
  a.align_corners = p->align_corners; a.mask_c = p->tgt_mask ? p->mask_c : 1;
  const bool fuse_rgb = p->rgb && p->out_rgb;
  a.rgb = fuse_rgb ? p->rgb : nullptr; a.out_rgb = fuse_rgb ? p->out_rgb : nullptr;
  a.feat = want_feat ? p->feat : nullptr; a.out_feat = want_feat ? p->out_feat : nullptr;
  a.ref_index = p->ref_index; a.logits = p->logits;
  a.tgt_mask = p->tgt_mask; a.fake = p->fake; a.conf = p->conf;
  a.zkeys = static_cast<unsigned long long*>(q->workspace);
  a.leave_clean = (q->flags & JAF_POSES_LEAVE_CLEAN) != 0;
  a.tgt_cam = q->tgt_cam; a.tgt_verts = q->tgt_verts; a.src_cam = q->src_cam; a.src_verts = q->src_verts;
  a.fidx = q->faces_idx; a.V = q->V; a.eye_z = q->eye_z; a.near_ = q->near_; a.far_ = q->far_;
  a.fim_out = q->fim; a.T_out = q->T;
  const int tw = 32;  // 8-lane groups: 4 pixel columns per warp, 8 warps
  a.tiles_x = (a.W + tw - 1) / tw;
  // 32 x 32 tiles for K <= 4 (212 k vs 208 k frames/s at 32 x 16); K = 5..8 keeps 16 rows (the K flows of a tile must fit
  // the 48 KB of default dynamic shared memory)
  int rows_dflt = wf_tune().rows_per_cta > 0 ? wf_tune().rows_per_cta : (a.K <= 4 ? 32 : 16);
  // small batches (the reference's per-frame loop runs batch 1): keep at least one wave of CTAs on the GPU
  if (wf_tune().rows_per_cta <= 0 && (long)a.tiles_x * ((a.H + rows_dflt - 1) / rows_dflt) * a.B < 148L * 4) rows_dflt = 8;
  a.rows_per_cta = a.H < rows_dflt ? a.H : rows_dflt;
  a.tiles_y = (a.H + a.rows_per_cta - 1) / a.rows_per_cta;
  const long grid = (long)a.tiles_x * a.tiles_y * a.B;
  JAF_REQUIRE(grid <= 0x7fffffffL, "too many tiles");
  const size_t tile_px = (size_t)tw * a.rows_per_cta;
  // K flows + coverage flag per tile pixel, and the compacted list of covered pixels
  const size_t smem = (((size_t)a.K * tile_px * sizeof(float2) + tile_px + 15) & ~(size_t)15) + tile_px * sizeof(int2);
  JAF_REQUIRE(smem <= 48 * 1024, "tile does not fit shared memory (JAF_WF_ROWS_PER_CTA too large)");
  switch (a.K) {
    case 1: launch_poses_k<1>(a, (unsigned)grid, smem, st); break;
    case 2: launch_poses_k<2>(a, (unsigned)grid, smem, st); break;
    case 3: launch_poses_k<3>(a, (unsigned)grid, smem, st); break;
    case 4: launch_poses_k<4>(a, (unsigned)grid, smem, st); break;
    case 5: launch_poses_k<5>(a, (unsigned)grid, smem, st); break;
    case 6: launch_poses_k<6>(a, (unsigned)grid, smem, st); break;
    case 7: launch_poses_k<7>(a, (unsigned)grid, smem, st); break;
    default: launch_poses_k<8>(a, (unsigned)grid, smem, st); break;
  }
  return jaf::finish_launch("jaf_warp_fuse_from_poses", launches + 1);
}

extern "C" int jaf_tuning_info(char* buf, int n) {
  const WFTune& t = wf_tune();
  char tmp[512];
  const int len = snprintf(tmp, sizeof(tmp),
                           "JAF_WF_WIDE=%d JAF_WF_WIDE_MINB=%d JAF_WF_WIDE_MINB8=%d JAF_WF_WIDE_ROWS_PER_CTA=%d "
                           "JAF_WF_RGB_MERGE=%d JAF_WF_MINB=%d JAF_WF_MINB_SKIP=%d JAF_WF_MINB_POSES=%d JAF_WF_ROWS=%d "
                           "JAF_WF_ROWS_PER_CTA=%d JAF_WF_WIDE8_ROUNDS=%d JAF_WF_WIDE8_SPLITK=%d JAF_WF_SPLITK_ROWS=%d JAF_WF_PAIR=%d JAF_WF_WIDE_NOSHFL=%d",
                           t.wide, t.wide_minb, t.wide_minb8, t.wide_rows, t.rgb_merge, t.minb_dense, t.minb_skip,
                           t.minb_poses, t.rows, t.rows_per_cta, t.wide8_rounds, t.wide8_splitk, t.splitk_rows, t.pair, t.noshfl);
  if (buf != nullptr && n > 0) snprintf(buf, (size_t)n, "%s", tmp);
  return len + 1;
}

extern "C" int jaf_warp_image(const float* src, const float* grid, int N, int C, int Hs, int Ws, int H, int W,
                              int align_corners, float* out, void* stream) {
  JAF_REQUIRE(src && grid && out, "null pointer");
  JafWarpFuseParams p = {};
  p.B = N; p.K = 1; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws; p.C = C;
  p.align_corners = align_corners;
  p.feat_layout = JAF_LAYOUT_PLANAR;
  p.feat_dtype = JAF_DTYPE_F32;
  p.grid = grid;
  if (C == 3) {  // an RGB frame (src/cal_flow.py:37-39): the templated per-pixel RGB kernel
    p.C = 0;
    p.rgb = src;
    p.out_rgb = out;
  } else {
    p.feat = src;
    p.out_feat = out;
  }
  p.stream = stream;
  return jaf_warp_fuse(&p);
}


extern "C" int jaf_flow_warp_pair(const float* feat_fwd, const float* feat_bwd, const float* base_grid, const float* flow,
                                  int B, int C, int h, int w, int H, int W, int align_corners, float* out_fwd,
                                  float* out_bwd, void* stream) {
  JAF_REQUIRE(base_grid && flow, "null pointer");
  JAF_REQUIRE((feat_fwd == nullptr) == (out_fwd == nullptr) && (feat_bwd == nullptr) == (out_bwd == nullptr),
              "each feature tensor needs its output");
  JAF_REQUIRE(feat_fwd || feat_bwd, "nothing to warp");
  JAF_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "bad sizes");
  PairArgs a;
  a.feat_fwd = feat_fwd; a.feat_bwd = feat_bwd; a.base_grid = base_grid; a.flow = flow;
  a.out_fwd = out_fwd; a.out_bwd = out_bwd;
  a.B = B; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.align_corners = align_corners;
  a.scale_y = (float)H / (float)h;  // ATen compute_scales_value for mode='nearest' with an explicit size
  a.scale_x = (float)W / (float)w;
  // small maps with many channels: slice the channels over blockIdx.y so the grid still fills the GPU
  const long pixels = (long)B * h * w;
  int slices = 1;
  while (slices < C && pixels * slices < 148L * 2048 && slices < 64) slices *= 2;
  a.c_chunk = jaf::ceil_div(C, slices);
  dim3 grid(jaf::ceil_div(pixels, 256), jaf::ceil_div(C, a.c_chunk));
  k_flow_warp_pair<<<grid, 256, 0, jaf::as_stream(stream)>>>(a);
  return jaf::finish_launch("k_flow_warp_pair");
}
