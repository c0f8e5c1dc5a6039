// Face-index / barycentric-weight rasteriser and transfer-flow construction (rows a1-a9).
//
// The reference tests every pixel against every face
// (NR/cuda/rasterize_cuda_kernel.cu:70-169: 65,536 px x 13,776 faces per frame).  SMPL
// faces project to ~1 px^2, so this file inverts the loop: one thread per FACE walks a
// conservative pixel bounding box (~10^4 x fewer tests) and resolves visibility with a
// 64-bit atomicMin on (depth_bits << 32 | face_index), which reproduces the reference's
// "strict '<' while scanning faces in ascending order" rule (lowest index wins a tie).
// A second per-pixel pass decodes the winner and writes fim / wim (rows already flipped,
// NR/rasterize.py:334-338), or composes the transfer flow directly (src/nmr.py:617-659).
//
// Bit-exactness: every float operation below is pinned with round-to-nearest intrinsics
// in the exact order — including the FMA contractions — that nvcc emits for the
// reference source on sm_100 (read from its SASS; DESIGN.md "Raster arithmetic").
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"
#include "raster_math.cuh"

namespace {

using namespace jaf_raster;
constexpr int kBigBox = 64;     // boxes above this many pixels are walked by the whole warp
constexpr int kHugeBox = 2048;  // boxes above this go to a queue that a follow-up launch spreads over the GPU
constexpr int kHugeCap = 1024;  // queue capacity (entries beyond it are walked by their warp)
constexpr int kHugeRun = 128;   // entries x chunks covered by the follow-up grid per launch
constexpr int kHugeChunks = 32;

struct HugeEntry {
  int b, fn, x_lo, y_lo, bw, npx, pad0, pad1;
};
struct HugeQueue {
  unsigned int count;  // entries pushed MINUS ONE: the queue is reset by the same 0xFF memset that clears the z-buffer
  unsigned int pad[15];
  HugeEntry e[kHugeCap];
};

// Order-preserving float -> uint map so atomicMin on the packed key orders by depth first.
__device__ __forceinline__ unsigned int float_order_bits(float z) {
  const unsigned int u = __float_as_uint(z);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct Box {
  int x_lo, x_hi, y_lo, y_hi;
};

// Conservative pixel box of a front-facing face.  The reference accepts a pixel when three fp32
// edge comparisons do not fail, so a pixel can pass while lying marginally outside the exact
// triangle; the margin grows with the sliver-ness of the face, and degenerate faces (zero or
// non-finite determinant — where the reference's tests hold on a whole line or everywhere)
// are tested against the full image exactly like the reference does.
// Extra whole pixels added to the analytic margin.  0 is enough: the box already runs from floor(min) to ceil(max), so an
// excluded pixel centre is at least one pixel away from the box, while the fp32 edge tests can only accept points within
// ~1e-5 px of an edge line (times the sliver factor, which the analytic term covers).  Validated bit-for-bit against the
// reference kernels on 72 SMPL frames and 380k sub-pixel triangles (tools/probes/raster_stress.py); JAF_RASTER_MARGIN=1
// restores the conservative setting (2.5x more pixel tests per face).
__constant__ float c_margin_base = 0.0f;

__device__ __forceinline__ Box face_box(const float* px, const float* py, float den, int is) {
  Box bx;
  const float xmin = fminf(px[0], fminf(px[1], px[2])), xmax = fmaxf(px[0], fmaxf(px[1], px[2]));
  const float ymin = fminf(py[0], fminf(py[1], py[2])), ymax = fmaxf(py[0], fmaxf(py[1], py[2]));
  const float ex0 = px[1] - px[0], ey0 = py[1] - py[0], ex1 = px[2] - px[1], ey1 = py[2] - py[1];
  const float ex2 = px[0] - px[2], ey2 = py[0] - py[2];
  const float l2 = fmaxf(ex0 * ex0 + ey0 * ey0, fmaxf(ex1 * ex1 + ey1 * ey1, ex2 * ex2 + ey2 * ey2));
  const float aden = fabsf(den);
  const bool finite = isfinite(xmin) && isfinite(xmax) && isfinite(ymin) && isfinite(ymax) && isfinite(l2);
  if (!finite || !(aden > 0.0f)) {
    bx.x_lo = 0; bx.y_lo = 0; bx.x_hi = is - 1; bx.y_hi = is - 1;
    return bx;
  }
  // A pixel outside the exact triangle can only pass the three fp32 edge comparisons if its distance d to
  // each violated edge line satisfies d <= 3.6e-7 * D (three roundings of 2^-24 per product; D = distance
  // to the edge's vertex <= ~1.5 * (is + |p|max) px); beyond the apex of a needle of apex angle ~ 1/sliver
  // that allows an overshoot t <= 2 * sliver * d.  Twice that bound is the margin.
  const float sliver = l2 / aden;  // ~2.3 for an equilateral face, large for slivers
  const float pmax = fmaxf(fmaxf(fabsf(xmin), fabsf(xmax)), fmaxf(fabsf(ymin), fabsf(ymax)));
  const float mf = fminf((float)is, c_margin_base + floorf(2.2e-6f * ((float)is + pmax) * sliver));
  const float lim = 2.0f * (float)is + 4.0f;
  bx.x_lo = max(0, (int)fmaxf(floorf(xmin) - mf, -lim));
  bx.y_lo = max(0, (int)fmaxf(floorf(ymin) - mf, -lim));
  bx.x_hi = min(is - 1, (int)fminf(ceilf(xmax) + mf, lim));
  bx.y_hi = min(is - 1, (int)fminf(ceilf(ymax) + mf, lim));
  return bx;
}

__device__ __forceinline__ void zbuf_try(unsigned long long* __restrict__ zb, const float* f, const float* inv,
                                         int b, int fn, int xi, int yi, int is, float near_, float far_) {
  float w[3], zp;
  if (pixel_test(f, inv, xi, yi, is, near_, far_, w, &zp)) {
    const unsigned long long key = ((unsigned long long)float_order_bits(zp) << 32) | (unsigned int)fn;
    atomicMin(zb + ((size_t)b * is + yi) * is + xi, key);
  }
}

// ---- pass 1: one thread per (batch, face)
template <bool PROJECT>
__global__ void __launch_bounds__(256)
k_raster_scatter(const float* __restrict__ faces_xyz, const float* __restrict__ cam,
                 const float* __restrict__ verts, const int* __restrict__ fidx, int B, int V, int F, int is,
                 float eye_z, float near_, float far_, unsigned long long* __restrict__ zbuf,
                 float* __restrict__ faces_out, HugeQueue* __restrict__ huge) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  float f[9], inv[9];
  Box bx = {0, -1, 0, -1};
  int b = 0, fn = 0;
  if (i < (long)B * F) {
    b = (int)(i / F);
    fn = (int)(i % F);
    if (PROJECT) {
      load_face_projected(cam, verts, fidx, b, fn, V, eye_z, f);
      if (faces_out) {
#pragma unroll
        for (int k = 0; k < 9; ++k) faces_out[i * 9 + k] = f[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) f[k] = faces_xyz[i * 9 + k];
    }
    if (!face_is_back(f)) {
      float px[3], py[3];
      const float den = face_setup(f, is, inv, px, py);
      bx = face_box(px, py, den, is);
    }
  }
  const int bw = bx.x_hi - bx.x_lo + 1, bh = bx.y_hi - bx.y_lo + 1;
  const int npx = (bw > 0 && bh > 0) ? bw * bh : 0;
  bool big = npx > kBigBox;
  if (npx > kHugeBox) {  // needle-like or degenerate face with a (near) full-image box: defer
    const unsigned slot = atomicAdd(&huge->count, 1u) + 1u;  // the counter starts at 0xFFFFFFFF
    if (slot < (unsigned)kHugeCap) {
      HugeEntry en = {b, fn, bx.x_lo, bx.y_lo, bw, npx, 0, 0};
      huge->e[slot] = en;
      big = false;
      bx.y_hi = bx.y_lo - 1;  // nothing left to do here
    }
  }
  if (!big) {
    for (int yi = bx.y_lo; yi <= bx.y_hi; ++yi)
      for (int xi = bx.x_lo; xi <= bx.x_hi; ++xi) zbuf_try(zbuf, f, inv, b, fn, xi, yi, is, near_, far_);
  }
  // faces with large boxes: the whole warp walks each of them in turn
  unsigned m = __ballot_sync(0xffffffffu, big);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    float ff[9], fi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      ff[k] = __shfl_sync(0xffffffffu, f[k], src);
      fi[k] = __shfl_sync(0xffffffffu, inv[k], src);
    }
    const int sx = __shfl_sync(0xffffffffu, bx.x_lo, src), sy = __shfl_sync(0xffffffffu, bx.y_lo, src);
    const int sw = __shfl_sync(0xffffffffu, bw, src), sn = __shfl_sync(0xffffffffu, npx, src);
    const int sb = __shfl_sync(0xffffffffu, b, src), sf = __shfl_sync(0xffffffffu, fn, src);
    for (int t = (int)lane; t < sn; t += 32)
      zbuf_try(zbuf, ff, fi, sb, sf, sx + t % sw, sy + t / sw, is, near_, far_);
  }
}

// ---- pass 1, flattened: the same work as k_raster_scatter with the divergence taken out.
// One thread per face leaves a warp waiting for its largest box (measured: 2,000 warp instructions per 32 faces for ~110
// useful pixel tests) and half the lanes idle behind the back-face cull.  Here a CTA of 256 faces (a) culls and compacts
// the front faces into shared memory, (b) runs the per-face setup on the compacted list (full warps), (c) takes a prefix
// sum over the box sizes and (d) splits the CTA's concatenated pixel tests into 256 equal contiguous shares, one per thread
// (one binary search for the first face, then a linear walk): every lane does one test per step whatever the boxes look like.  Same per-(face, pixel)
// arithmetic and the same 64-bit atomicMin: identical z-buffer.
struct FlatFace {
  float f[9], inv[9];
  int x_lo, y_lo, bw, b, fn;
};

template <bool PROJECT>
__global__ void __launch_bounds__(256)
k_raster_scatter_flat(const float* __restrict__ faces_xyz, const float* __restrict__ cam,
                      const float* __restrict__ verts, const int* __restrict__ fidx, int B, int V, int F, int is,
                      float eye_z, float near_, float far_, unsigned long long* __restrict__ zbuf,
                      float* __restrict__ faces_out, HugeQueue* __restrict__ huge) {
  __shared__ FlatFace s_face[256];
  __shared__ int s_pref[257];   // exclusive prefix sum of the box sizes
  __shared__ int s_wsum[8];
  __shared__ unsigned s_n;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  // (a) load / project, cull, compact
  if (i < (long)B * F) {
    float f[9];
    const int b = (int)(i / F), fn = (int)(i % F);
    if (PROJECT) {
      load_face_projected(cam, verts, fidx, b, fn, V, eye_z, f);
      if (faces_out) {
#pragma unroll
        for (int k = 0; k < 9; ++k) faces_out[i * 9 + k] = f[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) f[k] = faces_xyz[i * 9 + k];
    }
    if (!face_is_back(f)) {
      FlatFace& e = s_face[atomicAdd(&s_n, 1u)];
#pragma unroll
      for (int k = 0; k < 9; ++k) e.f[k] = f[k];
      e.b = b;
      e.fn = fn;
    }
  }
  __syncthreads();
  const int n = (int)s_n;
  // (b) per-face setup + box on the compacted faces; needle / degenerate faces with (near) full-image boxes are deferred
  int npx = 0;
  if ((int)threadIdx.x < n) {
    FlatFace& e = s_face[threadIdx.x];
    float f[9], inv[9], px[3], py[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = e.f[k];
    const float den = face_setup(f, is, inv, px, py);
    const Box bx = face_box(px, py, den, is);
    const int bw = bx.x_hi - bx.x_lo + 1, bh = bx.y_hi - bx.y_lo + 1;
    npx = (bw > 0 && bh > 0) ? bw * bh : 0;
    if (npx > kHugeBox) {
      const unsigned slot = atomicAdd(&huge->count, 1u) + 1u;  // the counter starts at 0xFFFFFFFF
      if (slot < (unsigned)kHugeCap) {
        HugeEntry en = {e.b, e.fn, bx.x_lo, bx.y_lo, bw, npx, 0, 0};
        huge->e[slot] = en;
        npx = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) e.inv[k] = inv[k];
    e.x_lo = bx.x_lo;
    e.y_lo = bx.y_lo;
    e.bw = bw;
  }
  // (c) block-wide exclusive prefix sum of npx
  int incl = npx;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane >= d) incl += t;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < (int)warp) base += s_wsum[w];
  s_pref[threadIdx.x] = base + incl - npx;
  if (threadIdx.x == 255) s_pref[256] = base + incl;
  __syncthreads();
  const int total = s_pref[256];
  // (d) the CTA's concatenated pixel tests, an equal contiguous share per thread: one binary search for the face of the
  // first test, then a linear walk (faces with empty boxes are stepped over)
  const int chunk = (total + 255) >> 8;
  int t = (int)threadIdx.x * chunk;
  const int t_end = min(total, t + chunk);
  if (t < t_end) {
    int lo = 0, hi = n - 1;  // largest j with s_pref[j] <= t
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_pref[mid] <= t) lo = mid; else hi = mid - 1;
    }
    int j = lo;
    for (; t < t_end; ++t) {
      while (t >= s_pref[j + 1]) ++j;
      const FlatFace& e = s_face[j];
      const int local = t - s_pref[j];
      zbuf_try(zbuf, e.f, e.inv, e.b, e.fn, e.x_lo + local % e.bw, e.y_lo + local / e.bw, is, near_, far_);
    }
  }
}

// ---- pass 1b: the deferred huge boxes, each spread over kHugeChunks CTAs
template <bool PROJECT>
__global__ void __launch_bounds__(256)
k_raster_huge(const float* __restrict__ faces_xyz, const float* __restrict__ cam, const float* __restrict__ verts,
              const int* __restrict__ fidx, int V, int F, int is, float eye_z, float near_, float far_,
              unsigned long long* __restrict__ zbuf, const HugeQueue* __restrict__ huge, int first) {
  const unsigned count = min(huge->count + 1u, (unsigned)kHugeCap);
  for (unsigned e = (unsigned)first + blockIdx.x; e < count; e += gridDim.x) {
  const HugeEntry en = huge->e[e];
  float f[9], inv[9], px[3], py[3];
  if (PROJECT) {
    load_face_projected(cam, verts, fidx, en.b, en.fn, V, eye_z, f);
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = faces_xyz[((size_t)en.b * F + en.fn) * 9 + k];
  }
  face_setup(f, is, inv, px, py);
  const int per = (en.npx + kHugeChunks - 1) / kHugeChunks;
  const int t0 = blockIdx.y * per, t1 = min(en.npx, t0 + per);
  for (int t = t0 + (int)threadIdx.x; t < t1; t += blockDim.x)
    zbuf_try(zbuf, f, inv, en.b, en.fn, en.x_lo + t % en.bw, en.y_lo + t / en.bw, is, near_, far_);
  }
}

// ---- pass 2: one thread per pixel; decode the winner, recompute its weights with the same pinned
// arithmetic (deterministic => identical bits), write outputs with rows flipped, and/or compose the
// transfer flow (src/nmr.py:644-653; src/cal_flow.py:30-31 for the source x, -y_raster).
template <bool PROJECT, bool COMPOSE>
__global__ void __launch_bounds__(256)
k_raster_resolve(const unsigned long long* __restrict__ zbuf, const float* __restrict__ faces_xyz,
                 const float* __restrict__ cam, const float* __restrict__ verts, const int* __restrict__ fidx,
                 const float* __restrict__ src_cam, const float* __restrict__ src_verts, int B, int V, int F,
                 int is, float eye_z, float near_, float far_, int flip_rows, int* __restrict__ fim,
                 float* __restrict__ wim, float* __restrict__ depth, float* __restrict__ T, int Ksrc,
                 int keep_bg = 0, float* __restrict__ face_inv_map = nullptr) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * is * is) return;
  const int b = (int)(i / ((long)is * is));
  const int pn = (int)(i % ((long)is * is));
  const int yi = pn / is, xi = pn % is;
  const int yo = flip_rows ? (is - 1 - yi) : yi;
  const size_t o = ((size_t)b * is + yo) * is + xi;
  const unsigned long long key = zbuf[i];
  int fn = -1;
  float w[3] = {0.f, 0.f, 0.f};
  float zp = far_;
  if (key != kEmptyKey) {
    fn = (int)(unsigned int)(key & 0xffffffffull);
    float f[9], inv[9], px[3], py[3];
    if (PROJECT) {
      load_face_projected(cam, verts, fidx, b, fn, V, eye_z, f);
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) f[k] = faces_xyz[((size_t)b * F + fn) * 9 + k];
    }
    face_setup(f, is, inv, px, py);
    pixel_test(f, inv, xi, yi, is, near_, far_, w, &zp);
    if (face_inv_map) {  // rasterize_cuda_kernel.cu:163-167 (return_depth)
#pragma unroll
      for (int k = 0; k < 9; ++k) face_inv_map[o * 9 + k] = inv[k];
    }
  }
  if (keep_bg) {
    // the extension's own contract (rasterize_cuda_kernel.cu:156-168): the caller pre-fills the outputs
    // (NR/rasterize.py:50-52) and only pixels covered by a face are written
    if (fn >= 0) {
      if (fim) fim[o] = fn;
      if (wim) {
        wim[o * 3 + 0] = w[0];
        wim[o * 3 + 1] = w[1];
        wim[o * 3 + 2] = w[2];
      }
      if (depth) depth[o] = zp;
    }
    return;
  }
  if (fim) fim[o] = fn;
  if (wim) {
    if ((is & 31) == 0) {
      // a warp's 32 pixels are 96 consecutive floats of wim: hand the values round so that every store instruction
      // writes 128 contiguous bytes (three strided 4-byte stores per lane would touch every sector three times)
      const unsigned lane = threadIdx.x & 31u;
      float* wbase = wim + (o - lane) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const unsigned e = lane + 32u * c, src = e / 3u, comp = e - 3u * src;
        const float v0 = __shfl_sync(0xffffffffu, w[0], src), v1 = __shfl_sync(0xffffffffu, w[1], src);
        const float v2 = __shfl_sync(0xffffffffu, w[2], src);
        wbase[e] = comp == 0 ? v0 : (comp == 1 ? v1 : v2);
      }
    } else {
      wim[o * 3 + 0] = w[0];
      wim[o * 3 + 1] = w[1];
      wim[o * 3 + 2] = w[2];
    }
  }
  if (depth) depth[o] = zp;
  if (COMPOSE) {
    // one target raster serves Ksrc source poses: T[b, ks] for ks < Ksrc (Ksrc == 1: float_estimate.cal_flow)
    int vi[3] = {0, 0, 0};
    if (fn >= 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) vi[k] = fidx[fn * 3 + k];
    }
    for (int ks = 0; ks < Ksrc; ++ks) {
      float tx = -2.0f, ty = -2.0f;  // src/nmr.py:627
      const size_t sb = (size_t)b * Ksrc + ks;
      if (fn >= 0) {
        const float s = src_cam[sb * 3 + 0], ctx = src_cam[sb * 3 + 1], cty = src_cam[sb * 3 + 2];
        float ax[3], ay[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float* p = src_verts + (sb * V + vi[k]) * 3;
          ax[k] = __fmul_rn(s, __fadd_rn(p[0], ctx));
          ay[k] = __fmul_rn(s, __fadd_rn(p[1], cty));  // -(-(s*(Y+ty))): raster flip then cal_flow.py:31
        }
        tx = __fadd_rn(__fadd_rn(__fmul_rn(ax[0], w[0]), __fmul_rn(ax[1], w[1])), __fmul_rn(ax[2], w[2]));
        ty = __fadd_rn(__fadd_rn(__fmul_rn(ay[0], w[0]), __fmul_rn(ay[1], w[1])), __fmul_rn(ay[2], w[2]));
      }
      reinterpret_cast<float2*>(T)[(sb * is + yo) * is + xi] = make_float2(tx, ty);
    }
  }
}

// ---- a1-a3 standalone (when the caller wants the [B,F,3,3] tensor)
__global__ void __launch_bounds__(256)
k_project_gather(const float* __restrict__ cam, const float* __restrict__ verts, const int* __restrict__ fidx,
                 int B, int V, int F, float eye_z, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * F) return;
  float f[9];
  load_face_projected(cam, verts, fidx, (int)(i / F), (int)(i % F), V, eye_z, f);
#pragma unroll
  for (int k = 0; k < 9; ++k) out[i * 9 + k] = f[k];
}

// ---- kernel_1's visible output (rasterize_cuda_kernel.cu:24-67): the per-face inverse matrices; culled faces are
// left as the caller filled them (zeros, NR/rasterize.py:164)
__global__ void __launch_bounds__(256)
k_faces_inv(const float* __restrict__ faces_xyz, long n, int is, float* __restrict__ faces_inv) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float f[9], inv[9], px[3], py[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) f[k] = faces_xyz[i * 9 + k];
  if (face_is_back(f)) return;
  face_setup(f, is, inv, px, py);
#pragma unroll
  for (int k = 0; k < 9; ++k) faces_inv[i * 9 + k] = inv[k];
}

// ---- a9 standalone: src/nmr.py:617-659 on explicit fim / wim tensors
__global__ void __launch_bounds__(256)
k_flow_compose(const float* __restrict__ src_pts, int stride, int negate_y, const int* __restrict__ fim,
               const float* __restrict__ wim, int B, int F, long HW, float* __restrict__ T) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * HW) return;
  const int b = (int)(i / HW);
  const int fn = ld_stream_s32(fim + i);
  float tx = -2.0f, ty = -2.0f;
  if (fn != -1) {
    const float* p = src_pts + ((size_t)b * F + fn) * 3 * stride;
    const float w0 = ld_stream_f32(wim + i * 3 + 0), w1 = ld_stream_f32(wim + i * 3 + 1),
                w2 = ld_stream_f32(wim + i * 3 + 2);
    float ax0 = p[0], ax1 = p[stride], ax2 = p[2 * stride];
    float ay0 = p[1], ay1 = p[stride + 1], ay2 = p[2 * stride + 1];
    if (negate_y) {
      ay0 = -ay0;
      ay1 = -ay1;
      ay2 = -ay2;
    }
    tx = __fadd_rn(__fadd_rn(__fmul_rn(ax0, w0), __fmul_rn(ax1, w1)), __fmul_rn(ax2, w2));
    ty = __fadd_rn(__fadd_rn(__fmul_rn(ay0, w0), __fmul_rn(ay1, w1)), __fmul_rn(ay2, w2));
  }
  reinterpret_cast<float2*>(T)[i] = make_float2(tx, ty);
}

size_t zbuf_bytes(int B, int is) { return ((size_t)B * is * is * sizeof(unsigned long long) + 255) & ~(size_t)255; }

// pass 1 (+1b): clear the z-buffer and the huge-box queue, scatter the faces, spread the deferred boxes
template <bool PROJECT>
int run_pass1(const float* faces_xyz, const float* cam, const float* verts, const int* fidx, int B, int V, int F,
              int is, float eye_z, float near_, float far_, void* workspace, float* faces_out, cudaStream_t st,
              int* launches, bool keys_clean = false) {
  // debug knob; a __constant__ symbol exists once per device, so it is set on every device that rasterises
  static const char* margin_env = getenv("JAF_RASTER_MARGIN");
  if (margin_env != nullptr) {
    static jaf::PerDeviceOnce margin_once;
    const int dev = jaf::current_device();
    if (!margin_once.done(dev)) {
      const float v = (float)atof(margin_env);
      JAF_CUDA(cudaMemcpyToSymbolAsync(c_margin_base, &v, sizeof(float), 0, cudaMemcpyHostToDevice, st));
      JAF_CUDA(cudaStreamSynchronize(st));
      margin_once.mark(dev);
    }
  }
  auto* zb = static_cast<unsigned long long*>(workspace);
  auto* hq = reinterpret_cast<HugeQueue*>(static_cast<char*>(workspace) + zbuf_bytes(B, is));
  // one clear for the z-buffer keys (all ones = empty) and the queue header behind them (counter = -1)
  if (!keys_clean)
    JAF_CUDA(cudaMemsetAsync(zb, 0xff, zbuf_bytes(B, is) + 64, st));
  else
    JAF_CUDA(cudaMemsetAsync(hq, 0xff, 64, st));
  if (F > 0) {
    static const bool flat = [] {  // JAF_RASTER_FLAT=0: the one-thread-per-face scatter (A/B runs)
      const char* e = getenv("JAF_RASTER_FLAT");
      return e == nullptr || atoi(e) != 0;
    }();
    if (flat)
      k_raster_scatter_flat<PROJECT><<<jaf::ceil_div((long)B * F, 256), 256, 0, st>>>(faces_xyz, cam, verts, fidx, B, V, F,
                                                                                   is, eye_z, near_, far_, zb, faces_out, hq);
    else
      k_raster_scatter<PROJECT><<<jaf::ceil_div((long)B * F, 256), 256, 0, st>>>(faces_xyz, cam, verts, fidx, B, V, F, is,
                                                                              eye_z, near_, far_, zb, faces_out, hq);
    // ~0.2 deferred faces per frame: a batch-1 call (the reference's per-frame loop) launches 4 x 32 CTAs, not 128 x 32
    const int huge_x = B * 2 < kHugeRun ? (B * 2 < 4 ? 4 : B * 2) : kHugeRun;
    k_raster_huge<PROJECT><<<dim3(huge_x, kHugeChunks), 256, 0, st>>>(faces_xyz, cam, verts, fidx, V, F, is, eye_z,
                                                                      near_, far_, zb, hq, 0);
    *launches += 2;
  }
  return JAF_OK;
}

int check_raster_args(int B, int F, int is) {
  return B >= 0 && F >= 0 && is > 0 && is <= 16384;
}

}  // namespace

namespace jaf {
// Pass 1 of the rasteriser on projected poses (z-buffer keys only), for callers that resolve the keys themselves
// (the pose-driven fused warp kernel).
int raster_keys_from_poses(const float* cam, const float* verts, const int* fidx, int B, int V, int F, int is,
                           float eye_z, float near_, float far_, void* workspace, cudaStream_t st, int* launches,
                           bool keys_clean) {
  if (!check_raster_args(B, F, is) || V <= 0) {
    set_error("raster_keys_from_poses: bad sizes");
    return JAF_ERR_INVALID;
  }
  return run_pass1<true>(nullptr, cam, verts, fidx, B, V, F, is, eye_z, near_, far_, workspace, nullptr, st, launches,
                         keys_clean);
}
}  // namespace jaf

extern "C" {

int jaf_project_gather(const float* cam, const float* verts, const int32_t* faces_idx, int B, int V, int F,
                       float eye_z, float* faces_xyz, void* stream) {
  JAF_REQUIRE(cam && verts && faces_idx && faces_xyz, "null pointer");
  JAF_REQUIRE(B >= 0 && V > 0 && F >= 0, "bad sizes");
  if ((long)B * F == 0) return JAF_OK;
  k_project_gather<<<jaf::ceil_div((long)B * F, 256), 256, 0, jaf::as_stream(stream)>>>(cam, verts, faces_idx, B, V,
                                                                                      F, eye_z, faces_xyz);
  return jaf::finish_launch("k_project_gather");
}

size_t jaf_raster_workspace_bytes(int B, int image_size) {
  if (B <= 0 || image_size <= 0) return 0;
  return zbuf_bytes(B, image_size) + sizeof(HugeQueue);
}

int jaf_raster_fim_wim(const float* faces_xyz, int B, int F, int image_size, float near_, float far_,
                       int flip_rows, int32_t* fim, float* wim, float* depth, void* workspace, void* stream) {
  JAF_REQUIRE((faces_xyz || F == 0) && fim && wim && workspace, "null pointer");
  JAF_REQUIRE(check_raster_args(B, F, image_size), "bad sizes");
  if (B == 0) return JAF_OK;
  cudaStream_t st = jaf::as_stream(stream);
  auto* zb = static_cast<unsigned long long*>(workspace);
  const long npix = (long)B * image_size * image_size;
  int launches = 0;
  {
    const int st1 = run_pass1<false>(faces_xyz, nullptr, nullptr, nullptr, B, 0, F, image_size, 0.f, near_, far_,
                                     workspace, nullptr, st, &launches);
    if (st1 != JAF_OK) return st1;
  }
  k_raster_resolve<false, false><<<jaf::ceil_div(npix, 256), 256, 0, st>>>(
      zb, faces_xyz, nullptr, nullptr, nullptr, nullptr, nullptr, B, 0, F, image_size, 0.f, near_, far_, flip_rows,
      fim, wim, depth, nullptr, 0);
  return jaf::finish_launch("jaf_raster_fim_wim", launches + 1);
}

int jaf_forward_face_index_map(const float* faces, int32_t* face_index_map, float* weight_map, float* depth_map,
                               float* face_inv_map, float* faces_inv, int B, int F, int image_size, float near_,
                               float far_, int return_depth, void* workspace, void* stream) {
  JAF_REQUIRE((faces || F == 0) && face_index_map && weight_map && depth_map && workspace, "null pointer");
  JAF_REQUIRE(!return_depth || face_inv_map, "return_depth needs face_inv_map [B,S,S,3,3]");
  JAF_REQUIRE(check_raster_args(B, F, image_size), "bad sizes");
  if (B == 0) return JAF_OK;
  cudaStream_t st = jaf::as_stream(stream);
  auto* zb = static_cast<unsigned long long*>(workspace);
  const long npix = (long)B * image_size * image_size;
  int launches = 0;
  if (faces_inv != nullptr && (long)B * F > 0) {
    k_faces_inv<<<jaf::ceil_div((long)B * F, 256), 256, 0, st>>>(faces, (long)B * F, image_size, faces_inv);
    ++launches;
  }
  {
    const int st1 = run_pass1<false>(faces, nullptr, nullptr, nullptr, B, 0, F, image_size, 0.f, near_, far_, workspace,
                                     nullptr, st, &launches);
    if (st1 != JAF_OK) return st1;
  }
  k_raster_resolve<false, false><<<jaf::ceil_div(npix, 256), 256, 0, st>>>(
      zb, faces, nullptr, nullptr, nullptr, nullptr, nullptr, B, 0, F, image_size, 0.f, near_, far_, 0, face_index_map,
      weight_map, depth_map, nullptr, 0, 1, return_depth ? face_inv_map : nullptr);
  return jaf::finish_launch("jaf_forward_face_index_map", launches + 1);
}

int jaf_render_fim_wim(const float* cam, const float* verts, const int32_t* faces_idx, int B, int V, int F,
                       int image_size, float eye_z, float near_, float far_, float* faces_xyz, int32_t* fim,
                       float* wim, void* workspace, void* stream) {
  JAF_REQUIRE(cam && verts && faces_idx && fim && wim && workspace, "null pointer");
  JAF_REQUIRE(check_raster_args(B, F, image_size) && V > 0, "bad sizes");
  if (B == 0) return JAF_OK;
  cudaStream_t st = jaf::as_stream(stream);
  auto* zb = static_cast<unsigned long long*>(workspace);
  const long npix = (long)B * image_size * image_size;
  int launches = 0;
  {
    const int st1 = run_pass1<true>(nullptr, cam, verts, faces_idx, B, V, F, image_size, eye_z, near_, far_, workspace,
                                    faces_xyz, st, &launches);
    if (st1 != JAF_OK) return st1;
  }
  k_raster_resolve<true, false><<<jaf::ceil_div(npix, 256), 256, 0, st>>>(
      zb, nullptr, cam, verts, faces_idx, nullptr, nullptr, B, V, F, image_size, eye_z, near_, far_, 1, fim, wim,
      nullptr, nullptr, 0);
  return jaf::finish_launch("jaf_render_fim_wim", launches + 1);
}

int jaf_flow_compose(const float* src_pts, int stride, int negate_y, const int32_t* fim, const float* wim, int B,
                     int F, int H, int W, float* T, void* stream) {
  JAF_REQUIRE(src_pts && fim && wim && T, "null pointer");
  JAF_REQUIRE(stride == 2 || stride == 3, "stride must be 2 or 3");
  JAF_REQUIRE(B >= 0 && F > 0 && H > 0 && W > 0, "bad sizes");
  JAF_REQUIRE((reinterpret_cast<uintptr_t>(T) & 7u) == 0, "T must be 8-byte aligned");
  if (B == 0) return JAF_OK;
  const long n = (long)B * H * W;
  k_flow_compose<<<jaf::ceil_div(n, 256), 256, 0, jaf::as_stream(stream)>>>(src_pts, stride, negate_y, fim, wim, B,
                                                                           F, (long)H * W, T);
  return jaf::finish_launch("k_flow_compose");
}

int jaf_cal_flow_multi(const float* src_cam, const float* src_verts, const float* tgt_cam, const float* tgt_verts,
                       const int32_t* faces_idx, int B, int K, int V, int F, int image_size, float eye_z, float near_,
                       float far_, float* T, int32_t* fim, float* wim, void* workspace, void* stream) {
  JAF_REQUIRE(src_cam && src_verts && tgt_cam && tgt_verts && faces_idx && T && workspace, "null pointer");
  JAF_REQUIRE(check_raster_args(B, F, image_size) && V > 0 && K >= 1, "bad sizes");
  JAF_REQUIRE((reinterpret_cast<uintptr_t>(T) & 7u) == 0, "T must be 8-byte aligned");
  if (B == 0) return JAF_OK;
  cudaStream_t st = jaf::as_stream(stream);
  auto* zb = static_cast<unsigned long long*>(workspace);
  const long npix = (long)B * image_size * image_size;
  int launches = 0;
  {
    const int st1 = run_pass1<true>(nullptr, tgt_cam, tgt_verts, faces_idx, B, V, F, image_size, eye_z, near_, far_,
                                    workspace, nullptr, st, &launches);
    if (st1 != JAF_OK) return st1;
  }
  k_raster_resolve<true, true><<<jaf::ceil_div(npix, 256), 256, 0, st>>>(
      zb, nullptr, tgt_cam, tgt_verts, faces_idx, src_cam, src_verts, B, V, F, image_size, eye_z, near_, far_, 1,
      fim, wim, nullptr, T, K);
  return jaf::finish_launch("jaf_cal_flow", launches + 1);
}

int jaf_cal_flow(const float* src_cam, const float* src_verts, const float* tgt_cam, const float* tgt_verts,
                 const int32_t* faces_idx, int B, int V, int F, int image_size, float eye_z, float near_,
                 float far_, float* T, int32_t* fim, float* wim, void* workspace, void* stream) {
  return jaf_cal_flow_multi(src_cam, src_verts, tgt_cam, tgt_verts, faces_idx, B, 1, V, F, image_size, eye_z, near_,
                            far_, T, fim, wim, workspace, stream);
}

}  // extern "C"
