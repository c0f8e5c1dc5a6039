// Pinned fp32 arithmetic of the rasteriser (rows a1-a6), shared by raster.cu and the pose-driven flavour of the fused
// warp kernel (warp_fuse.cu), which recomputes the winning face's barycentric weights per tile instead of reading them
// from HBM.  Every float operation is pinned with round-to-nearest intrinsics in the exact order — including the FMA
// contractions — that nvcc emits for the reference source on sm_100 (DESIGN.md "Raster arithmetic").
#pragma once
#include <cuda_runtime.h>

namespace jaf_raster {

constexpr unsigned long long kEmptyKey = ~0ull;  // z-buffer key of a pixel no face covers

// ---- a1-a3: src/nmr.py:10-28 (s*(X+t)), :271 (y *= -1), NR/look_at.py:59 (v - eye; the rotation
// is exactly the identity for SMPLRenderer's eye), NR/vertices_to_faces.py:19-22 (gather).
__device__ __forceinline__ void project_vertex(const float* __restrict__ p, float s, float tx, float ty,
                                               float eye_z, float* o) {
  o[0] = __fmul_rn(s, __fadd_rn(p[0], tx));
  o[1] = -__fmul_rn(s, __fadd_rn(p[1], ty));
  o[2] = __fsub_rn(p[2], eye_z);
}

__device__ __forceinline__ void load_face_projected(const float* __restrict__ cam, const float* __restrict__ verts,
                                                    const int* __restrict__ fidx, int b, int fn, int V,
                                                    float eye_z, float* f) {
  const float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int v = fidx[fn * 3 + k];
    project_vertex(verts + ((size_t)b * V + v) * 3, s, tx, ty, eye_z, f + 3 * k);
  }
}

// ---- rasterize_cuda_kernel.cu:40 / :111
__device__ __forceinline__ bool face_is_back(const float* f) {
  return __fmul_rn(__fsub_rn(f[7], f[1]), __fsub_rn(f[3], f[0])) <
         __fmul_rn(__fsub_rn(f[4], f[1]), __fsub_rn(f[6], f[0]));
}

// ---- rasterize_cuda_kernel.cu:44-62.  Returns the (pixel-space) denominator; also reports the
// pixel-space vertex positions for the bounding box.
__device__ __forceinline__ float face_setup(const float* f, int is, float* inv, float* px, float* py) {
  const float isf = (float)is;
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    px[n] = __fmul_rn(__fadd_rn(__fmaf_rn(f[3 * n + 0], isf, isf), -1.0f), 0.5f);
    py[n] = __fmul_rn(__fadd_rn(__fmaf_rn(f[3 * n + 1], isf, isf), -1.0f), 0.5f);
  }
  inv[0] = __fsub_rn(py[1], py[2]);
  inv[1] = __fsub_rn(px[2], px[1]);
  inv[2] = __fmaf_rn(px[1], py[2], -__fmul_rn(px[2], py[1]));
  inv[3] = __fsub_rn(py[2], py[0]);
  inv[4] = __fsub_rn(px[0], px[2]);
  inv[5] = __fmaf_rn(px[2], py[0], -__fmul_rn(px[0], py[2]));
  inv[6] = __fsub_rn(py[0], py[1]);
  inv[7] = __fsub_rn(px[1], px[0]);
  inv[8] = __fmaf_rn(px[0], py[1], -__fmul_rn(px[1], py[0]));
  const float den = __fmaf_rn(px[1], inv[3], __fmaf_rn(px[2], inv[6], __fmul_rn(px[0], inv[0])));
#pragma unroll
  for (int k = 0; k < 9; ++k) inv[k] = __fdiv_rn(inv[k], den);
  return den;
}

// ---- rasterize_cuda_kernel.cu:96-97, :115-137.  true => the face is a z-buffer candidate at (xi, yi).
__device__ __forceinline__ bool pixel_test(const float* f, const float* inv, int xi, int yi, int is,
                                           float near_, float far_, float* w, float* zp_out) {
  // :96-97 `(2. * yi + 1 - is) / is` in double, rounded to float.  For is <= 4096 the numerator is a small
  // integer and a correctly rounded fp32 division gives the same bits (no double-rounding case exists for
  // quotients of integers below 2^13; checked exhaustively in tools/), without the fp64 divide.
  float yp, xp;
  if (is <= 4096) {
    yp = __fdiv_rn((float)(2 * yi + 1 - is), (float)is);
    xp = __fdiv_rn((float)(2 * xi + 1 - is), (float)is);
  } else {
    yp = (float)((2. * yi + 1 - is) / is);
    xp = (float)((2. * xi + 1 - is) / is);
  }
  if (__fmul_rn(__fsub_rn(yp, f[1]), __fsub_rn(f[3], f[0])) < __fmul_rn(__fsub_rn(xp, f[0]), __fsub_rn(f[4], f[1])))
    return false;
  if (__fmul_rn(__fsub_rn(yp, f[4]), __fsub_rn(f[6], f[3])) < __fmul_rn(__fsub_rn(xp, f[3]), __fsub_rn(f[7], f[4])))
    return false;
  if (__fmul_rn(__fsub_rn(yp, f[7]), __fsub_rn(f[0], f[6])) < __fmul_rn(__fsub_rn(xp, f[6]), __fsub_rn(f[1], f[7])))
    return false;
  const float xif = (float)xi, yif = (float)yi;
  float w_sum = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float wk = __fadd_rn(__fmaf_rn(xif, inv[3 * k + 0], __fmul_rn(yif, inv[3 * k + 1])), inv[3 * k + 2]);
    wk = (float)fmin(fmax((double)wk, 0.), 1.);  // :129, double min/max (NaN -> 0)
    w[k] = wk;
    w_sum = __fadd_rn(w_sum, wk);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = __fdiv_rn(w[k], w_sum);
  const float s = __fadd_rn(__fadd_rn(__fdiv_rn(w[0], f[2]), __fdiv_rn(w[1], f[5])), __fdiv_rn(w[2], f[8]));
  const float zp = __frcp_rn(s);  // :136 `1. / s`: double reciprocal of a float == correctly rounded fp32
  *zp_out = zp;
  // :137 skip if zp <= near || far <= zp; :142 accept only if zp < depth_min (<= far).  NaN fails.
  return (zp > near_) && (zp < far_);
}

}  // namespace jaf_raster
