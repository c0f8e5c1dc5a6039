// IUV texture lookup: test/conv_pro_test.py:41-74 / train/4.convLSTM_flowpro_interval.py:43-76.
//
// The reference loops over the 24 body parts and, for each, builds a full-frame grid, runs
// F.grid_sample(part_texture, grid, 'bilinear') (zero padding) and merges with torch.where: ~120 launches
// and 24 full-frame passes per target frame.  Here one thread per pixel reads its IUV triple (3 bytes),
// picks its own part and gathers the 2x2 taps of the 3 colour planes: every input byte is touched once.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_texture_warp(const float* __restrict__ tex, int P, int Ht, int Wt, const unsigned char* __restrict__ iuv, int B,
               long HW, int align_corners, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * HW) return;
  const int b = (int)(i / HW);
  const long p = i % HW;
  const int part = iuv[i * 3 + 0];
  float r[3] = {0.f, 0.f, 0.f};
  if (part >= 1 && part <= P) {
    const float u = (float)iuv[i * 3 + 1], v = (float)iuv[i * 3 + 2];
    // :63-64  x = ((255 - V)/255. - 0.5)*2 ; y = (U/255. - 0.5)*2   (separately rounded fp32 ops)
    const float gx = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(255.0f, v), 255.0f), 0.5f), 2.0f);
    const float gy = __fmul_rn(__fsub_rn(__fdiv_rn(u, 255.0f), 0.5f), 2.0f);
    // ATen grid_sampler_2d, bilinear, padding_mode='zeros': no clamp, out-of-range taps contribute nothing
    const float ix = align_corners ? __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(Wt - 1))
                                   : __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), (float)Wt, -1.f), 0.5f);
    const float iy = align_corners ? __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(Ht - 1))
                                   : __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), (float)Ht, -1.f), 0.5f);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float ax = __fsub_rn(fx + 1.f, ix), bx = __fsub_rn(ix, fx), ay = __fsub_rn(fy + 1.f, iy), by = __fsub_rn(iy, fy);
    const float nw = __fmul_rn(ax, ay), ne = __fmul_rn(bx, ay), sw = __fmul_rn(ax, by), se = __fmul_rn(bx, by);
    const bool inx0 = x0 >= 0 && x0 < Wt, inx1 = x1 >= 0 && x1 < Wt, iny0 = y0 >= 0 && y0 < Ht, iny1 = y1 >= 0 && y1 < Ht;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = tex + ((long)(part - 1) * 3 + c) * Ht * Wt;
      float acc = 0.f;
      if (inx0 && iny0) acc = fmaf(__ldg(pl + (long)y0 * Wt + x0), nw, acc);
      if (inx1 && iny0) acc = fmaf(__ldg(pl + (long)y0 * Wt + x1), ne, acc);
      if (inx0 && iny1) acc = fmaf(__ldg(pl + (long)y1 * Wt + x0), sw, acc);
      if (inx1 && iny1) acc = fmaf(__ldg(pl + (long)y1 * Wt + x1), se, acc);
      r[c] = acc;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) st_stream_f32(out + ((long)b * 3 + c) * HW + p, r[c]);
}

// ---------------------------------------------------------------------------------------------------
// texture-space assembly around Accumulate_LSTM_no_loss (test/conv_pro_test.py:209-239;
// train/4.convLSTM_flowpro_interval.py:269-298).  The atlas is rows x cols parts of ph x pw pixels (4 x 6 x 200 x 200).
//   gather : src_texture_im[:, ref[z], :, i*ph:(i+1)*ph, j*pw:(j+1)*pw] for every part and selected reference,
//            already in the [K*B, C, ph, pw] order Downsampler_convLSTM's torch.cat(x, dim=0) builds
//            (src/networks.py:1316) -> out [parts, K, B, C, ph, pw]                       (:209-217)
//   mask   : common = OR_z uint8(src_mask_im[:, ref[z]]) as float, parts[p] *= common     (:221-236)
//   scatter: texture_image[:, :, i*ph:.., j*pw:..] = parts[i*cols+j]                       (src/networks.py:1685-1691)
// One thread per output element, x fastest: reads and writes are coalesced part-row segments.
// ---------------------------------------------------------------------------------------------------
struct AtlasGeom {
  int B, Kmax, K, C, rows, cols, ph, pw;
};

__global__ void __launch_bounds__(256)
k_parts_gather(const float* __restrict__ atlas, const int* __restrict__ ref, AtlasGeom g, long n, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long r = i;
  const int x = (int)(r % g.pw); r /= g.pw;
  const int y = (int)(r % g.ph); r /= g.ph;
  const int c = (int)(r % g.C); r /= g.C;
  const int b = (int)(r % g.B); r /= g.B;
  const int z = (int)(r % g.K);
  const int part = (int)(r / g.K);
  const int pi = part / g.cols, pj = part % g.cols;
  const long AW = (long)g.cols * g.pw, AH = (long)g.rows * g.ph;
  const long src = ((((long)b * g.Kmax + __ldg(ref + z)) * g.C + c) * AH + (long)pi * g.ph + y) * AW + (long)pj * g.pw + x;
  st_stream_f32(out + i, ld_stream_f32(atlas + src));
}

__global__ void __launch_bounds__(256)
k_parts_common_mask(float* __restrict__ parts, const float* __restrict__ mask, const int* __restrict__ ref, AtlasGeom g,
                    long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // over [parts, B, ph, pw]: one mask value serves C channels
  if (i >= n) return;
  long r = i;
  const int x = (int)(r % g.pw); r /= g.pw;
  const int y = (int)(r % g.ph); r /= g.ph;
  const int b = (int)(r % g.B);
  const int part = (int)(r / g.B);
  const int pi = part / g.cols, pj = part % g.cols;
  const long AW = (long)g.cols * g.pw, AH = (long)g.rows * g.ph;
  unsigned common = 0;
  for (int z = 0; z < g.K; ++z) {
    const float m = ld_stream_f32(mask + (((long)b * g.Kmax + __ldg(ref + z)) * AH + (long)pi * g.ph + y) * AW + (long)pj * g.pw + x);
    common |= (unsigned)(unsigned char)(int)m;  // .byte(): truncate toward zero, keep the low 8 bits
  }
  const float cf = (float)common;
  const long plane = (long)g.ph * g.pw;
  float* p = parts + (((long)part * g.B + b) * g.C) * plane + (long)y * g.pw + x;
  for (int c = 0; c < g.C; ++c) p[c * plane] *= cf;
}

__global__ void __launch_bounds__(256)
k_parts_scatter(const float* __restrict__ parts, AtlasGeom g, long n, float* __restrict__ atlas) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // over the atlas [B, C, AH, AW]
  if (i >= n) return;
  const long AW = (long)g.cols * g.pw, AH = (long)g.rows * g.ph;
  long r = i;
  const int ax = (int)(r % AW); r /= AW;
  const int ay = (int)(r % AH); r /= AH;
  const int c = (int)(r % g.C);
  const int b = (int)(r / g.C);
  const int part = (ay / g.ph) * g.cols + ax / g.pw;
  const long src = ((((long)part * g.B + b) * g.C + c) * g.ph + ay % g.ph) * g.pw + ax % g.pw;
  st_stream_f32(atlas + i, ld_stream_f32(parts + src));
}

// ---------------------------------------------------------------------------------------------------
// per-frame IUV preprocessing of the reference's data loader (src/data.py:102-113,:504).
//   TransferTexture (src/utils.py:369-394): nearest texel lookup  U = rint(IUV[1]/255.*199.), V likewise,
//       out = tex[i*200 + U, j*200 + (199 - V)] for part 6*i + j + 1; channels that come out 0 take `im`.
//   compute_angle (src/computer_angle.py:4-39): needs the pixel count of every part and the sum of x of
//       parts 1 and 2 — a per-frame histogram; the scalar formula stays on the host in float64.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_transfer_texture(const unsigned char* __restrict__ tex, long tex_stride, int cols, int nparts, int ps,
                   const unsigned char* __restrict__ iuv, const unsigned char* __restrict__ im, long HW, long n,
                   unsigned char* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = (int)(i / HW);
  const int part = iuv[i * 3 + 0];
  unsigned char r[3] = {0, 0, 0};
  if (part >= 1 && part <= nparts) {
    // float64 like numpy: uint8 / 255. * 199. then round-half-even, then the uint8 cast
    const double last = (double)(ps - 1);
    const int u = (int)(unsigned char)rint((double)iuv[i * 3 + 1] / 255.0 * last);
    const int v = (int)(unsigned char)rint((double)iuv[i * 3 + 2] / 255.0 * last);
    const int ic = (part - 1) / cols, jc = part - ic * cols - 1;
    const long AW = (long)cols * ps;
    const unsigned char* t = tex + b * tex_stride + (((long)ic * ps + u) * AW + (long)jc * ps + (ps - 1 - v)) * 3;
    r[0] = __ldg(t);
    r[1] = __ldg(t + 1);
    r[2] = __ldg(t + 2);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    unsigned char o = r[c];
    if (im != nullptr && o == 0) o = im[i * 3 + c];  // BG_MASK = output_img == 0, per channel (:390-391)
    out[i * 3 + c] = o;
  }
}

__global__ void __launch_bounds__(256)
k_iuv_part_stats(const unsigned char* __restrict__ iuv, int W, long HW, int* __restrict__ counts,
                 long long* __restrict__ sumx) {
  __shared__ int s_cnt[32];
  __shared__ unsigned long long s_sx[32];
  if (threadIdx.x < 32) {
    s_cnt[threadIdx.x] = 0;
    s_sx[threadIdx.x] = 0ull;
  }
  __syncthreads();
  const int b = blockIdx.y;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long)gridDim.x * blockDim.x) {
    const int part = iuv[((long)b * HW + p) * 3];
    if (part >= 1 && part < 32) {
      atomicAdd(&s_cnt[part], 1);
      atomicAdd(&s_sx[part], (unsigned long long)(p % W));
    }
  }
  __syncthreads();
  if (threadIdx.x < 32 && s_cnt[threadIdx.x] != 0) {
    atomicAdd(&counts[b * 32 + threadIdx.x], s_cnt[threadIdx.x]);
    atomicAdd(reinterpret_cast<unsigned long long*>(&sumx[b * 32 + threadIdx.x]), s_sx[threadIdx.x]);
  }
}

bool geom_ok(const AtlasGeom& g) {
  return g.B > 0 && g.Kmax > 0 && g.K > 0 && g.C > 0 && g.rows > 0 && g.cols > 0 && g.ph > 0 && g.pw > 0;
}

}  // namespace

extern "C" int jaf_texture_parts_gather(const float* atlas, const int32_t* ref_index, int B, int Kmax, int K, int C,
                                        int rows, int cols, int ph, int pw, float* out, void* stream) {
  JAF_REQUIRE(atlas && ref_index && out, "null pointer");
  const AtlasGeom g{B, Kmax, K, C, rows, cols, ph, pw};
  JAF_REQUIRE(geom_ok(g), "bad sizes");
  const long n = (long)rows * cols * K * B * C * ph * pw;
  k_parts_gather<<<jaf::ceil_div(n, 256), 256, 0, jaf::as_stream(stream)>>>(atlas, ref_index, g, n, out);
  return jaf::finish_launch("k_parts_gather");
}

extern "C" int jaf_texture_parts_common_mask(float* parts, const float* mask, const int32_t* ref_index, int B, int Kmax,
                                             int K, int C, int rows, int cols, int ph, int pw, void* stream) {
  JAF_REQUIRE(parts && mask && ref_index, "null pointer");
  const AtlasGeom g{B, Kmax, K, C, rows, cols, ph, pw};
  JAF_REQUIRE(geom_ok(g), "bad sizes");
  const long n = (long)rows * cols * B * ph * pw;
  k_parts_common_mask<<<jaf::ceil_div(n, 256), 256, 0, jaf::as_stream(stream)>>>(parts, mask, ref_index, g, n);
  return jaf::finish_launch("k_parts_common_mask");
}

extern "C" int jaf_texture_parts_scatter(const float* parts, int B, int C, int rows, int cols, int ph, int pw,
                                         float* atlas, void* stream) {
  JAF_REQUIRE(parts && atlas, "null pointer");
  const AtlasGeom g{B, 1, 1, C, rows, cols, ph, pw};
  JAF_REQUIRE(geom_ok(g), "bad sizes");
  const long n = (long)B * C * rows * ph * cols * pw;
  k_parts_scatter<<<jaf::ceil_div(n, 256), 256, 0, jaf::as_stream(stream)>>>(parts, g, n, atlas);
  return jaf::finish_launch("k_parts_scatter");
}

extern "C" int jaf_transfer_texture(const uint8_t* tex, int tex_batched, int rows, int cols, int part_size,
                                    const uint8_t* iuv, const uint8_t* im, int B, int H, int W, uint8_t* out,
                                    void* stream) {
  JAF_REQUIRE(tex && iuv && out, "null pointer");
  JAF_REQUIRE(rows > 0 && cols > 0 && part_size > 0 && part_size <= 256 && rows * cols <= 255 && B > 0 && H > 0 && W > 0,
              "bad sizes");
  const long n = (long)B * H * W;
  const long tex_stride = tex_batched ? (long)rows * part_size * cols * part_size * 3 : 0;
  k_transfer_texture<<<jaf::ceil_div(n, 256), 256, 0, jaf::as_stream(stream)>>>(tex, tex_stride, cols, rows * cols,
                                                                               part_size, iuv, im, (long)H * W, n, out);
  return jaf::finish_launch("k_transfer_texture");
}

extern "C" int jaf_iuv_part_stats(const uint8_t* iuv, int B, int H, int W, int32_t* counts, int64_t* sumx, void* stream) {
  JAF_REQUIRE(iuv && counts && sumx, "null pointer");
  JAF_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "bad sizes");
  cudaStream_t st = jaf::as_stream(stream);
  JAF_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * 32 * sizeof(int32_t), st));
  JAF_CUDA(cudaMemsetAsync(sumx, 0, (size_t)B * 32 * sizeof(int64_t), st));
  const long HW = (long)H * W;
  int bx = jaf::ceil_div(HW, 256 * 8);
  if (bx > 64) bx = 64;
  k_iuv_part_stats<<<dim3(bx, B), 256, 0, st>>>(iuv, W, HW, counts, reinterpret_cast<long long*>(sumx));
  return jaf::finish_launch("k_iuv_part_stats");
}

// =====================================================================================
// DensePose texture extraction, get_texture (src/utils.py:232-255).
// Per part 1..24 the reference scatters the part's pixels into a tex_size x tex_size x 3 map
//   row = int((255 - V) * (tex_size-1) / 255.), col = int(U * (tex_size-1) / 255.)      (:248-249, float64 arithmetic)
// in np.where order (row-major pixels; numpy's fancy assignment keeps the LAST write, :250-251), resizes it to
// final_size with cv2.resize(..., INTER_LINEAR) (float64 data), reverses the channels and divides by 255 (:253).
// Here: (1) one thread per pixel elects the owner of its texel with atomicMax on the pixel's linear index (= the last
// writer in row-major order); (2) one thread per output texel evaluates the half-pixel-centre bilinear resize in fp64
// reading the four owners' pixels straight from the image (the small map is never materialised).
// cv::resize's coefficient rounding depends on the OpenCV build (generic float coefficients vs the IPP path of the pip
// wheel); ours are exact rationals rounded once to double: outputs agree to ~1e-14 (tests: <= 1e-12).
// =====================================================================================
__global__ void __launch_bounds__(256)
k_get_texture_scatter(const uint8_t* __restrict__ iuv, int B, long HW, int tex_size, unsigned int* __restrict__ owner) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * HW) return;
  const int part = iuv[i * 3 + 0];
  if (part < 1 || part > 24) return;
  const double sf = (double)tex_size - 1.0;
  const int u = iuv[i * 3 + 1], v = iuv[i * 3 + 2];
  const int row = (int)__ddiv_rn(__dmul_rn((double)(255 - v), sf), 255.0);
  const int col = (int)__ddiv_rn(__dmul_rn((double)u, sf), 255.0);
  const long b = i / HW;
  const unsigned int pix1 = (unsigned int)(i - b * HW) + 1u;  // 0 = empty texel
  atomicMax(owner + ((b * 24 + (part - 1)) * tex_size + row) * tex_size + col, pix1);
}

// offset and weight of output index d along one axis (sn source texels -> dn outputs)
__device__ __forceinline__ void resize_coeff(int d, int sn, int dn, int* s0, int* s1, double* w1) {
  double f = __ddiv_rn((double)((2L * d + 1) * sn - dn), (double)(2L * dn));
  int s = (int)floor(f);
  f = __dsub_rn(f, (double)s);
  if (s < 0) { s = 0; f = 0.0; }
  if (s >= sn - 1) { s = sn - 1; f = 0.0; }
  *s0 = s;
  *s1 = min(s + 1, sn - 1);
  *w1 = f;
}

__global__ void __launch_bounds__(256)
k_get_texture_resize(const uint8_t* __restrict__ im, const unsigned int* __restrict__ owner, int B, long HW, int tex_size,
                     int final_size, double* __restrict__ parts) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long per_part = (long)final_size * final_size;
  if (i >= (long)B * 24 * per_part) return;
  const long bp = i / per_part;           // b * 24 + part
  const int dy = (int)((i - bp * per_part) / final_size), dx = (int)(i % final_size);
  const long b = bp / 24;
  int x0, x1, y0, y1;
  double ax, ay;
  resize_coeff(dx, tex_size, final_size, &x0, &x1, &ax);
  resize_coeff(dy, tex_size, final_size, &y0, &y1, &ay);
  const unsigned int* o = owner + bp * tex_size * tex_size;
  const unsigned int t00 = o[y0 * tex_size + x0], t01 = o[y0 * tex_size + x1], t10 = o[y1 * tex_size + x0],
                     t11 = o[y1 * tex_size + x1];
  const uint8_t* imb = im + b * HW * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v00 = t00 ? (double)imb[(long)(t00 - 1) * 3 + c] : 0.0, v01 = t01 ? (double)imb[(long)(t01 - 1) * 3 + c] : 0.0;
    const double v10 = t10 ? (double)imb[(long)(t10 - 1) * 3 + c] : 0.0, v11 = t11 ? (double)imb[(long)(t11 - 1) * 3 + c] : 0.0;
    const double top = __dadd_rn(__dmul_rn(v00, __dsub_rn(1.0, ax)), __dmul_rn(v01, ax));
    const double bot = __dadd_rn(__dmul_rn(v10, __dsub_rn(1.0, ax)), __dmul_rn(v11, ax));
    const double r = __dadd_rn(__dmul_rn(top, __dsub_rn(1.0, ay)), __dmul_rn(bot, ay));
    parts[i * 3 + (2 - c)] = __ddiv_rn(r, 255.0);  // [:, :, ::-1] / 255.
  }
}

extern "C" size_t jaf_get_texture_workspace_bytes(int B, int tex_size) {
  if (B <= 0 || tex_size <= 0) return 0;
  return (size_t)B * 24 * tex_size * tex_size * sizeof(unsigned int);
}

extern "C" int jaf_get_texture(const uint8_t* im, const uint8_t* iuv, int B, int H, int W, int tex_size, int final_size,
                               double* parts, void* workspace, void* stream) {
  JAF_REQUIRE(im && iuv && parts && workspace, "null pointer");
  JAF_REQUIRE(B > 0 && H > 0 && W > 0 && tex_size >= 2 && tex_size <= 4096 && final_size >= 1 && final_size <= 8192, "bad sizes");
  JAF_REQUIRE((long)H * W < 0xffffffffL, "image too large");
  cudaStream_t st = jaf::as_stream(stream);
  JAF_CUDA(cudaMemsetAsync(workspace, 0, jaf_get_texture_workspace_bytes(B, tex_size), st));
  const long n = (long)B * H * W;
  k_get_texture_scatter<<<jaf::ceil_div(n, 256), 256, 0, st>>>(iuv, B, (long)H * W, tex_size, static_cast<unsigned int*>(workspace));
  const long m = (long)B * 24 * final_size * final_size;
  k_get_texture_resize<<<jaf::ceil_div(m, 256), 256, 0, st>>>(im, static_cast<const unsigned int*>(workspace), B, (long)H * W,
                                                            tex_size, final_size, parts);
  return jaf::finish_launch("jaf_get_texture", 2);
}

extern "C" int jaf_texture_warp(const float* tex_parts, int P, int Ht, int Wt, const uint8_t* iuv, int B, int H, int W,
                                int align_corners, float* out, void* stream) {
  JAF_REQUIRE(tex_parts && iuv && out, "null pointer");
  JAF_REQUIRE(P >= 1 && P <= 255 && Ht > 0 && Wt > 0 && B >= 0 && H > 0 && W > 0, "bad sizes");
  if (B == 0) return JAF_OK;
  const long n = (long)B * H * W;
  k_texture_warp<<<jaf::ceil_div(n, 256), 256, 0, jaf::as_stream(stream)>>>(tex_parts, P, Ht, Wt, iuv, B, (long)H * W,
                                                                           align_corners, out);
  return jaf::finish_launch("k_texture_warp");
}
