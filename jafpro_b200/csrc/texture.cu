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

}  // namespace

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
