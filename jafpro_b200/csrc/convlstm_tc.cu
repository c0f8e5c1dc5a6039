// ConvLSTM cell step as a tcgen05 implicit GEMM with the gate epilogue fused (row a13, wide cells;
// src/convLSTM.py:41-56).  BASELINE config 4: B=16, Cin=Ch=256, 64x64, 3x3 -> M = B*H*W = 65,536 pixels,
// N = 4*Ch = 1024 gate channels, K = 9*(Cin+Ch) = 4,608; 6.18e11 FLOP per step.
//
//   D[pixel, gate-channel] = sum over (tap, channel) A[pixel shifted by tap, channel] * Wp[gate-channel, tap, channel]
//
// * A is never materialised (no im2col, no cat(x,h), :43): for every (tap, 64-channel chunk) one TMA
//   tile load pulls the 128-pixel window shifted by (ky-1, kx-1) straight out of the NHWC activation
//   (x for channels < Cin, h above), and TMA's out-of-bounds zero fill IS the conv padding (:38).
// * Weights are repacked once (jaf_convlstm_pack_weight): bf16, K-major [N][9*(Cin+Ch)], rows reordered
//   so that one 256-wide N tile holds i,f,o,g of the same 64 hidden channels — the epilogue therefore
//   has all four gates of a hidden channel in one accumulator tile and the 4*Ch pre-activation tensor
//   (:45-46) never reaches HBM.
// * 128x256 fp32 accumulators live in TMEM (two buffers = all 512 columns), so the epilogue of tile t
//   (tcgen05.ld -> bias, sigmoid/tanh, c' = f*c + i*g, h' = o*tanh(c'), :48-54, fp32 state) overlaps
//   the MMAs of tile t+1.
// * Persistent: one CTA per SM, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one
//   elected thread issues tcgen05.mma), warps 2..5 = epilogue (one TMEM lane quarter each).
//   4-stage smem ring (A 16 KB + B 32 KB per stage), mbarrier full/empty pipeline.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;      // pixels per tile  (UMMA M)
constexpr int BN = 256;      // gate channels per tile (UMMA N) = 4 gates x 64 hidden channels
constexpr int BK = 64;       // channels per k-block (128 bytes of bf16 = one SWIZZLE_128B row)
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;

using namespace jaf::tc;

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version = 1 at [46,48),
// layout type SWIZZLE_128B = 2 at [61,64).  Rows are 128 B, 8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = F32 (bit 4), A = B = BF16 (bits 7, 10), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
struct TcArgs {
  const float* c;
  const float* bias;
  __nv_bfloat16* h_out;
  float* c_out;
  int B, Cin, Ch, H, W;
  int Wt, Ht;         // pixel tile = Ht rows x Wt columns (Wt * Ht == 128)
  int tiles_x, tiles_y;
  int num_m_tiles, num_n_blocks, num_tiles;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_convlstm_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_h,
              const __grid_constant__ CUtensorMap map_w, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]    epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ct = a.Cin + a.Ch;
  const int kchunks = Ct / BK;
  const int num_kb = 9 * kchunks;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&map_x);
    tmap_prefetch(&map_h);
    tmap_prefetch(&map_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is a warp-wide operation; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int nb = tile % a.num_n_blocks, mt = tile / a.num_n_blocks;
        const int per_img = a.tiles_x * a.tiles_y;
        const int b = mt / per_img, rem = mt % per_img;
        const int y0 = (rem / a.tiles_x) * a.Ht, x0 = (rem % a.tiles_x) * a.Wt;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / kchunks, kc = kb % kchunks;
          const int ky = tap / 3, kx = tap % 3;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int ch0 = kc * BK;
          if (ch0 < a.Cin)
            tma_load_4d(sa, &map_x, &full_bar[stage], ch0, x0 + kx - 1, y0 + ky - 1, b);
          else
            tma_load_4d(sa, &map_h, &full_bar[stage], ch0 - a.Cin, x0 + kx - 1, y0 + ky - 1, b);
          tma_load_2d(sb, &map_w, &full_bar[stage], tap * Ct + ch0, nb * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = umma_desc(sa + k * UMMA_K * 2);
            const uint64_t bd = umma_desc(sb + k * UMMA_K * 2);
            umma_f16(tmem_d, ad, bd, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;      // accumulator row == pixel within the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int nb = tile % a.num_n_blocks, mt = tile / a.num_n_blocks;
      const int per_img = a.tiles_x * a.tiles_y;
      const int b = mt / per_img, rem = mt % per_img;
      const int y = (rem / a.tiles_x) * a.Ht + row / a.Wt, x = (rem % a.tiles_x) * a.Wt + row % a.Wt;
      const size_t pix = ((size_t)b * a.H + y) * a.W + x;
      const float* cp = a.c + pix * a.Ch + nb * 64;
      float* cop = a.c_out + pix * a.Ch + nb * 64;
      __nv_bfloat16* hop = a.h_out + pix * a.Ch + nb * 64;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int j0 = 0; j0 < 64; j0 += 16) {
        float gi[16], gf[16], go[16], gg[16];
        tmem_ld16(taddr + 0 * 64 + j0, gi);
        tmem_ld16(taddr + 1 * 64 + j0, gf);
        tmem_ld16(taddr + 2 * 64 + j0, go);
        tmem_ld16(taddr + 3 * 64 + j0, gg);
        float cv[16];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(cp + j0) + v);
          cv[4 * v + 0] = t4.x; cv[4 * v + 1] = t4.y; cv[4 * v + 2] = t4.z; cv[4 * v + 3] = t4.w;
        }
        tmem_ld_wait();
        float cn[16], hn[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int ch = nb * 64 + j0 + i;
          float bi = 0.f, bf = 0.f, bo = 0.f, bg = 0.f;
          if (a.bias) {
            bi = __ldg(a.bias + 0 * a.Ch + ch);
            bf = __ldg(a.bias + 1 * a.Ch + ch);
            bo = __ldg(a.bias + 2 * a.Ch + ch);
            bg = __ldg(a.bias + 3 * a.Ch + ch);
          }
          const float ig = sigmoid_f(gi[i] + bi), fg = sigmoid_f(gf[i] + bf), og = sigmoid_f(go[i] + bo);
          const float g_ = tanh_f(gg[i] + bg);
          cn[i] = fg * cv[i] + ig * g_;   // src/convLSTM.py:53
          hn[i] = og * tanh_f(cn[i]);     // :54
        }
#pragma unroll
        for (int v = 0; v < 4; ++v)
          __stcs(reinterpret_cast<float4*>(cop + j0) + v, make_float4(cn[4 * v], cn[4 * v + 1], cn[4 * v + 2], cn[4 * v + 3]));
        uint4 h0, h1;
        h0.x = pack_bf16x2(hn[0], hn[1]);   h0.y = pack_bf16x2(hn[2], hn[3]);
        h0.z = pack_bf16x2(hn[4], hn[5]);   h0.w = pack_bf16x2(hn[6], hn[7]);
        h1.x = pack_bf16x2(hn[8], hn[9]);   h1.y = pack_bf16x2(hn[10], hn[11]);
        h1.z = pack_bf16x2(hn[12], hn[13]); h1.w = pack_bf16x2(hn[14], hn[15]);
        __stcs(reinterpret_cast<uint4*>(hop + j0), h0);
        __stcs(reinterpret_cast<uint4*>(hop + j0) + 1, h1);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);  // 128 arrivals release the accumulator to the MMA warp
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// weight [4Ch][Ct][3][3] f32  ->  wpack [N' = 4Ch][9*Ct] bf16 with N' = blk*256 + gate*64 + jj  (orig row gate*Ch + blk*64 + jj)
// and K index = tap*Ct + c
__global__ void __launch_bounds__(256)
k_pack_weight(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Ch, int Ct) {
  const size_t n = (size_t)4 * Ch * 9 * Ct;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int kidx = (int)(i % ((size_t)9 * Ct));
  const int np = (int)(i / ((size_t)9 * Ct));
  const int tap = kidx / Ct, c = kidx % Ct;
  const int blk = np / 256, gate = (np % 256) / 64, jj = np % 64;
  const int orow = gate * Ch + blk * 64 + jj;
  wp[i] = __float2bfloat16_rn(w[((size_t)orow * Ct + c) * 9 + tap]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_act_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int Wt, int Ht) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    jaf::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return JAF_ERR_CUDA;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)Wt, (cuuint32_t)Ht, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    jaf::set_error("cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
    return JAF_ERR_CUDA;
  }
  return JAF_OK;
}

int make_w_map(CUtensorMap* m, const void* ptr, int N, int Ktot) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    jaf::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return JAF_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    jaf::set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    return JAF_ERR_CUDA;
  }
  return JAF_OK;
}

}  // namespace

extern "C" {

size_t jaf_convlstm_wpack_bytes(int Cin, int Ch) {
  if (Cin <= 0 || Ch <= 0) return 0;
  return (size_t)9 * 4 * Ch * (Cin + Ch) * 2;
}

int jaf_convlstm_pack_weight(const float* weight, int Cin, int Ch, void* wpack, void* stream) {
  JAF_REQUIRE(weight && wpack, "null pointer");
  JAF_REQUIRE(Cin > 0 && Ch > 0 && Cin % 64 == 0 && Ch % 64 == 0, "Cin and Ch must be positive multiples of 64");
  const size_t n = (size_t)4 * Ch * 9 * (Cin + Ch);
  k_pack_weight<<<jaf::ceil_div((long)n, 256), 256, 0, jaf::as_stream(stream)>>>(
      weight, static_cast<__nv_bfloat16*>(wpack), Ch, Cin + Ch);
  return jaf::finish_launch("k_pack_weight");
}

int jaf_convlstm_step_tc(const void* x, const void* h, const float* c, const void* wpack, const float* bias, int B,
                         int Cin, int Ch, int H, int W, void* h_out, float* c_out, void* stream) {
  JAF_REQUIRE(x && h && c && wpack && h_out && c_out, "null pointer");
  JAF_REQUIRE(B > 0 && H > 0 && W > 0, "bad sizes");
  JAF_REQUIRE(Cin > 0 && Ch > 0 && Cin % 64 == 0 && Ch % 64 == 0, "Cin and Ch must be positive multiples of 64");
  int Wt, Ht;
  if (W >= BM) {
    JAF_REQUIRE(W % BM == 0, "W must be a multiple of 128 when W >= 128");
    Wt = BM;
    Ht = 1;
  } else {
    JAF_REQUIRE(BM % W == 0, "W must divide 128 when W < 128");
    Wt = W;
    Ht = BM / W;
    JAF_REQUIRE(H % Ht == 0, "H must be a multiple of 128 / W");
  }
  JAF_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)h & 15) == 0 && ((uintptr_t)wpack & 15) == 0 &&
                  ((uintptr_t)c & 15) == 0 && ((uintptr_t)c_out & 15) == 0 && ((uintptr_t)h_out & 15) == 0,
              "tensors must be 16-byte aligned");
  CUtensorMap mx, mh, mw;
  int st = make_act_map(&mx, x, B, H, W, Cin, Wt, Ht);
  if (st != JAF_OK) return st;
  st = make_act_map(&mh, h, B, H, W, Ch, Wt, Ht);
  if (st != JAF_OK) return st;
  st = make_w_map(&mw, wpack, 4 * Ch, 9 * (Cin + Ch));
  if (st != JAF_OK) return st;

  TcArgs a;
  a.c = c; a.bias = bias; a.h_out = static_cast<__nv_bfloat16*>(h_out); a.c_out = c_out;
  a.B = B; a.Cin = Cin; a.Ch = Ch; a.H = H; a.W = W; a.Wt = Wt; a.Ht = Ht;
  a.tiles_x = W / Wt; a.tiles_y = H / Ht;
  a.num_m_tiles = B * a.tiles_x * a.tiles_y;
  a.num_n_blocks = (4 * Ch) / BN;
  a.num_tiles = a.num_m_tiles * a.num_n_blocks;

  static jaf::PerDeviceOnce attr_once;  // the shared-memory attribute is per device
  const int dev = jaf::current_device();
  const int sm_count = jaf::sm_count(dev);
  JAF_REQUIRE(sm_count > 0, "no current CUDA device");
  if (!attr_once.done(dev)) {
    JAF_CUDA(cudaFuncSetAttribute(k_convlstm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_once.mark(dev);
  }
  const int grid = a.num_tiles < sm_count ? a.num_tiles : sm_count;
  k_convlstm_tc<<<grid, NUM_THREADS, SMEM_BYTES, jaf::as_stream(stream)>>>(mx, mh, mw, a);
  return jaf::finish_launch("k_convlstm_tc");
}

}  // extern "C"
