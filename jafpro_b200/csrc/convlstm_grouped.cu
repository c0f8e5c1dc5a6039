// Reference-sized ConvLSTM cells on the tensor cores, G independent cells per launch (row a13 at the shapes
// the reference really runs: Accumulate_LSTM_no_loss builds 24 part-specific Downsampler_convLSTM stacks with
// cells of 12/24/24/48/96 channels on 200^2 ... 13^2 maps, src/networks.py:1290-1355,:1641-1662; one cell
// step is src/convLSTM.py:41-56).  fp32 in, fp32 out, reference (NCHW) layout, fp32-grade accuracy:
//
// * Split-bf16 arithmetic: every fp32 operand is written as hi + lo (two bf16 numbers, 16 mantissa bits
//   together) and the product is accumulated as hi*hi + hi*lo + lo*hi in fp32 TMEM accumulators — three
//   tcgen05.mma kind::f16 per k-step instead of one, relative error per product ~2^-17.
// * Implicit GEMM without im2col by flattening the ZERO-PADDED image: with p = (y+1)*(W+2) + (x+1) a 3x3 tap
//   is the constant row shift (ky-1)*(W+2) + (kx-1), so ONE shared-memory copy of the rows
//   [p0 - (W+3), p0 + 128*MT + (W+3)) x channels serves all nine taps: the A descriptor of a tap is the same
//   buffer with a different start address.  That needs rows at a uniform 16-byte pitch, i.e. the
//   no-swizzle K-major canonical layout [channel chunk of 8][row][16 B] with SBO = 128 B and
//   LBO = rows * 16 B.  Outputs computed for the two padding columns of each row are discarded.
//   The batch is flattened into the same index (padded images stacked), so tiles may straddle images.  Wide maps are
//   cut into vertical bands of ~25 columns, each flattened on its own with its neighbours' columns as halo: the row
//   window of a tile is then 128*MT + 2*(band width + 3) rows instead of 128*MT + 2*(W + 3).
// * The fp32 NCHW activations (x for channels < Cin, h above: the cat of :43 never exists) are read
//   coalesced along x, split to hi/lo bf16 in registers and stored with conflict-free 16-byte stores.
// * Weights are repacked once per model (jaf_convlstm_gpack_weight) into the exact shared-memory image of
//   the B operand, hi and lo, in k-step order; stages of KS k-steps stream through a 3-slot ring with
//   cp.async.bulk + mbarrier while the MMAs of the previous stage run.
// * MT accumulator tiles (128 pixels x 4*Ch gate channels each) live in TMEM at once; the epilogue reads
//   them with tcgen05.ld (thread = pixel, so global accesses are coalesced in NCHW), adds the bias and applies
//   the gates (:46-54) without the 4*Ch pre-activation tensor ever reaching HBM.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace jaf::tc;

constexpr int kThreads = 1024;  // warp 0: weight stream + MMA issue; the other warps: staging + epilogue
constexpr int kThreadsHalf = 512;  // two CTAs per SM (64 registers per thread either way)
constexpr int kHalfSmem = 112 * 1024;
constexpr int kThreadsQuarter = 256;  // four CTAs per SM: more staging / MMA / epilogue phases of different tiles overlap
constexpr int kQuarterSmem = 55 * 1024;
constexpr int kRing = 3;        // weight stages in flight the plan guarantees (it then deepens the ring into the free shared memory)
constexpr int kMaxRing = 16;
constexpr int kMaxSmem = 232448 - 1024;

struct FastDiv {
  uint32_t d, m;
};

struct GArgs {
  const float* x;
  const float* h;
  const float* c;
  const float* bias;
  const uint8_t* wpack;
  float* h_out;
  float* c_out;
  int G, B, Cin, Ch, H, W;
  // elements between consecutive (cell, batch) items of x, h and h_out: Cin*H*W / Ch*H*W for plain [G,B,C,H,W]
  // tensors, T times that for the time-step slices of a [G,B,T,C,H,W] sequence (jaf_convlstm_sequence_grouped)
  long xs, hs, hos;
  int nb, Wb;        // vertical bands per image and their width: each band is flattened on its own (short halo)
  int Wp, HpWp;      // padded row pitch of a band (Wb + 2), padded band size (H + 2) * Wp
  long Q;            // B * nb * HpWp flattened padded positions per group
  int Ct, Ctp, N, nsplit, Nsub;
  int CS, Chs, Ns;   // hidden-channel slices per cell, channels per slice, gate rows per slice (= 4 * Chs)
  int MT, R;         // accumulator tiles per CTA, staged rows
  int S, KS, nstages;
  int tiles_per_group;
  FastDiv d_hpwp, d_wp, d_nb;  // persistent kernel: position -> (image, band, row, column) without divisions
  FastDiv d_tpg;               // persistent kernel: tile -> cell (ntiles < 2^31)
  int C8, U;         // persistent kernel: 8-channel chunks of the staged window, K units (9 * C8)
  int ring;          // weight stages (k_convlstm_grouped) / k-steps (persistent kernel) in flight
  long ntiles;       // persistent kernel: G * tiles_per_group
  uint32_t idesc, a_half, stage_bytes, tmem_cols;
};

// no-swizzle K-major matrix descriptor: rows at a 16-byte pitch (SBO = 128 B per 8 rows), the two 8-element
// K chunks of one k-step `lbo` bytes apart
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Issued from warp-uniform code under elect_one(): ptxas then knows exactly one thread is active and moves the
// operands to uniform registers directly.  (Under a plain divergent `if (lane == 0)` it wraps every UTCHMMA in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop, ~120 cycles per instruction — measured — which bounds small-N MMAs.)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred px;\n"
      "elect.sync _|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// The three split products of one k-step into one accumulator: hi*hi, hi*lo, lo*hi.  Only the low words of the
// descriptors vary (start address >> 4); the high words are loop invariants, so the issuing thread moves five
// values to uniform registers per group instead of twelve (R2UR latency is what bounds the issue rate).
__device__ __forceinline__ void umma_split3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_top, uint32_t b_hi,
                                            uint32_t b_lo, uint32_t b_top, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, t;\n"
      ".reg .b64 dah, dal, dbh, dbl;\n"
      "mov.b64 dah, {%1, %3};\n"
      "mov.b64 dal, {%2, %3};\n"
      "mov.b64 dbh, {%4, %6};\n"
      "mov.b64 dbl, {%5, %6};\n"
      "setp.ne.b32 p, %8, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %7, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %7, t;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %7, t;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_hi), "r"(a_lo), "r"(a_top), "r"(b_hi), "r"(b_lo), "r"(b_top), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(taddr));
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
  v[2] = __uint_as_float(r2);
  v[3] = __uint_as_float(r3);
}
// fp32 -> (hi, lo) bf16 pair; packs two consecutive channels per 32-bit word
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  const float ar = a - __bfloat162float(ah), br = b - __bfloat162float(bh);
  hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  lo = pack_bf16x2(ar, br);
}

#ifdef JAF_GROUPED_PROFILE
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PROF_STAMP(var) const unsigned long long var = gtime()
#define PROF_DECL(var) unsigned long long var = 0
#define PROF_ACC(acc, ...)                  \
  {                                         \
    const unsigned long long t0_ = gtime(); \
    __VA_ARGS__;                            \
    acc += gtime() - t0_;                   \
  }
#else
#define PROF_STAMP(var)
#define PROF_DECL(var)
#define PROF_ACC(acc, ...) \
  { __VA_ARGS__; }
#endif

// Warp 0: MMA issue (one thread).  Last warp: weight stream (one thread).  The warps between: activation staging, then
// the gate epilogue (warps 1..28 of 32: seven warps per TMEM lane quarter).
__global__ void __launch_bounds__(kThreads, 1)
k_convlstm_grouped(const GArgs a) {
  PROF_STAMP(t_start);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* sA = smem;                       // [hi | lo] x [Ctp/8 chunks][R rows][16 B]
  uint8_t* sB = smem + 2 * a.a_half;        // ring x stage_bytes
  const int ring = a.ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)ring * a.stage_bytes);
  uint64_t* full_bar = bars;                      // [ring]  bulk copy -> MMA
  uint64_t* empty_bar = bars + kMaxRing;          // [ring]  MMA -> bulk copy
  uint64_t* aready_bar = bars + 2 * kMaxRing;     // staging -> MMA
  uint64_t* tfull_bar = bars + 2 * kMaxRing + 1;  // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxRing + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x % a.tiles_per_group;
  const int gs = blockIdx.x / a.tiles_per_group;
  const int g = gs / a.CS, sl = gs % a.CS;  // cell, hidden-channel slice
  const long p_end = a.Q - a.Wp - 1;                          // one past the last output position
  const long p0 = (long)a.Wp + 1 + (long)t * a.MT * 128;      // first output position of this CTA
  const int mt_here = (int)min((long)a.MT, (p_end - p0 + 127) / 128);

  if (threadIdx.x == 0) {
    for (int s = 0; s < ring; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(aready_bar, blockDim.x - 64);
    mbar_init(tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);  // warp-uniform by construction

  const int last_warp = (int)(blockDim.x >> 5) - 1;  // weight stream (never an epilogue warp, see nsets below)
  if (warp == last_warp) {
    // ===================== weight stream: bulk copies through the ring, decoupled from the MMA issue =====================
    // (Issued from the MMA warp the refill of a slot had to wait for that slot's MMAs to COMPLETE before the next
    // stage's MMAs could be issued; per-role timers put the MMA warp of the 48- and 96-channel levels at 250-320 cycles
    // per MMA against the 128-cycle floor.)
    if (lane == 0) {
      const uint8_t* wg = a.wpack + (size_t)g * a.S * 64 * a.N;
      const uint32_t stage_bytes = a.stage_bytes;
      const int nstages = a.nstages, KS = a.KS;
      int slot = 0;
      uint32_t phase = 0;
      for (int st = 0; st < nstages; ++st) {
        if (st >= ring) mbar_wait_relaxed(&empty_bar[slot], phase ^ 1u);  // the MMAs that read the slot's previous stage are done
        // one stage = KS k-steps of this slice's rows.  A k-step of the pack is [hi | lo][2 chunks][N rows][16 B]; a
        // slice is a contiguous run of Ns rows in each of the four blocks (one copy when the cell is not sliced)
        uint8_t* dst = sB + (size_t)slot * stage_bytes;
        mbar_expect_tx(&full_bar[slot], stage_bytes);
        if (a.CS == 1) {
          bulk_g2s(dst, wg + (size_t)st * stage_bytes, stage_bytes, &full_bar[slot]);
        } else {
          const uint32_t run = (uint32_t)a.Ns * 16u;
          for (int j = 0; j < KS; ++j) {
            const uint8_t* src = wg + (size_t)(st * KS + j) * 64 * a.N + (size_t)sl * run;
#pragma unroll
            for (int blk = 0; blk < 4; ++blk)
              bulk_g2s(dst + (size_t)(j * 4 + blk) * run, src + (size_t)blk * 16 * a.N, run, &full_bar[slot]);
          }
        }
        if (++slot == ring) {
          slot = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 0) {
    // ===================== MMA issue (warp-uniform; lane 0 is the leader) =====================
    // The issue rate of this single thread bounds the small-N MMAs, so everything is kept in descriptor
    // units (16 B) and advanced with adds: a descriptor's low 14 bits are the start address >> 4.
    {
      const bool leader = lane == 0;  // == the lane elect.sync picks (lowest active), so commits track these MMAs
      const uint32_t stage_bytes = a.stage_bytes;
      const int nstages = a.nstages, KS = a.KS;
      const uint32_t N = (uint32_t)a.Ns, Nsub = (uint32_t)a.Nsub, R = (uint32_t)a.R, Wp = (uint32_t)a.Wp;  // N: rows of this slice
      const uint32_t idesc = a.idesc;
      const int nsplit = a.nsplit, spt = a.Ctp / 16;
      const uint64_t ad0 = desc_noswz(smem_u32(sA), R * 16u), bd0 = desc_noswz(smem_u32(sB), N * 16u);
      const uint32_t a_top = (uint32_t)(ad0 >> 32), b_top = (uint32_t)(bd0 >> 32);
      const uint32_t a_hi0 = (uint32_t)ad0, a_lo_delta = a.a_half >> 4, b0 = (uint32_t)bd0;
      const uint32_t stage_u = stage_bytes >> 4;
      mbar_wait(aready_bar, 0);
      tc_fence_after();
      PROF_STAMP(t_aready);
      int kc = 0, kx = 0, ky = 0;
      uint32_t accf = 0;
      int slot = 0;
      uint32_t phase = 0;
      for (int st = 0; st < nstages; ++st) {
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        uint32_t bh = b0 + (uint32_t)slot * stage_u;
        for (int j = 0; j < KS; ++j, bh += 4 * N) {
          uint32_t ah = a_hi0 + (uint32_t)(kc * 2) * R + (uint32_t)ky * Wp + (uint32_t)kx;
          uint32_t tm = tmem_base;
          for (int m = 0; m < mt_here; ++m, ah += 128) {
            uint32_t bhh = bh;
            for (int hn = 0; hn < nsplit; ++hn, bhh += Nsub, tm += Nsub) {
              if (elect_one()) umma_split3(tm, ah, ah + a_lo_delta, a_top, bhh, bhh + 2 * N, b_top, idesc, accf);
            }
          }
          accf = 1u;
          if (++kc == spt) {
            kc = 0;
            if (++kx == 3) {
              kx = 0;
              ++ky;
            }
          }
        }
        if (leader) umma_commit(&empty_bar[slot]);
        if (++slot == ring) {
          slot = 0;
          phase ^= 1u;
        }
      }
      if (leader) umma_commit(tfull_bar);
#ifdef JAF_GROUPED_PROFILE
      PROF_STAMP(t_issued);
      mbar_wait(tfull_bar, 0);
      PROF_STAMP(t_done);
      if (blockIdx.x == 0 && leader)
        printf("cta %d: MT=%d R=%d KS=%d nstages=%d | staged at %llu ns, mma issued +%llu, mma done +%llu\n", blockIdx.x,
               a.MT, a.R, a.KS, a.nstages, t_aready - t_start, t_issued - t_aready, t_done - t_aready);
#endif
    }
  } else {
    // ===================== stage the activation rows (all warps but the first and the last) =====================
    const size_t HW = (size_t)a.H * a.W;
    {
      const int wid = threadIdx.x - 32;
      const long q0 = p0 - a.Wp - 1;
      const int npairs = a.Ctp / 16;
      const int items = a.R * npairs;  // (row, pair of 8-channel chunks): 16 independent loads in flight per thread
      const int nworkers = (int)blockDim.x - 64;
      for (int it = wid; it < items; it += nworkers) {
        const int cp = it / a.R, i = it - cp * a.R;
        const long q = q0 + i;
        bool inside = q < a.Q;
        size_t pix = 0;
        int b = 0;
        if (inside) {
          const int u = (int)(q / a.HpWp);  // (image, band)
          const int rem = (int)(q - (long)u * a.HpWp);
          const int yp = rem / a.Wp, xp = rem - yp * a.Wp;
          b = u / a.nb;
          const int x = (u - b * a.nb) * a.Wb + xp - 1;  // a band's halo columns are its neighbours' pixels
          inside = x >= 0 && x < a.W && yp >= 1 && yp <= a.H;
          pix = (size_t)(yp - 1) * a.W + (size_t)x;
        }
        const float* xb = a.x + ((size_t)g * a.B + b) * (size_t)a.xs + pix;
        const float* hb = a.h + ((size_t)g * a.B + b) * (size_t)a.hs + pix;
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int ch = cp * 16 + e;
          float val = 0.f;
          if (inside && ch < a.Ct) val = ch < a.Cin ? __ldg(xb + (size_t)ch * HW) : __ldg(hb + (size_t)(ch - a.Cin) * HW);
          v[e] = val;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint4 hi, lo;
          split2(v[8 * u + 0], v[8 * u + 1], hi.x, lo.x);
          split2(v[8 * u + 2], v[8 * u + 3], hi.y, lo.y);
          split2(v[8 * u + 4], v[8 * u + 5], hi.z, lo.z);
          split2(v[8 * u + 6], v[8 * u + 7], hi.w, lo.w);
          uint8_t* dst = sA + ((size_t)(cp * 2 + u) * a.R + i) * 16;
          *reinterpret_cast<uint4*>(dst) = hi;
          *reinterpret_cast<uint4*>(dst + a.a_half) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core reads
      mbar_arrive(aready_bar);
    }
    // ===================== epilogue: gates (thread = pixel; 7 warps per TMEM lane quarter) =====================
    const int nsets = ((int)(blockDim.x >> 5) - 1) >> 2;  // epilogue warps per TMEM lane quarter (7 or 3)
    if (warp <= 4 * nsets) {
      const int qd = warp & 3, set = (warp - 1) >> 2;  // lane quarter this warp may read; which share of the items
      const int nblk = a.Chs / 4;
      const int nitems = mt_here * nblk;  // (accumulator tile, block of 4 hidden channels)
      const float* bp = a.bias ? a.bias + (size_t)g * a.N : nullptr;
      mbar_wait_relaxed(tfull_bar, 0);
      tc_fence_after();
      for (int it0 = set; it0 < nitems; it0 += 2 * nsets) {
        // two items per pass so that two sets of TMEM / global loads overlap.  The packed row order puts the four
        // gates of a block of four channels in 16 consecutive accumulator columns: one tcgen05.ld per item.
        float acc[2][16], cv[2][4];
        size_t base[2], hbase[2];
        bool valid[2];
        int c0[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int it = it0 + nsets * u;
          valid[u] = false;
          base[u] = 0;
          hbase[u] = 0;
          c0[u] = 0;
          if (it < nitems) {  // warp-uniform
            const int m = it / nblk, jb = it - m * nblk;
            c0[u] = sl * a.Chs + jb * 4;
            const long p = p0 + (long)m * 128 + qd * 32 + lane;
            bool ok = p < p_end;
            if (ok) {
              const int un = (int)(p / a.HpWp);
              const int rem = (int)(p - (long)un * a.HpWp);
              const int yp = rem / a.Wp, xp = rem - yp * a.Wp;
              const int b = un / a.nb;
              const int x = (un - b * a.nb) * a.Wb + xp - 1;
              ok = xp >= 1 && xp <= a.Wb && x < a.W && yp >= 1 && yp <= a.H;
              base[u] = ((size_t)g * a.B + b) * a.Ch * HW + (size_t)(yp - 1) * a.W + (size_t)x;
              hbase[u] = ((size_t)g * a.B + b) * (size_t)a.hos + (size_t)(yp - 1) * a.W + (size_t)x;
            }
            valid[u] = ok;
            tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(m * a.Ns + jb * 16), acc[u]);
#pragma unroll
            for (int e = 0; e < 4; ++e) cv[u][e] = ok ? __ldg(a.c + base[u] + (size_t)(c0[u] + e) * HW) : 0.f;
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!valid[u]) continue;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = c0[u] + e;
            float bi = 0.f, bf = 0.f, bo = 0.f, bg = 0.f;
            if (bp) {
              bi = __ldg(bp + ch);
              bf = __ldg(bp + a.Ch + ch);
              bo = __ldg(bp + 2 * a.Ch + ch);
              bg = __ldg(bp + 3 * a.Ch + ch);
            }
            const float ig = sigmoid_f(acc[u][e] + bi), fg = sigmoid_f(acc[u][4 + e] + bf);
            const float og = sigmoid_f(acc[u][8 + e] + bo), g_ = tanh_f(acc[u][12 + e] + bg);
            const float cn = fg * cv[u][e] + ig * g_;  // src/convLSTM.py:53
            const float hn = og * tanh_f(cn);          // :54
            a.c_out[base[u] + (size_t)ch * HW] = cn;
            a.h_out[hbase[u] + (size_t)ch * HW] = hn;
          }
        }
      }
    }
  }

#ifdef JAF_GROUPED_PROFILE
  if (threadIdx.x == 32 && blockIdx.x == 0) {
    PROF_STAMP(t_end);
    printf("cta %d: epilogue warp 1 done at %llu ns after start\n", blockIdx.x, t_end - t_start);
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
  }
}

// weight [G][4Ch][Ct][3][3] f32 -> per group, per k-step (tap, 16 channels): [hi | lo] x [2 chunks][N rows][8 bf16]
__global__ void __launch_bounds__(256)
k_gpack_weight(const float* __restrict__ w, uint8_t* __restrict__ wp, int G, int N, int Ct, int Ctp) {
  const int spt = Ctp / 16, S = 9 * spt;
  const size_t total = (size_t)G * S * 2 * N * 8;  // one thread per (g, step, chunk, n, e)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = (int)(i % 8);
  size_t r = i / 8;
  const int n = (int)(r % N);
  r /= N;
  const int cc = (int)(r % 2);
  r /= 2;
  const int step = (int)(r % S);
  const int g = (int)(r / S);
  const int tap = step / spt, kc = step % spt;
  const int ch = kc * 16 + cc * 8 + e;
  // packed row n = (hidden channel / 4) * 16 + gate * 4 + hidden channel % 4: any run of 4k channels is a contiguous
  // run of rows (so a cell can be sliced over CTAs) and the four gates of a channel block sit in 16 adjacent columns
  const int Chh = N / 4;
  const int orow = ((n % 16) / 4) * Chh + (n / 16) * 4 + n % 4;  // reference row: gate * Ch + channel (:46)
  const float v = ch < Ct ? w[(((size_t)g * N + orow) * Ct + ch) * 9 + tap] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  uint8_t* stepb = wp + ((size_t)g * S + step) * 64 * N;
  const size_t off = (size_t)cc * 16 * N + (size_t)n * 16 + (size_t)e * 2;
  *reinterpret_cast<__nv_bfloat16*>(stepb + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(stepb + 32 * (size_t)N + off) = lo;
}


// =====================================================================================
// Operand-swapped, persistent, warp-specialised flavour for the narrow cells (Ch <= 32, Cin + Ch <= 48: the 12- and
// 24-channel levels of the reference pyramid, which hold 85 % of its pixels).
//
// Why swapped.  In k_convlstm_grouped the pixels are the M side (128 rows) and the gate rows the N side: an MMA costs
// 128 cycles whatever N is (tools/probes/umma_probe.cu), so with N = 48 / 96 the tensor pipe does 19 / 38 % of its work
// per issued MMA (ncu: tensor pipe 16-20 % active).  Here the WEIGHTS are the A operand (M = 128, of which the 4*Ch rows
// `gate*Ch + channel` — the reference's own row order, :46 — are real) and 256 PIXELS the B operand (N = 256): the same
// 128 cycles cover 256 pixels.  Both operands keep the K-major no-swizzle layout, so the staged pixel window serves all
// nine taps exactly as before.
//
// K is flattened over (8-channel chunk, tap): a k-step is two "units" of 8 channels, each with its own tap shift — the
// two K chunks of a no-swizzle descriptor may lie anywhere (LBO), so unit (chunk c, tap t) is simply the address
// c*R + shift(t).  Units are ordered chunk-major, which keeps LBO positive.  24 input channels x 9 taps are 27 units =
// 14 k-steps instead of 18 with 16-channel k-steps (Ctp = 32), and the staged window holds 3 chunks instead of 4.
//
// Weight rows are packed densely: one k-step is [hi | lo][2 units][4*Ch rows][16 B] = 256*Ch bytes (3 KB at Ch = 12
// instead of 8 KB for 128 padded rows).  The M = 128 descriptor reads on into whatever follows the 4*Ch rows; those
// accumulator rows (TMEM lanes >= 4*Ch) are never read.
//
// Why persistent.  A 256-pixel tile costs ~3 us of staging, ~3-5 us of MMAs and ~4 us of gate epilogue; run one after
// the other by two CTAs per SM (the first swapped kernel) the tensor pipe was 28 % active.  ONE CTA per SM now walks a
// contiguous run of tiles and the phases of consecutive tiles overlap:
//   warp 0        MMA issue: waits for the staged pixels of tile i and a free accumulator, issues the 3*S MMAs
//   warp 1        weight stream: one cp.async.bulk per k-step through a deep ring, running ahead across tile boundaries
//   warps 2..11   stage the pixel window of tile i+1 (thread = row, every channel in flight; fp32 NCHW -> hi/lo bf16)
//   warps 12..27  gate epilogue of tile i-1 from the other of two 256-column TMEM accumulators: four sets of four warps.
//                 Warp q of a set may read TMEM lane quarter q only, and a lane holds ONE gate row for 32 pixels, so the
//                 set transposes through shared memory: each warp dumps its RAW rows, then thread = pixel applies bias,
//                 the five activations and c' = f*c + i*g, h' = o*tanh(c') for channels q, q+4, ... on full warps with
//                 coalesced NCHW accesses.  The previous cell state of the NEXT chunk is loaded one chunk ahead.
// =====================================================================================
constexpr int kTileT = 256;          // pixels (flattened padded positions) per tile = N of the MMA
constexpr int kXPitch = 33;          // floats per gate row of the exchange buffer (32 pixels + 1: no bank conflicts)
constexpr int kThreadsP = 896;
constexpr int kStagerWarp0 = 2, kStagerWarps = 10, kEpiWarp0 = 12, kEpiSets = 4;
constexpr int kChunksPerSet = kTileT / 32 / kEpiSets;  // 2
constexpr int kMaxRingP = 24;
constexpr int kRingPad = 2048;       // the M = 128 descriptor of the last ring slot reads up to 128 rows x 16 B past a unit

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// n / d for n < 2^32 with m = min(floor(2^32 / d), 2^32 - 1): umulhi gives the quotient or one less
__device__ __forceinline__ uint32_t fdiv(uint32_t n, FastDiv f, uint32_t& rem) {
  uint32_t q = __umulhi(n, f.m);
  rem = n - q * f.d;
  if (rem >= f.d) {
    ++q;
    rem -= f.d;
  }
  return q;
}

// Gate epilogue role of the persistent kernel, NU = Ch / 4 channels per thread (a template parameter, so that the NU
// channels of the pixel stage are independent instruction streams without per-channel branches).
template <int NU>
__device__ __forceinline__ void epilogue_role_p(const GArgs& a, float* sX, uint64_t* tfull, uint64_t* tempty, uint32_t tmem_base,
                                                long first, int n_here, long p_end) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t HW = (size_t)a.H * a.W;
  const int q = warp & 3, set = (warp - kEpiWarp0) >> 2;
  constexpr int Ch = 4 * NU, rows = 4 * Ch;
  constexpr int xset_floats = rows * kXPitch + rows;  // raw gate rows [gate*Ch + channel][kXPitch] + the cell's bias
  float* X = sX + (size_t)set * xset_floats;
  float* sbias = X + rows * kXPitch;
  const int my_rows = rows - 32 * q;  // rows of this warp's TMEM lane quarter that exist (<= 0: nothing to dump)
  const int nit = kChunksPerSet * n_here;
  PROF_DECL(w_f);
  PROF_STAMP(t_begin);

  // where the pixel of (iteration, lane) lives + its previous cell state, one iteration ahead of their use
  bool ok_c = false, ok_n = false;
  float* cop_c = a.c_out;
  float* hop_c = a.h_out;
  float* cop_n = a.c_out;
  float* hop_n = a.h_out;
  float c_c[NU], c_n[NU];
  int g_c = 0, g_n = 0;
  bool act_c = false, act_n = false;
  auto prep = [&](int it) {
    const long tile = first + it / kChunksPerSet;
    const int chunk = set + kEpiSets * (it % kChunksPerSet);
    uint32_t t_in_g;
    g_n = (int)fdiv((uint32_t)tile, a.d_tpg, t_in_g);
    const long p0 = (long)a.Wp + 1 + (long)t_in_g * kTileT + (long)chunk * 32;
    act_n = p0 < p_end;  // uniform over the set
#ifdef JAF_PROBE_SKIP_EPI
    act_n = false;
#endif
    const long p = p0 + lane;
    ok_n = p < p_end;
    const float* cp = a.c;
    cop_n = a.c_out;
    hop_n = a.h_out;
    if (ok_n) {
      uint32_t rem, xp, ub;
      const uint32_t un = fdiv((uint32_t)p, a.d_hpwp, rem);
      const uint32_t yp = fdiv(rem, a.d_wp, xp);
      const uint32_t b = fdiv(un, a.d_nb, ub);
      const int x = (int)ub * a.Wb + (int)xp - 1;
      ok_n = xp >= 1 && (int)xp <= a.Wb && x < a.W && yp >= 1 && (int)yp <= a.H;
      const size_t pix = (size_t)(yp - 1) * a.W + (size_t)x + (size_t)q * HW;
      const size_t base = ((size_t)g_n * a.B + b) * Ch * HW + pix;
      cp += base;
      cop_n += base;
      hop_n += ((size_t)g_n * a.B + b) * (size_t)a.hos + pix;
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) c_n[u] = ok_n ? __ldg(cp + (size_t)(4 * u) * HW) : 0.f;
  };
  auto shift = [&]() {
    ok_c = ok_n; act_c = act_n; g_c = g_n; cop_c = cop_n; hop_c = hop_n;
#pragma unroll
    for (int u = 0; u < NU; ++u) c_c[u] = c_n[u];
  };
  if (nit > 0) prep(0);
  int g_bias = -1;
  for (int it = 0; it < nit; ++it) {
    shift();
    const int i = it / kChunksPerSet, part = it % kChunksPerSet;
    const int buf = i & 1;
    const uint32_t ph = (uint32_t)(i >> 1) & 1u;
    const int chunk = set + kEpiSets * part;
    if (part == 0) {
      if (g_c != g_bias) {  // set-uniform; the last barrier of the previous chunk has passed, the next one orders the writes
        g_bias = g_c;
        const int t4 = q * 32 + lane;
        if (t4 < rows) sbias[t4] = a.bias != nullptr ? __ldg(a.bias + (size_t)g_c * rows + t4) : 0.f;
      }
      PROF_ACC(w_f, mbar_wait_relaxed(&tfull[buf], ph));
      tc_fence_after();
    }
    // (1) dump this warp's raw gate rows of the chunk (TMEM lane = row gate*Ch + channel, column = pixel)
    if (act_c && my_rows > 0) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kTileT + chunk * 32), v);
      tmem_ld_wait();
      if (lane < my_rows) {
        float* xr = X + (size_t)(q * 32 + lane) * kXPitch;
#pragma unroll
        for (int j = 0; j < 32; ++j) xr[j] = v[j];
      }
    }
    if (part == kChunksPerSet - 1) {  // this warp has read everything it needs from the accumulator: free it early
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    // (0') next chunk's addresses and cell-state loads: in flight during the pixel stage below
    if (it + 1 < nit) prep(it + 1);
    if (act_c) {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
      // (2) thread = pixel, channels q, q + 4, ...
      if (ok_c) {
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int ch = q + 4 * u;
          const float* xc = X + ch * kXPitch + lane;
          const float ig = sigmoid_f(xc[0 * Ch * kXPitch] + sbias[ch]), fg = sigmoid_f(xc[1 * Ch * kXPitch] + sbias[Ch + ch]);
          const float og = sigmoid_f(xc[2 * Ch * kXPitch] + sbias[2 * Ch + ch]), g_ = tanh_f(xc[3 * Ch * kXPitch] + sbias[3 * Ch + ch]);
          const float cn = fg * c_c[u] + ig * g_;        // src/convLSTM.py:53
          cop_c[(size_t)(4 * u) * HW] = cn;
          hop_c[(size_t)(4 * u) * HW] = og * tanh_f(cn);  // :54
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");  // the buffer is rewritten by the next chunk
    }
  }
#ifdef JAF_GROUPED_PROFILE
  if (lane == 0 && set == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
    printf("cta %d epilogue warp %d: total %llu ns | wait tfull %llu\n", blockIdx.x, warp, gtime() - t_begin, w_f);
#endif
}

__global__ void __launch_bounds__(kThreadsP, 1)
k_convlstm_grouped_p(const GArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const uint32_t pbuf_bytes = (2u * a.a_half + 127u) & ~127u;
  uint8_t* sP = smem;                                    // 2 x pixels: [hi | lo] x [C8 chunks][R rows][16 B]
  float* sX = reinterpret_cast<float*>(smem + 2 * pbuf_bytes);  // kEpiSets x ([4*Ch rows][kXPitch] + bias)
  const uint32_t xset_floats = 4u * (uint32_t)a.Ch * (kXPitch + 1);
  uint8_t* sW = smem + 2 * pbuf_bytes + ((kEpiSets * xset_floats * 4u + 127u) & ~127u);  // ring x one k-step (+ kRingPad)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)a.ring * a.stage_bytes + kRingPad);
  uint64_t* wfull = bars;                     // [ring] bulk copy -> MMA
  uint64_t* wempty = bars + kMaxRingP;        // [ring] MMA -> bulk copy
  uint64_t* pfull = bars + 2 * kMaxRingP;     // [2] stagers -> MMA
  uint64_t* pempty = pfull + 2;               // [2] MMA -> stagers
  uint64_t* tfull = pfull + 4;                // [2] MMA -> epilogue
  uint64_t* tempty = pfull + 6;               // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pfull + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long first = (long)blockIdx.x * a.ntiles / gridDim.x;
  const int n_here = (int)((long)(blockIdx.x + 1) * a.ntiles / gridDim.x - first);
  const long p_end = a.Q - a.Wp - 1;  // one past the last output position of a cell

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.ring; ++s) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pfull[s], kStagerWarps);
      mbar_init(&pempty[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiSets * 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);

  if (warp == 0) {
    // ===================== MMA issue =====================
    const bool leader = lane == 0;
    const uint32_t R = (uint32_t)a.R, Wp = (uint32_t)a.Wp, rows = 4u * (uint32_t)a.Ch;
    const int S = a.S, ring = a.ring, U = a.U;
    // A = weights: 128 rows of which `rows` are packed, the two units of a k-step `rows` rows apart; B = pixels
    const uint64_t wd0 = desc_noswz(smem_u32(sW), rows * 16u), pd0 = desc_noswz(smem_u32(sP), 0);
    const uint32_t w_top = (uint32_t)(wd0 >> 32), p_top = (uint32_t)(pd0 >> 32);
    const uint32_t w0 = (uint32_t)wd0, p0u = (uint32_t)pd0, p_lo_delta = a.a_half >> 4;
    const uint32_t stage_u = a.stage_bytes >> 4;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int slot = 0;
    uint32_t wph = 0;
    PROF_DECL(w_t);
    PROF_DECL(w_p);
    PROF_DECL(w_w);
    PROF_STAMP(t_begin);
    for (int i = 0; i < n_here; ++i) {
      const int buf = i & 1;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      PROF_ACC(w_t, mbar_wait(&tempty[buf], ph ^ 1u));
      PROF_ACC(w_p, mbar_wait(&pfull[buf], ph));
      tc_fence_after();
      const uint32_t tm = tmem_base + (uint32_t)buf * kTileT;
      const uint32_t pb = p0u + (uint32_t)buf * (pbuf_bytes >> 4);
      // unit = (8-channel chunk c8, tap (ky, kx)) at c8*R + ky*Wp + kx rows; chunk-major order
      uint32_t ubase = 0, kx = 0, ky = 0;
      int u = 0;
      uint32_t accf = 0;
      for (int ks = 0; ks < S; ++ks) {
        const uint32_t ad0 = ubase + ky * Wp + kx;
        if (++kx == 3) { kx = 0; if (++ky == 3) { ky = 0; ubase += R; } }
        uint32_t ad1 = ad0;  // odd unit count: the last k-step's second unit has zero weights and re-reads the first
        if (++u < U) {
          ad1 = ubase + ky * Wp + kx;
          if (++kx == 3) { kx = 0; if (++ky == 3) { ky = 0; ubase += R; } }
          ++u;
        }
        PROF_ACC(w_w, mbar_wait(&wfull[slot], wph));
        tc_fence_after();
        const uint32_t wh = w0 + (uint32_t)slot * stage_u;
        const uint32_t phd = (pb + ad0) | ((ad1 - ad0) << 16);
        if (elect_one()) umma_split3(tm, wh, wh + 2u * rows, w_top, phd, phd + p_lo_delta, p_top, idesc, accf);
        if (leader) umma_commit(&wempty[slot]);
        accf = 1u;
        if (++slot == ring) {
          slot = 0;
          wph ^= 1u;
        }
      }
      if (leader) {
        umma_commit(&pempty[buf]);
        umma_commit(&tfull[buf]);
      }
    }
#ifdef JAF_GROUPED_PROFILE
    if (leader && (blockIdx.x == 0 || blockIdx.x == 77))
      printf("cta %d mma: %d tiles, total %llu ns | wait tempty %llu, pfull %llu, wfull %llu\n", blockIdx.x, n_here,
             gtime() - t_begin, w_t, w_p, w_w);
#endif
  } else if (warp == 1) {
    // ===================== weight stream =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t wph = 0;
      for (int i = 0; i < n_here; ++i) {
        uint32_t t_in_g_unused;
        const int g = (int)fdiv((uint32_t)(first + i), a.d_tpg, t_in_g_unused);
        const uint8_t* wg = a.wpack + (size_t)g * a.S * a.stage_bytes;
        for (int ks = 0; ks < a.S; ++ks) {
          mbar_wait_relaxed(&wempty[slot], wph ^ 1u);
          mbar_expect_tx(&wfull[slot], a.stage_bytes);
          bulk_g2s(sW + (size_t)slot * a.stage_bytes, wg + (size_t)ks * a.stage_bytes, a.stage_bytes, &wfull[slot]);
          if (++slot == a.ring) {
            slot = 0;
            wph ^= 1u;
          }
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ===================== stage the pixel rows: thread = row, every channel of the row in flight at once =====================
    const size_t HW = (size_t)a.H * a.W;
    const int wid = threadIdx.x - kStagerWarp0 * 32;
    const int C8 = a.C8;
    PROF_DECL(w_e);
    PROF_STAMP(t_begin);
    for (int i = 0; i < n_here; ++i) {
      const int buf = i & 1;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      const long tile = first + i;
      uint32_t t_in_g;
      const int g = (int)fdiv((uint32_t)tile, a.d_tpg, t_in_g);
      const long p0 = (long)a.Wp + 1 + (long)t_in_g * kTileT;
      const long q0 = p0 - a.Wp - 1;
      uint8_t* dstb = sP + (size_t)buf * pbuf_bytes;
      PROF_ACC(w_e, mbar_wait_relaxed(&pempty[buf], ph ^ 1u));
#ifdef JAF_PROBE_SKIP_STAGE
      for (int r = wid + (1 << 30); r < a.R; r += kStagerWarps * 32) {
#else
      for (int r = wid; r < a.R; r += kStagerWarps * 32) {
#endif
        const long q = q0 + r;
        bool inside = q < a.Q;
        size_t pix = 0;
        uint32_t b = 0;
        if (inside) {
          uint32_t rem, xp, ub;
          const uint32_t u = fdiv((uint32_t)q, a.d_hpwp, rem);  // (image, band)
          const uint32_t yp = fdiv(rem, a.d_wp, xp);
          b = fdiv(u, a.d_nb, ub);
          const int x = (int)ub * a.Wb + (int)xp - 1;  // a band's halo columns are its neighbours' pixels
          inside = x >= 0 && x < a.W && yp >= 1 && (int)yp <= a.H;
          pix = inside ? (size_t)(yp - 1) * a.W + (size_t)x : 0;
          if (!inside) b = 0;
        }
        // padding rows read pixel 0 of image 0 (a valid address) and are zeroed below: the loads stay unconditional
        const float* xb = a.x + ((size_t)g * a.B + b) * (size_t)a.xs + pix;
        const float* hb = a.h + ((size_t)g * a.B + b) * (size_t)a.hs + pix - (size_t)a.Cin * HW;
        for (int c0 = 0; c0 < C8; c0 += 4) {  // batches of 4 chunks = 32 channels: all their loads are issued before the first use
          float v[32];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + u < C8) {  // uniform
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int ch = min((c0 + u) * 8 + e, a.Ct - 1);  // a partial last chunk re-reads the last channel (zeroed below)
                v[8 * u + e] = __ldg((ch < a.Cin ? xb : hb) + (size_t)ch * HW);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + u < C8) {
              float w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w[e] = (inside && (c0 + u) * 8 + e < a.Ct) ? v[8 * u + e] : 0.f;
              uint4 hi, lo;
              split2(w[0], w[1], hi.x, lo.x);
              split2(w[2], w[3], hi.y, lo.y);
              split2(w[4], w[5], hi.z, lo.z);
              split2(w[6], w[7], hi.w, lo.w);
              uint8_t* dst = dstb + ((size_t)(c0 + u) * a.R + r) * 16;
              *reinterpret_cast<uint4*>(dst) = hi;
              *reinterpret_cast<uint4*>(dst + a.a_half) = lo;
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[buf]);
    }
#ifdef JAF_GROUPED_PROFILE
    if (wid == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
      printf("cta %d stager: total %llu ns | wait pempty %llu\n", blockIdx.x, gtime() - t_begin, w_e);
#endif
  } else {
    // ===================== epilogue: four sets of four warps =====================
    switch (a.Ch >> 2) {
      case 1: epilogue_role_p<1>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
      case 2: epilogue_role_p<2>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
      case 3: epilogue_role_p<3>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
      case 4: epilogue_role_p<4>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
      case 5: epilogue_role_p<5>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
      case 6: epilogue_role_p<6>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
      case 7: epilogue_role_p<7>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
      default: epilogue_role_p<8>(a, sX, tfull, tempty, tmem_base, first, n_here, p_end); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// weight [G][4Ch][Ct][3][3] f32 -> per cell, per k-step j (units 2j, 2j+1; unit u = 8-channel chunk u/9, tap u%9):
// [hi | lo] x [2 units][4*Ch rows][8 bf16], rows in the reference's order gate*Ch + channel; zero where the unit or the
// input channel does not exist
__global__ void __launch_bounds__(256)
k_gpack_weight_p(const float* __restrict__ w, uint8_t* __restrict__ wp, int G, int Ch, int Ct, int U, int S) {
  const int rows = 4 * Ch;
  const size_t total = (size_t)G * S * 2 * rows * 8;  // one thread per (g, k-step, unit slot, row, e)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = (int)(i % 8);
  size_t r = i / 8;
  const int row = (int)(r % rows);
  r /= rows;
  const int us = (int)(r % 2);
  r /= 2;
  const int step = (int)(r % S);
  const int g = (int)(r / S);
  const int u = 2 * step + us;
  const int c8 = u / 9, tap = u % 9;
  const int ch_in = c8 * 8 + e;
  float v = 0.f;
  if (u < U && ch_in < Ct) v = w[(((size_t)g * rows + row) * Ct + ch_in) * 9 + tap];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  uint8_t* stepb = wp + ((size_t)g * S + step) * (64 * (size_t)rows);
  const size_t off = (size_t)us * 16 * rows + (size_t)row * 16 + (size_t)e * 2;
  *reinterpret_cast<__nv_bfloat16*>(stepb + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(stepb + 32 * (size_t)rows + off) = lo;
}

// Cells served by the operand-swapped persistent kernel (the packed-weight format follows this rule, so it must not
// change between packing and stepping: the environment is read once per process).  Ch <= 32 keeps the 4*Ch gate rows
// within one M = 128 tile; two pixel buffers + the exchange buffers + a weight ring of >= 4 k-steps must fit 227 KB for
// every map size (the band width is capped, see plan_grouped_p): Cin + Ch <= 48 at Ch <= 24, <= 40 at Ch = 32.  Measured on the reference pyramid (24 cells,
// B = 1; profiles/r02_convlstm_grouped.jsonl).  JAF_CG_SWAP=0 sends every cell to k_convlstm_grouped (A/B runs).
bool grouped_swapped(int Cin, int Ch) {
  static const int mode = [] {
    const char* e = getenv("JAF_CG_SWAP");
    return e == nullptr ? 1 : atoi(e);
  }();
  if (mode == 0 || Ch > 32) return false;
  // worst case over map sizes: bands are at most 37 columns wide (plan_grouped_p), i.e. R <= 256 + 2 * 39 + 2 rows
  const size_t pbuf = ((size_t)2 * ((Cin + Ch + 7) / 8) * 336 * 16 + 127) & ~(size_t)127;
  const size_t xb = ((size_t)kEpiSets * 4 * Ch * (kXPitch + 1) * 4 + 127) & ~(size_t)127;
  return 2 * pbuf + xb + kRingPad + (2 * kMaxRingP + 9) * 8 + 128 + 4 * (size_t)64 * 4 * Ch <= (size_t)kMaxSmem;
}

bool grouped_shape_ok(int Cin, int Ch) { return Cin > 0 && Ch > 0 && Ch % 4 == 0 && 4 * Ch <= 512 && ((4 * Ch <= 256) || (2 * Ch) % 16 == 0); }

}  // namespace

extern "C" {

// k-steps of the swapped format: two 8-channel units per step, 9 taps per 8-channel chunk
static inline int swapped_units(int Cin, int Ch) { return 9 * ((Cin + Ch + 7) / 8); }
static inline int swapped_steps(int Cin, int Ch) { return (swapped_units(Cin, Ch) + 1) / 2; }

size_t jaf_convlstm_gpack_bytes(int G, int Cin, int Ch) {
  if (G <= 0 || !grouped_shape_ok(Cin, Ch)) return 0;
  if (grouped_swapped(Cin, Ch)) return (size_t)G * swapped_steps(Cin, Ch) * 64 * (4 * Ch);
  const int Ctp = (Cin + Ch + 15) / 16 * 16;
  return (size_t)G * 9 * (Ctp / 16) * 64 * (4 * Ch);
}

int jaf_convlstm_gpack_weight(const float* weight, int G, int Cin, int Ch, void* wpack, void* stream) {
  JAF_REQUIRE(weight && wpack, "null pointer");
  JAF_REQUIRE(G > 0 && grouped_shape_ok(Cin, Ch), "Ch must be a multiple of 4 and at most 128");
  const int Ct = Cin + Ch, Ctp = (Ct + 15) / 16 * 16, N = 4 * Ch;
  if (grouped_swapped(Cin, Ch)) {
    const int U = swapped_units(Cin, Ch), S = swapped_steps(Cin, Ch);
    const size_t total_p = (size_t)G * S * 2 * N * 8;
    k_gpack_weight_p<<<jaf::ceil_div((long)total_p, 256), 256, 0, jaf::as_stream(stream)>>>(weight, static_cast<uint8_t*>(wpack),
                                                                                         G, Ch, Ct, U, S);
    return jaf::finish_launch("k_gpack_weight_p");
  }
  const size_t total = (size_t)G * 9 * (Ctp / 16) * 2 * N * 8;
  k_gpack_weight<<<jaf::ceil_div((long)total, 256), 256, 0, jaf::as_stream(stream)>>>(
      weight, static_cast<uint8_t*>(wpack), G, N, Ct, Ctp);
  return jaf::finish_launch("k_gpack_weight");
}

// Launch plan of one grouped step (shared by the launcher and jaf_convlstm_grouped_supported): fills `a`, the launch
// shape and the dynamic shared memory.  Returns JAF_OK, or JAF_ERR_UNSUPPORTED when the cell does not fit the SM.
static int plan_grouped(int G, int B, int Cin, int Ch, int H, int W, int sm_count, GArgs& a, int& threads, size_t& smem,
                        long& grid) {
  a.G = G; a.B = B; a.Cin = Cin; a.Ch = Ch; a.H = H; a.W = W;
  a.xs = (long)Cin * H * W;
  a.hs = a.hos = (long)Ch * H * W;
  // vertical bands of ~25 columns: the halo of a flattened tile is two padded rows, so narrow bands keep it short
  // (W = 200: 2 x 28 rows instead of 2 x 203) at the price of two extra columns per band row
  static const int band_target = [] {
    const char* e = getenv("JAF_CG_BAND");
    const int v = e ? atoi(e) : 25;  // measured best of 16 / 25 / 34 / 50 / 67 / 100 on the reference pyramid
    return v >= 8 ? v : 25;
  }();
  a.nb = (W + band_target / 2) / band_target;
  if (a.nb < 1) a.nb = 1;
  a.Wb = (W + a.nb - 1) / a.nb;
  a.Wp = a.Wb + 2;
  a.HpWp = (H + 2) * a.Wp;
  a.Q = (long)B * a.nb * a.HpWp;
  a.Ct = Cin + Ch;
  a.Ctp = (a.Ct + 15) / 16 * 16;
  a.N = 4 * Ch;
  a.S = 9 * (a.Ctp / 16);
  const long out_rows = a.Q - 2L * a.Wp - 2;
  const int tiles_total = jaf::ceil_div(out_rows, 128);
  static const int forced_mode = [] {
    const char* e = getenv("JAF_CG_MODE");
    return e ? atoi(e) : 0;
  }();
  static const int forced_cs = [] {
    const char* e = getenv("JAF_CG_SLICES");
    return e ? atoi(e) : 0;
  }();
  // Hidden-channel slices: a small map with wide cells (the 96-channel 13x13 level) has too few pixel tiles to fill
  // the GPU; its cell is then split over CS CTAs, each with Ch/CS hidden channels (4*Ch/CS gate rows, all inputs).
  int CS = 1;
  if ((long)G * tiles_total * 2 <= sm_count) {
    for (int cs = 2; cs <= 8; ++cs)
      if (Ch % cs == 0 && (Ch / cs) % 4 == 0 && (long)G * tiles_total * cs <= sm_count + sm_count / 4) CS = cs;
  }
  if (forced_cs >= 1 && Ch % forced_cs == 0 && (Ch / forced_cs) % 4 == 0) CS = forced_cs;
  a.CS = CS;
  a.Chs = Ch / CS;
  a.Ns = 4 * a.Chs;
  a.nsplit = a.Ns > 256 ? 2 : 1;
  a.Nsub = a.Ns / a.nsplit;
  // Two launch shapes.  "Half": 512 threads, <= 112 KB and <= 256 TMEM columns per CTA, so two CTAs share an SM and
  // one stages / runs its epilogue while the other's MMAs execute.  "Full": 1024 threads, the whole SM, the largest
  // MT (smallest halo overhead).  Half is used when it fits and the grid is more than one wave of it.
  constexpr long kBarBytes = (2 * kMaxRing + 2) * 8 + 16 + 128;  // barriers + TMEM slot + alignment slack
  // the largest stage (KS k-steps, a divisor of S, at most stage_cap bytes) with which at least one accumulator tile and
  // a ring of kRing stages fit; then the largest MT
  auto plan = [&](long smem_cap, int col_cap, long stage_cap, int& KS, int& MT) {
    KS = 1;
    MT = 0;
    for (int ks = a.S; ks >= 1 && MT == 0; --ks) {
      if (a.S % ks != 0 || (ks > 1 && (long)ks * 64 * a.Ns > stage_cap)) continue;
      for (int m = 1; m <= 8; ++m) {
        const long R = (long)m * 128 + 2L * a.Wp + 2;
        const long sm = 4L * a.Ctp * R + (long)kRing * ks * 64 * a.Ns + kBarBytes;
        if ((long)m * a.Ns <= col_cap && sm <= smem_cap && m <= tiles_total) MT = m;
      }
      if (MT > 0) KS = ks;
    }
  };
  int ks_full, mt_full, ks_half, mt_half, ks_q, mt_q;
  // small stages: what counts for the weight stream is the bytes in flight, and a slot is refilled only when all its
  // k-steps have been consumed (per-role timers on the 48- and 96-channel levels: the MMA warp spent half its time
  // waiting for weights with three 24 KB stages)
  static const long stage_cap_full = [] {
    const char* e = getenv("JAF_CG_STAGE_KB");
    const int v = e ? atoi(e) : 28;
    return (long)(v >= 1 ? v : 28) * 1024;
  }();
  plan(kMaxSmem, 512, stage_cap_full, ks_full, mt_full);
  plan(kHalfSmem, 256, 10 * 1024, ks_half, mt_half);
  plan(kQuarterSmem, 128, 6 * 1024, ks_q, mt_q);
  if (mt_full < 1) {
    jaf::set_error("jaf_convlstm_step_grouped: cell does not fit shared memory (Cin=%d Ch=%d W=%d); use "
                   "jaf_convlstm_step_f32", Cin, Ch, W);
    return JAF_ERR_UNSUPPORTED;
  }
  bool half = mt_half >= 1 && (long)G * CS * jaf::ceil_div(tiles_total, mt_half) > sm_count;
  if (forced_mode == 1) half = false;
  if (forced_mode == 2 && mt_half >= 1) half = true;
  // quarter shape: JAF_CG_MODE=3 forces it; JAF_CG_MODE=0 (auto) takes it when it fits and the grid is at least two
  // waves of it (measured on the reference pyramid: profiles/r02_convlstm_grouped.jsonl)
  bool quarter = mt_q >= 1 && forced_mode == 3;
  static const int auto_quarter = [] {
    const char* e = getenv("JAF_CG_AUTO_QUARTER");
    return e ? atoi(e) : 0;
  }();
  if (forced_mode == 0 && auto_quarter && mt_q >= 1 && (long)G * CS * jaf::ceil_div(tiles_total, mt_q) > 8L * sm_count) quarter = true;
  if (quarter) half = true;  // shares the "not one CTA per SM" handling below
  int MT = quarter ? mt_q : (half ? mt_half : mt_full);
  a.KS = quarter ? ks_q : (half ? ks_half : ks_full);
  threads = quarter ? kThreadsQuarter : (half ? kThreadsHalf : kThreads);
  a.nstages = a.S / a.KS;
  a.stage_bytes = (uint32_t)a.KS * 64u * (uint32_t)a.Ns;
  // keep at least one CTA per SM
  while (!half && MT > 1 && (long)G * CS * jaf::ceil_div(tiles_total, MT) < sm_count) --MT;
  a.MT = MT;
  a.R = MT * 128 + 2 * a.Wp + 2;
  a.a_half = (uint32_t)(a.Ctp / 8) * (uint32_t)a.R * 16u;
  if (!((uint32_t)a.R * 16u < (1u << 18))) {
    jaf::set_error("jaf_convlstm_step_grouped: row window too large for the descriptor");
    return JAF_ERR_UNSUPPORTED;
  }
  a.tiles_per_group = jaf::ceil_div(tiles_total, MT);
  uint32_t cols = 32;
  while (cols < (uint32_t)(MT * a.Ns)) cols <<= 1;
  a.tmem_cols = cols;
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.Nsub >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // the ring takes the shared memory the shape leaves free (the whole SM for the full shape)
  const long cap = quarter ? kQuarterSmem : (half ? kHalfSmem : kMaxSmem);
  long ring = (cap - 2L * a.a_half - kBarBytes) / (long)a.stage_bytes;
  static const int ring_cap = [] {
    const char* e = getenv("JAF_CG_RING");
    const int v = e ? atoi(e) : kMaxRing;
    return v >= kRing && v <= kMaxRing ? v : kMaxRing;
  }();
  if (ring > ring_cap) ring = ring_cap;
  if (ring > a.nstages) ring = a.nstages;
  if (ring < 1) ring = 1;
  a.ring = (int)ring;
  smem = 2 * (size_t)a.a_half + (size_t)ring * a.stage_bytes + kBarBytes;
  grid = (long)G * CS * a.tiles_per_group;
  if (grid >= (1L << 31)) {
    jaf::set_error("jaf_convlstm_step_grouped: too many tiles");
    return JAF_ERR_UNSUPPORTED;
  }
  return JAF_OK;
}

// Launch plan of the operand-swapped persistent kernel: the geometry of plan_grouped (bands of <= 25 columns) with
// 256-pixel tiles, one CTA per SM walking a contiguous run of tiles.
static int plan_grouped_p(int G, int B, int Cin, int Ch, int H, int W, int sm_count, GArgs& a, size_t& smem, long& grid) {
  a.G = G; a.B = B; a.Cin = Cin; a.Ch = Ch; a.H = H; a.W = W;
  a.xs = (long)Cin * H * W;
  a.hs = a.hos = (long)Ch * H * W;
  static const int band_target = [] {
    const char* e = getenv("JAF_CG_BAND");
    const int v = e ? atoi(e) : 25;
    return (v >= 8 && v <= 25) ? v : 25;  // capped: the shared-memory budget of grouped_swapped() assumes it
  }();
  a.nb = (W + band_target / 2) / band_target;
  if (a.nb < 1) a.nb = 1;
  a.Wb = (W + a.nb - 1) / a.nb;
  a.Wp = a.Wb + 2;
  a.HpWp = (H + 2) * a.Wp;
  a.Q = (long)B * a.nb * a.HpWp;
  a.Ct = Cin + Ch;
  a.C8 = (a.Ct + 7) / 8;
  a.Ctp = a.C8 * 8;
  a.U = 9 * a.C8;
  a.S = (a.U + 1) / 2;
  a.N = 4 * Ch;
  a.CS = 1; a.Chs = Ch; a.Ns = a.N; a.nsplit = 1; a.Nsub = a.N;
  a.MT = kTileT / 128;
  a.R = kTileT + 2 * a.Wp + 2;
  a.a_half = (uint32_t)a.C8 * (uint32_t)a.R * 16u;
  a.KS = 1;
  a.nstages = a.S;
  a.stage_bytes = 64u * (uint32_t)a.N;
  a.tmem_cols = 512;
  a.idesc = 0;
  const long out_rows = a.Q - 2L * a.Wp - 2;
  a.tiles_per_group = jaf::ceil_div(out_rows, kTileT);
  a.ntiles = (long)G * a.tiles_per_group;
  auto mk = [](int d) {
    FastDiv f;
    f.d = (uint32_t)d;
    const uint64_t m = (1ull << 32) / (uint64_t)d;
    f.m = m > 0xffffffffull ? 0xffffffffu : (uint32_t)m;
    return f;
  };
  a.d_hpwp = mk(a.HpWp);
  a.d_wp = mk(a.Wp);
  a.d_nb = mk(a.nb);
  a.d_tpg = mk(a.tiles_per_group);
  const size_t pbuf = ((size_t)2 * a.a_half + 127) & ~(size_t)127;
  const size_t xb = ((size_t)kEpiSets * 4 * Ch * (kXPitch + 1) * 4 + 127) & ~(size_t)127;
  const size_t fixed = 2 * pbuf + xb + kRingPad + (2 * kMaxRingP + 9) * 8 + 128;
  if (a.Q >= (1L << 31) || a.ntiles >= (1L << 31) || (size_t)2 * pbuf >= ((size_t)1 << 18) ||
      fixed + 4 * (size_t)a.stage_bytes > (size_t)kMaxSmem) {
    jaf::set_error("jaf_convlstm_step_grouped: cell does not fit the operand-swapped plan (Cin=%d Ch=%d H=%d W=%d B=%d)", Cin, Ch,
                   H, W, B);
    return JAF_ERR_UNSUPPORTED;
  }
  size_t ring = ((size_t)kMaxSmem - fixed) / a.stage_bytes;
  if (ring > (size_t)kMaxRingP) ring = kMaxRingP;
  a.ring = (int)ring;
  smem = fixed + ring * a.stage_bytes;
  grid = a.ntiles < sm_count ? a.ntiles : sm_count;
  return JAF_OK;
}

int jaf_convlstm_grouped_supported(int G, int B, int Cin, int Ch, int H, int W) {
  if (G <= 0 || B <= 0 || H <= 0 || W <= 0 || !grouped_shape_ok(Cin, Ch)) return 0;
  int sms = jaf::sm_count(jaf::current_device());
  if (sms <= 0) sms = 148;  // no device yet: plan for a B200
  GArgs a;
  int threads;
  size_t smem;
  long grid;
  if (grouped_swapped(Cin, Ch)) return plan_grouped_p(G, B, Cin, Ch, H, W, sms, a, smem, grid) == JAF_OK ? 1 : 0;
  return plan_grouped(G, B, Cin, Ch, H, W, sms, a, threads, smem, grid) == JAF_OK ? 1 : 0;
}

int jaf_convlstm_step_grouped(const float* x, const float* h, const float* c, const void* wpack, const float* bias,
                              int G, int B, int Cin, int Ch, int H, int W, float* h_out, float* c_out, void* stream) {
  JAF_REQUIRE(x && h && c && wpack && h_out && c_out, "null pointer");
  JAF_REQUIRE(G > 0 && B > 0 && H > 0 && W > 0, "bad sizes");
  JAF_REQUIRE(grouped_shape_ok(Cin, Ch), "Ch must be a multiple of 4 and at most 128");
  JAF_REQUIRE(((uintptr_t)wpack & 15) == 0, "wpack must be 16-byte aligned");
  static jaf::PerDeviceOnce attr_once;  // the shared-memory attribute is per device
  const int dev = jaf::current_device();
  const int sm_count = jaf::sm_count(dev);
  JAF_REQUIRE(sm_count > 0, "no current CUDA device");
  if (!attr_once.done(dev)) {
    JAF_CUDA(cudaFuncSetAttribute(k_convlstm_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_once.mark(dev);
  }
  GArgs a;
  a.x = x; a.h = h; a.c = c; a.bias = bias; a.wpack = static_cast<const uint8_t*>(wpack);
  a.h_out = h_out; a.c_out = c_out;
  int threads;
  size_t smem;
  long grid;
  if (grouped_swapped(Cin, Ch)) {
    static jaf::PerDeviceOnce attr_once_p;
    if (!attr_once_p.done(dev)) {
      JAF_CUDA(cudaFuncSetAttribute(k_convlstm_grouped_p, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
      attr_once_p.mark(dev);
    }
    const int stt = plan_grouped_p(G, B, Cin, Ch, H, W, sm_count, a, smem, grid);
    if (stt != JAF_OK) return stt;
    k_convlstm_grouped_p<<<(unsigned)grid, kThreadsP, smem, jaf::as_stream(stream)>>>(a);
    return jaf::finish_launch("k_convlstm_grouped_p");
  }
  const int st = plan_grouped(G, B, Cin, Ch, H, W, sm_count, a, threads, smem, grid);
  if (st != JAF_OK) return st;
  k_convlstm_grouped<<<(unsigned)grid, threads, smem, jaf::as_stream(stream)>>>(a);
  return jaf::finish_launch("k_convlstm_grouped");
}

int jaf_convlstm_sequence_grouped(const float* x_seq, const float* h0, const float* c0, const void* wpack, const float* bias,
                                  int G, int B, int T, int Cin, int Ch, int H, int W, float* h_seq, float* c_last,
                                  float* c_tmp, void* stream) {
  JAF_REQUIRE(x_seq && h0 && c0 && wpack && h_seq && c_last, "null pointer");
  JAF_REQUIRE(T == 1 || c_tmp, "c_tmp [G,B,Ch,H,W] is required for T > 1");
  JAF_REQUIRE(G > 0 && B > 0 && T > 0 && H > 0 && W > 0, "bad sizes");
  JAF_REQUIRE(grouped_shape_ok(Cin, Ch), "Ch must be a multiple of 4 and at most 128");
  JAF_REQUIRE(((uintptr_t)wpack & 15) == 0, "wpack must be 16-byte aligned");
  static jaf::PerDeviceOnce attr_once;
  const int dev = jaf::current_device();
  const int sm_count = jaf::sm_count(dev);
  JAF_REQUIRE(sm_count > 0, "no current CUDA device");
  if (!attr_once.done(dev)) {
    JAF_CUDA(cudaFuncSetAttribute(k_convlstm_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_once.mark(dev);
  }
  GArgs a;
  a.bias = bias; a.wpack = static_cast<const uint8_t*>(wpack);
  int threads = kThreadsP;
  size_t smem;
  long grid;
  const bool swapped = grouped_swapped(Cin, Ch);
  if (swapped) {
    static jaf::PerDeviceOnce attr_once_p;
    if (!attr_once_p.done(dev)) {
      JAF_CUDA(cudaFuncSetAttribute(k_convlstm_grouped_p, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
      attr_once_p.mark(dev);
    }
  }
  const int st = swapped ? plan_grouped_p(G, B, Cin, Ch, H, W, sm_count, a, smem, grid)
                         : plan_grouped(G, B, Cin, Ch, H, W, sm_count, a, threads, smem, grid);
  if (st != JAF_OK) return st;
  const long HW = (long)H * W;
  // the recurrence of src/convLSTM.py:131-134 in one call: step t reads x[:, :, t] and h[:, :, t-1] straight from the
  // sequence tensors (strided views, no slicing copies) and writes h[:, :, t]; the cell state ping-pongs between two
  // buffers so that the last step lands in c_last
  float* cbuf[2] = {(T & 1) ? c_last : c_tmp, (T & 1) ? c_tmp : c_last};
  for (int t = 0; t < T; ++t) {
    a.x = x_seq + (long)t * Cin * HW;
    a.xs = (long)T * Cin * HW;
    a.h = t == 0 ? h0 : h_seq + (long)(t - 1) * Ch * HW;
    a.hs = t == 0 ? (long)Ch * HW : (long)T * Ch * HW;
    a.c = t == 0 ? c0 : cbuf[(t - 1) & 1];
    a.h_out = h_seq + (long)t * Ch * HW;
    a.hos = (long)T * Ch * HW;
    a.c_out = cbuf[t & 1];
    if (swapped) k_convlstm_grouped_p<<<(unsigned)grid, kThreadsP, smem, jaf::as_stream(stream)>>>(a);
    else k_convlstm_grouped<<<(unsigned)grid, threads, smem, jaf::as_stream(stream)>>>(a);
  }
  return jaf::finish_launch("k_convlstm_grouped (sequence)", T);
}

}  // extern "C"
