// jaf_warp_fuse_host / jaf_warp_fuse_from_poses_host: the row-F operation for callers whose tensors live in HOST
// memory (the reference's per-frame loop hands frames over from DataLoader workers, test/conv_pro_test.py:168-195,
// and reads every result back, :282-302).  Target frames are streamed through two device-side slots: while slot s
// computes, slot s^1 uploads the next chunk and downloads the previous result, on three streams tied together with
// events.  Reference sets (and, for the pose-driven entry point, the reference poses) addressed through ref_index are
// uploaded once, right before the first chunk that uses them, and stay resident for the whole call.
//
// One pipeline PER DEVICE (streams, events and grow-only staging buffers are cached per device, so steady-state calls
// do no cudaMalloc): a process that drives several GPUs — the reference's nn.DataParallel(float_estimate()),
// test/conv_pro_test.py:140-141 — gets an independent pipeline on each.  Blocking: returns after the last D2H copy has
// landed; on an error every stream of the pipeline is drained before returning, so no copy to or from the caller's
// buffers is still in flight.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return JAF_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return jaf::cuda_status(e, "cudaMalloc(staging)");
    cap = bytes;
    return JAF_OK;
  }
};

struct Slot {
  DevBuf grid, logits, vis, fim, mask, fake, conf, rgb, feat, out_rgb, out_feat, warped;
  DevBuf tgt_cam, tgt_verts, raster_ws;  // pose-driven entry point
  size_t clean_keys = 0;                 // leading z-buffer keys of raster_ws known to be empty (self-cleaning kernel)
  cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr;
};

struct Pipeline {
  std::mutex mu;
  Slot slot[2];
  DevBuf ref_rgb, ref_feat, ref_index, src_cam, src_verts, faces_idx;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  bool ready = false;
  int init() {
    if (ready) return JAF_OK;
    JAF_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    JAF_CUDA(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
    JAF_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (auto& s : slot) {
      JAF_CUDA(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
      JAF_CUDA(cudaEventCreateWithFlags(&s.computed, cudaEventDisableTiming));
      JAF_CUDA(cudaEventCreateWithFlags(&s.downloaded, cudaEventDisableTiming));
    }
    ready = true;
    return JAF_OK;
  }
  void drain() {
    if (!ready) return;
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_k);
    cudaStreamSynchronize(s_out);
  }
};

Pipeline g_pipes[jaf::kMaxDevices];

// Drains the pipeline's streams on every exit path (error returns included) while the caller's buffers are still valid.
struct DrainGuard {
  Pipeline& p;
  ~DrainGuard() { p.drain(); }
};

inline size_t esz(int dtype) { return dtype == JAF_DTYPE_BF16 ? 2 : 4; }

#define JAF_TRY(expr)              \
  do {                             \
    int st__ = (expr);             \
    if (st__ != JAF_OK) return st__; \
  } while (0)

int h2d(DevBuf& d, const void* src, size_t bytes, cudaStream_t st) {
  JAF_TRY(d.ensure(bytes));
  JAF_CUDA(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, st));
  return JAF_OK;
}

// hq == nullptr: flows come from the caller (hp->grid).  hq != nullptr: flows come from the poses (hq->* are HOST
// pointers except hq->workspace, which is ignored: the pipeline owns the raster workspaces).
int run_pipeline(const JafWarpFuseParams* hp, const JafPoseFlowParams* hq, int frames_per_chunk, void* out_feat_device) {
  JAF_REQUIRE(hp != nullptr, "null params");
  JAF_REQUIRE(hp->B >= 0 && hp->K >= 1 && hp->H > 0 && hp->W > 0 && hp->Hs > 0 && hp->Ws > 0, "bad sizes");
  const bool poses = hq != nullptr;
  JAF_REQUIRE(poses || hp->grid != nullptr, "grid is required");
  JAF_REQUIRE(!poses || (hq->tgt_cam && hq->tgt_verts && hq->src_cam && hq->src_verts && hq->faces_idx && hq->V > 0 && hq->F >= 0),
              "null / bad pose input");
  JAF_REQUIRE(!poses || (!hq->T && !hq->fim), "the host entry point does not return T / fim");
  const bool want_rgb = hp->rgb && (hp->out_rgb || hp->warped_rgb);
  const bool feat_to_device = out_feat_device != nullptr;
  const bool want_feat = hp->feat && (hp->out_feat || feat_to_device) && hp->C > 0;
  JAF_REQUIRE(want_rgb || want_feat, "nothing to do");
  JAF_REQUIRE(!poses || want_feat, "the pose-driven entry point needs features");
  if (hp->B == 0) return JAF_OK;

  const int dev = jaf::current_device();
  JAF_REQUIRE(dev >= 0 && dev < jaf::kMaxDevices, "no current CUDA device");
  Pipeline& P = g_pipes[dev];
  std::lock_guard<std::mutex> lock(P.mu);
  JAF_TRY(P.init());
  DrainGuard guard{P};

  const int B = hp->B, K = hp->K;
  const size_t HW = (size_t)hp->H * hp->W, HWs = (size_t)hp->Hs * hp->Ws;
  const size_t fe = esz(hp->feat_dtype);
  const size_t rgb_set = (size_t)K * 3 * HWs * 4;          // one reference set, RGB
  const size_t feat_set = (size_t)K * hp->C * HWs * fe;    // one reference set, features
  const size_t pose_set = poses ? (size_t)K * hq->V * 3 * 4 : 0;  // the K reference poses of one set
  // bytes a target frame moves through a slot (resident reference sets addressed by ref_index do not count)
  const size_t per_frame = (poses ? (size_t)hq->V * 12 + 8 * HW : K * HW * 8) + K * HW * 8 + HW * 64 +
                           (want_feat ? (size_t)hp->C * HW * fe : 0) +
                           (hp->ref_index ? 0 : (want_rgb ? rgb_set : 0) + (want_feat ? feat_set : 0) + pose_set);
  int n = frames_per_chunk;
  if (n <= 0) {
    // ~96 MB per slot: the transfer of a chunk (~2 ms over PCIe 5) dwarfs launch overheads, and a short first chunk
    // keeps the pipeline prologue (nothing overlaps the first upload) small
    const size_t fit = ((size_t)96 << 20) / (per_frame > 0 ? per_frame : 1);
    n = fit < 1 ? 1 : (fit > 64 ? 64 : (int)fit);
  }
  if (n > B) n = B;

  // resident reference sets (ref_index given): each is uploaded once
  std::vector<char> ref_done;
  const float* d_ref_rgb = nullptr;
  const void* d_ref_feat = nullptr;
  if (hp->ref_index) {
    int R = 0;
    for (int b = 0; b < B; ++b) {
      JAF_REQUIRE(hp->ref_index[b] >= 0, "negative ref_index");
      if (hp->ref_index[b] + 1 > R) R = hp->ref_index[b] + 1;
    }
    // reference sets go up lazily, right before the first chunk that uses them: the pipeline starts after ONE set has
    // arrived instead of all R (for a 30-frame video per set that is most of the prologue)
    if (want_rgb) {
      JAF_TRY(P.ref_rgb.ensure(rgb_set * R));
      d_ref_rgb = static_cast<const float*>(P.ref_rgb.p);
    }
    if (want_feat) {
      JAF_TRY(P.ref_feat.ensure(feat_set * R));
      d_ref_feat = P.ref_feat.p;
    }
    if (poses) {
      JAF_TRY(P.src_cam.ensure((size_t)R * K * 12));
      JAF_TRY(P.src_verts.ensure(pose_set * R));
    }
    JAF_TRY(h2d(P.ref_index, hp->ref_index, sizeof(int32_t) * B, P.s_in));
    ref_done.assign((size_t)R, 0);
  }
  if (poses) JAF_TRY(h2d(P.faces_idx, hq->faces_idx, (size_t)hq->F * 3 * 4, P.s_in));

  const int nchunks = (B + n - 1) / n;
  for (int ci = 0; ci < nchunks; ++ci) {
    Slot& S = P.slot[ci & 1];
    const int b0 = ci * n;
    const int nb = (b0 + n <= B) ? n : (B - b0);
    // the slot's previous result must have left before its inputs/outputs are overwritten
    if (ci >= 2) JAF_CUDA(cudaStreamWaitEvent(P.s_in, S.downloaded, 0));
    // ---- upload
    if (hp->ref_index) {
      for (int bb = b0; bb < b0 + nb; ++bb) {
        const int rset = hp->ref_index[bb];
        if (ref_done[rset]) continue;
        ref_done[rset] = 1;
        if (want_rgb)
          JAF_CUDA(cudaMemcpyAsync(static_cast<char*>(P.ref_rgb.p) + rgb_set * rset, hp->rgb + (size_t)rset * K * 3 * HWs,
                                   rgb_set, cudaMemcpyHostToDevice, P.s_in));
        if (want_feat)
          JAF_CUDA(cudaMemcpyAsync(static_cast<char*>(P.ref_feat.p) + feat_set * rset,
                                   static_cast<const char*>(hp->feat) + feat_set * rset, feat_set,
                                   cudaMemcpyHostToDevice, P.s_in));
        if (poses) {
          JAF_CUDA(cudaMemcpyAsync(static_cast<char*>(P.src_cam.p) + (size_t)rset * K * 12, hq->src_cam + (size_t)rset * K * 3,
                                   (size_t)K * 12, cudaMemcpyHostToDevice, P.s_in));
          JAF_CUDA(cudaMemcpyAsync(static_cast<char*>(P.src_verts.p) + pose_set * rset,
                                   hq->src_verts + (size_t)rset * K * hq->V * 3, pose_set, cudaMemcpyHostToDevice, P.s_in));
        }
      }
    }
    if (poses) {
      JAF_TRY(h2d(S.tgt_cam, hq->tgt_cam + (size_t)b0 * 3, (size_t)nb * 12, P.s_in));
      JAF_TRY(h2d(S.tgt_verts, hq->tgt_verts + (size_t)b0 * hq->V * 3, (size_t)nb * hq->V * 12, P.s_in));
      if (jaf_raster_workspace_bytes(nb, hp->H) > S.raster_ws.cap) S.clean_keys = 0;  // a new buffer holds garbage
      JAF_TRY(S.raster_ws.ensure(jaf_raster_workspace_bytes(nb, hp->H)));
    } else {
      JAF_TRY(h2d(S.grid, hp->grid + (size_t)b0 * K * HW * 2, (size_t)nb * K * HW * 8, P.s_in));
      if (hp->vis) JAF_TRY(h2d(S.vis, hp->vis + (size_t)b0 * K * HW, (size_t)nb * K * HW * 4, P.s_in));
      if (hp->fim) JAF_TRY(h2d(S.fim, hp->fim + (size_t)b0 * HW, (size_t)nb * HW * 4, P.s_in));
    }
    if (hp->logits) JAF_TRY(h2d(S.logits, hp->logits + (size_t)b0 * K * HW, (size_t)nb * K * HW * 4, P.s_in));
    if (hp->tgt_mask)
      JAF_TRY(h2d(S.mask, hp->tgt_mask + (size_t)b0 * hp->mask_c * HW, (size_t)nb * hp->mask_c * HW * 4, P.s_in));
    if (hp->fake && hp->conf) {
      JAF_TRY(h2d(S.fake, hp->fake + (size_t)b0 * 3 * HW, (size_t)nb * 3 * HW * 4, P.s_in));
      JAF_TRY(h2d(S.conf, hp->conf + (size_t)b0 * HW, (size_t)nb * HW * 4, P.s_in));
    }
    if (!hp->ref_index) {
      if (want_rgb) JAF_TRY(h2d(S.rgb, hp->rgb + (size_t)b0 * K * 3 * HWs, rgb_set * nb, P.s_in));
      if (want_feat)
        JAF_TRY(h2d(S.feat, static_cast<const char*>(hp->feat) + feat_set * b0, feat_set * nb, P.s_in));
      if (poses) {  // one set of reference poses per target frame, streamed with it (the raster slot buffers are reused)
        JAF_TRY(P.src_cam.ensure((size_t)2 * n * K * 12));
        JAF_TRY(P.src_verts.ensure(pose_set * 2 * n));
        JAF_CUDA(cudaMemcpyAsync(static_cast<char*>(P.src_cam.p) + (size_t)(ci & 1) * n * K * 12, hq->src_cam + (size_t)b0 * K * 3,
                                 (size_t)nb * K * 12, cudaMemcpyHostToDevice, P.s_in));
        JAF_CUDA(cudaMemcpyAsync(static_cast<char*>(P.src_verts.p) + pose_set * (size_t)(ci & 1) * n,
                                 hq->src_verts + (size_t)b0 * K * hq->V * 3, pose_set * nb, cudaMemcpyHostToDevice, P.s_in));
      }
    }
    JAF_CUDA(cudaEventRecord(S.uploaded, P.s_in));
    // ---- compute
    JAF_CUDA(cudaStreamWaitEvent(P.s_k, S.uploaded, 0));
    JafWarpFuseParams d = *hp;
    d.B = nb;
    d.grid = poses ? nullptr : static_cast<const float*>(S.grid.p);
    d.logits = hp->logits ? static_cast<const float*>(S.logits.p) : nullptr;
    d.vis = (!poses && hp->vis) ? static_cast<const float*>(S.vis.p) : nullptr;
    d.fim = (!poses && hp->fim) ? static_cast<const int32_t*>(S.fim.p) : nullptr;
    d.tgt_mask = hp->tgt_mask ? static_cast<const float*>(S.mask.p) : nullptr;
    d.fake = (hp->fake && hp->conf) ? static_cast<const float*>(S.fake.p) : nullptr;
    d.conf = (hp->fake && hp->conf) ? static_cast<const float*>(S.conf.p) : nullptr;
    if (hp->ref_index) {
      d.rgb = want_rgb ? d_ref_rgb : nullptr;
      d.feat = want_feat ? d_ref_feat : nullptr;
      d.ref_index = static_cast<const int32_t*>(P.ref_index.p) + b0;
    } else {
      d.rgb = want_rgb ? static_cast<const float*>(S.rgb.p) : nullptr;
      d.feat = want_feat ? S.feat.p : nullptr;
      d.ref_index = nullptr;
    }
    d.out_rgb = nullptr;
    d.out_feat = nullptr;
    d.warped_rgb = nullptr;
    if (want_rgb && hp->out_rgb) {
      JAF_TRY(S.out_rgb.ensure((size_t)nb * 3 * HW * 4));
      d.out_rgb = static_cast<float*>(S.out_rgb.p);
    }
    if (want_rgb && hp->warped_rgb) {
      JAF_TRY(S.warped.ensure((size_t)nb * K * 3 * HW * 4));
      d.warped_rgb = static_cast<float*>(S.warped.p);
    }
    if (want_feat) {
      if (feat_to_device) {  // device-resident feature output: the kernel writes the caller's buffer, nothing is staged
        d.out_feat = static_cast<char*>(out_feat_device) + (size_t)b0 * hp->C * HW * fe;
      } else {
        JAF_TRY(S.out_feat.ensure((size_t)nb * hp->C * HW * fe));
        d.out_feat = S.out_feat.p;
      }
    }
    d.stream = P.s_k;
    if (poses) {
      JafPoseFlowParams dq = *hq;
      dq.tgt_cam = static_cast<const float*>(S.tgt_cam.p);
      dq.tgt_verts = static_cast<const float*>(S.tgt_verts.p);
      dq.faces_idx = static_cast<const int32_t*>(P.faces_idx.p);
      if (hp->ref_index) {
        dq.src_cam = static_cast<const float*>(P.src_cam.p);
        dq.src_verts = static_cast<const float*>(P.src_verts.p);
      } else {
        dq.src_cam = reinterpret_cast<const float*>(static_cast<char*>(P.src_cam.p) + (size_t)(ci & 1) * n * K * 12);
        dq.src_verts = reinterpret_cast<const float*>(static_cast<char*>(P.src_verts.p) + pose_set * (size_t)(ci & 1) * n);
      }
      dq.T = nullptr;
      dq.fim = nullptr;
      dq.workspace = S.raster_ws.p;
      // the pipeline owns this workspace: the fused kernel leaves the keys empty and the next chunk skips the clear
      const size_t need = (size_t)nb * hp->H * hp->H;
      dq.flags = JAF_POSES_LEAVE_CLEAN | (S.clean_keys >= need ? JAF_POSES_KEYS_CLEAN : 0);
      S.clean_keys = 0;  // not clean again until this call has succeeded
      JAF_TRY(jaf_warp_fuse_from_poses(&d, &dq));
      S.clean_keys = need;
    } else {
      JAF_TRY(jaf_warp_fuse(&d));
    }
    JAF_CUDA(cudaEventRecord(S.computed, P.s_k));
    // ---- download
    JAF_CUDA(cudaStreamWaitEvent(P.s_out, S.computed, 0));
    if (d.out_rgb)
      JAF_CUDA(cudaMemcpyAsync(hp->out_rgb + (size_t)b0 * 3 * HW, d.out_rgb, (size_t)nb * 3 * HW * 4,
                               cudaMemcpyDeviceToHost, P.s_out));
    if (d.warped_rgb)
      JAF_CUDA(cudaMemcpyAsync(hp->warped_rgb + (size_t)b0 * K * 3 * HW, d.warped_rgb, (size_t)nb * K * 3 * HW * 4,
                               cudaMemcpyDeviceToHost, P.s_out));
    if (want_feat && !feat_to_device)
      JAF_CUDA(cudaMemcpyAsync(static_cast<char*>(hp->out_feat) + (size_t)b0 * hp->C * HW * fe, d.out_feat,
                               (size_t)nb * hp->C * HW * fe, cudaMemcpyDeviceToHost, P.s_out));
    JAF_CUDA(cudaEventRecord(S.downloaded, P.s_out));
  }
  JAF_CUDA(cudaStreamSynchronize(P.s_out));
  JAF_CUDA(cudaStreamSynchronize(P.s_k));
  JAF_CUDA(cudaStreamSynchronize(P.s_in));
  return JAF_OK;
}

}  // namespace

extern "C" int jaf_warp_fuse_host(const JafWarpFuseParams* hp, int frames_per_chunk) {
  return run_pipeline(hp, nullptr, frames_per_chunk, nullptr);
}

extern "C" int jaf_warp_fuse_from_poses_host(const JafWarpFuseParams* hp, const JafPoseFlowParams* hq, int frames_per_chunk,
                                             void* out_feat_device) {
  JAF_REQUIRE(hq != nullptr, "null pose params");
  return run_pipeline(hp, hq, frames_per_chunk, out_feat_device);
}
