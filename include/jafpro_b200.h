/*
 * jafpro_b200 — C ABI of the B200-native appearance warp-and-fuse path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every
 * entry point names the reference interface it replaces (paths relative to the
 * Larry-u/JAFPro tree; "NR" = third_party/neural_renderer/neural_renderer).
 * INTEGRATION.md shows the reference-side binding for each one.
 *
 * Conventions (kept from the reference's extension, NR/cuda/rasterize_cuda.cpp:70-95):
 *   - the CALLER allocates every output and workspace; functions fill them in place;
 *   - all pointers are DEVICE pointers of the current CUDA device unless the
 *     function name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which
 *     is what the reference launches on, rasterize_cuda_kernel.cu:616);
 *   - tensors are dense, row-major, in the shapes written next to each argument.
 * Unlike the reference (which only printf()s launch failures,
 * rasterize_cuda_kernel.cu:624-626,647-649) every function returns a status:
 *   0 success, <0 failure; jaf_last_error() gives a thread-local message.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef JAFPRO_B200_H_
#define JAFPRO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JAF_OK 0
#define JAF_ERR_INVALID (-1)     /* bad argument (NULL pointer, unsupported shape ...) */
#define JAF_ERR_CUDA (-2)        /* a CUDA call or kernel launch failed                */
#define JAF_ERR_UNSUPPORTED (-3) /* valid request this build cannot serve              */

#define JAF_LAYOUT_PLANAR 0 /* [.., C, H, W]  (the reference's NCHW)          */
#define JAF_LAYOUT_NHWC 1   /* [.., H, W, C]  (channels-last)                 */
#define JAF_DTYPE_F32 0
#define JAF_DTYPE_BF16 1

int jaf_version(void);
const char* jaf_last_error(void);
/* Number of kernels this library has launched in the calling process (monotonic). */
uint64_t jaf_launch_count(void);
/* Measurement aids (no reference counterpart): the kernel variant the calling thread launched last through
 * jaf_warp_fuse / jaf_warp_fuse_from_maps, e.g. "k_warp_fuse_nhwc_wide<K=4,MINB=4,SKIP=0,RGBM=1>", and the effective
 * values of the JAF_* tuning knobs (environment, read once per process) as "NAME=value ..." written into buf
 * (returns the length needed). */
const char* jaf_last_kernel(void);
int jaf_tuning_info(char* buf, int n);

/* ---------------------------------------------------------------------------------
 * a1-a3  projection + y flip + look_at + face gather
 * replaces: orthographic_proj_withz_idrot (src/nmr.py:10-28), `proj_verts[:,:,1] *= -1`
 *           (src/nmr.py:271), nr.look_at(proj, eye) (NR/look_at.py:6-62; identity rotation
 *           for SMPLRenderer's eye, src/nmr.py:177), nr.vertices_to_faces
 *           (NR/vertices_to_faces.py:4-22).
 * cam [B,3] (s,tx,ty); verts [B,V,3]; faces_idx [F,3] int32 (shared by the batch,
 * src/nmr.py:266); eye_z = float32(-(1/tan(30deg)+1)); faces_xyz out [B,F,3,3].
 * --------------------------------------------------------------------------------- */
int jaf_project_gather(const float* cam, const float* verts, const int32_t* faces_idx, int B, int V,
                       int F, float eye_z, float* faces_xyz, void* stream);

/* ---------------------------------------------------------------------------------
 * a4-a6  face-index + barycentric-weight rasteriser
 * replaces: rasterize_cuda.forward_face_index_map (NR/cuda/rasterize_cuda.cpp:70-95 ->
 *           rasterize_cuda_kernel.cu:24-169,596-651) together with the Python around it:
 *           the output fills of NR/rasterize.py:50-52 and the row flips of :334-338
 *           (flip_rows=1), i.e. nr.rasterize_face_index_map_and_weight_map
 *           (NR/rasterize.py:543-571) with anti_aliasing=False.
 * faces_xyz [B,F,3,3]; fim out [B,S,S] int32 (-1 = background); wim out [B,S,S,3];
 * depth out [B,S,S] or NULL (background = far); workspace: jaf_raster_workspace_bytes().
 * fim is bit-exact with the reference kernels; wim/depth reproduce their fp32 operation
 * order (including nvcc's FMA contraction) and are bit-exact as well.
 * --------------------------------------------------------------------------------- */
size_t jaf_raster_workspace_bytes(int B, int image_size);
int jaf_raster_fim_wim(const float* faces_xyz, int B, int F, int image_size, float near_, float far_,
                       int flip_rows, int32_t* fim, float* wim, float* depth, void* workspace,
                       void* stream);

/* ---------------------------------------------------------------------------------
 * a4-a6 with the extension's OWN contract: rasterize_cuda.forward_face_index_map(faces, face_index_map, weight_map,
 * depth_map, face_inv_map, faces_inv, image_size, near, far, return_rgb, return_alpha, return_depth)
 * (NR/cuda/rasterize_cuda.cpp:70-95,194-200 -> rasterize_cuda_kernel.cu:24-169,596-651).
 * Outputs are PRE-FILLED by the caller (fim -1, wim 0, depth far, faces_inv 0: NR/rasterize.py:50-52,164) and filled
 * in place: only pixels covered by a face are written (rasterize_cuda_kernel.cu:156-168), rows are NOT flipped (the
 * flip is Python's, NR/rasterize.py:334-338), faces_inv [B,F,3,3] (nullable) receives kernel_1's per-face inverse
 * matrices for front faces, face_inv_map [B,S,S,3,3] is written only when return_depth != 0 (else it may be the
 * reference's 1-element dummy or NULL).  return_rgb / return_alpha do not reach the kernels (:105-169) and are
 * not parameters here.  Every written value is bit-identical to the reference kernels'.
 * --------------------------------------------------------------------------------- */
int jaf_forward_face_index_map(const float* faces, int32_t* face_index_map, float* weight_map, float* depth_map,
                               float* face_inv_map, float* faces_inv, int B, int F, int image_size, float near_,
                               float far_, int return_depth, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------
 * a7  SMPLRenderer.render_fim_wim (src/nmr.py:263-278) in one call: a1-a3 + a4-a6.
 * faces_xyz out [B,F,3,3] may be NULL when the caller does not need `faces`
 * (then the [B,13776,3,3] tensor never touches HBM).
 * --------------------------------------------------------------------------------- */
int jaf_render_fim_wim(const float* cam, const float* verts, const int32_t* faces_idx, int B, int V,
                       int F, int image_size, float eye_z, float near_, float far_, float* faces_xyz,
                       int32_t* fim, float* wim, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------
 * a9  barycentric flow compose
 * replaces: SMPLRenderer.cal_bc_transform (src/nmr.py:617-659).
 * src_pts: per-face source coordinates, `stride` floats per vertex: 2 for the
 * [B,F,3,2] tensor of the reference signature, 3 to read x,y straight from a
 * faces_xyz [B,F,3,3] tensor.  negate_y=1 applies `src_f2verts[:,:,:,1] *= -1`
 * (src/cal_flow.py:31) on the fly.  fim [B,H,W] int32; wim [B,H,W,3]; T out [B,H,W,2],
 * -2 where fim == -1 (src/nmr.py:627).
 * --------------------------------------------------------------------------------- */
int jaf_flow_compose(const float* src_pts, int stride, int negate_y, const int32_t* fim,
                     const float* wim, int B, int F, int H, int W, float* T, void* stream);

/* ---------------------------------------------------------------------------------
 * a8  float_estimate.cal_flow (src/cal_flow.py:28-35) in one call.  The source pose is
 * projected but NOT rasterised (its fim/wim are dead in the reference, cal_flow.py:29).
 * T out [B,S,S,2]; fim/wim out may be NULL (then they never touch HBM).
 * workspace: jaf_raster_workspace_bytes(B, S).
 * --------------------------------------------------------------------------------- */
int jaf_cal_flow(const float* src_cam, const float* src_verts, const float* tgt_cam,
                 const float* tgt_verts, const int32_t* faces_idx, int B, int V, int F,
                 int image_size, float eye_z, float near_, float far_, float* T, int32_t* fim,
                 float* wim, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------
 * a8 for K references: the transfer flows from K source poses into ONE target pose per frame.
 * The reference calls float_estimate.cal_flow once per (source, target) pair and rasterises the
 * target every time (src/cal_flow.py:33); here the target is rasterised once and composed K times.
 * src_cam [B,K,3]; src_verts [B,K,V,3]; tgt_cam [B,3]; tgt_verts [B,V,3];
 * T out [B,K,S,S,2] (the `grid` of jaf_warp_fuse); fim/wim out [B,S,S]/[B,S,S,3] may be NULL.
 * T[b,k] is bit-identical to jaf_cal_flow(src[b,k], tgt[b]).
 * --------------------------------------------------------------------------------- */
int jaf_cal_flow_multi(const float* src_cam, const float* src_verts, const float* tgt_cam,
                       const float* tgt_verts, const int32_t* faces_idx, int B, int K, int V, int F,
                       int image_size, float eye_z, float near_, float far_, float* T, int32_t* fim,
                       float* wim, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------
 * row F  fused bilinear backward warp of K references + visibility/softmax fusion
 * replaces, in ONE pass over every input byte:
 *   float_estimate.warp_image = F.grid_sample(src, flow, padding_mode='border')
 *       (src/cal_flow.py:37-39; feature-map warps src/crn_model.py:463-566),
 *   tsf * tgt_smpl_mask and the confidence blend (src/flow_net.py:91,98),
 *   softmax-over-K weighting and the sum over K (src/networks.py:1230-1244,1264-1286).
 * Per target pixel p of frame b (fp32 arithmetic, bf16 only as storage):
 *   alpha = softmax_k(logits[b,k,p])            (uniform 1/K when logits == NULL)
 *   v_k   = vis[b,k,p]  |  (fim[b,p] != -1)     (1 when both NULL)
 *   fused = sum_k (alpha_k * v_k) * bilinear_border(ref[r,k], grid[b,k,p]);  r = ref_index[b] | b
 *   fused *= tgt_mask[b,c|0,p]                  (when tgt_mask != NULL)
 *   out_rgb = fake*conf + fused*(1-conf)        (RGB only, when fake and conf != NULL)
 * K == 1 without logits/vis/fim reduces exactly to warp_image(src, T) * mask.
 * --------------------------------------------------------------------------------- */
typedef struct JafWarpFuseParams {
  int32_t B, K, H, W;    /* target frames, references per frame, output size        */
  int32_t Hs, Ws;        /* reference (source) size                                 */
  int32_t C;             /* feature channels (0 = no feature tensor)                */
  int32_t align_corners; /* torch 1.2 (pinned by the reference) == 1, torch >= 1.3 default == 0 */
  int32_t feat_layout;   /* JAF_LAYOUT_*                                            */
  int32_t feat_dtype;    /* JAF_DTYPE_*                                             */
  int32_t mask_c;        /* channels of tgt_mask: 1 or 3                            */
  int32_t reserved;
  const float* rgb;         /* [R,K,3,Hs,Ws] f32 planar, or NULL                    */
  const void* feat;         /* [R,K,C,Hs,Ws] | [R,K,Hs,Ws,C], or NULL               */
  const int32_t* ref_index; /* [B] reference-set index per target frame, or NULL    */
  const float* grid;        /* [B,K,H,W,2] transfer flows, (x,y) in NDC             */
  const float* logits;      /* [B,K,H,W] or NULL                                    */
  const float* vis;         /* [B,K,H,W] or NULL                                    */
  const int32_t* fim;       /* [B,H,W] or NULL (default visibility)                 */
  const float* tgt_mask;    /* [B,mask_c,H,W] or NULL                               */
  const float* fake;        /* [B,3,H,W] or NULL                                    */
  const float* conf;        /* [B,1,H,W] or NULL                                    */
  float* out_rgb;           /* [B,3,H,W] or NULL                                    */
  void* out_feat;           /* [B,C,H,W] | [B,H,W,C] (layout/dtype of feat) or NULL */
  float* warped_rgb;        /* [B,K,3,H,W] per-reference warps, or NULL             */
  void* stream;
} JafWarpFuseParams;

int jaf_warp_fuse(const JafWarpFuseParams* p);

/* ---------------------------------------------------------------------------------
 * rows a7-a12 in ONE pass from the poses: the transfer flows of the K reference poses into every
 * target pose are composed inside the warp kernel, per tile, in shared memory — the flow tensor never round-trips HBM.
 * replaces, per target frame: K x float_estimate.cal_flow (src/cal_flow.py:28-35: render_fim_wim of the target,
 * src/nmr.py:263-278, + cal_bc_transform, :617-659) followed by the row-F operation above with the default visibility
 * "the target pixel is on the body" (fim != -1, the -2 sentinel of src/nmr.py:627).
 * `p` as for jaf_warp_fuse except: grid, vis and fim are ignored (they are what this call computes), H == W = the raster
 * size.  Source poses are indexed like the reference sets (r = ref_index[b] | b).  RGB-only calls (feat == NULL: the
 * reference's own per-frame warp_image + mask + blend chain) are served as well.
 * T / fim (optional outputs) are bit-identical to jaf_cal_flow_multi's; out_rgb / out_feat are bit-identical to
 * jaf_cal_flow_multi + jaf_warp_fuse(fim).  Served shapes: jaf_warp_fuse_from_poses_supported() (C = 64 channels-last
 * bf16 or C = 0 = RGB only, K <= 8); others return JAF_ERR_UNSUPPORTED and take the two-call path.
 * --------------------------------------------------------------------------------- */
/* The z-buffer keys live in the caller's workspace.  A caller that OWNS the workspace between calls can save the
 * per-call clear (8 B per target pixel): JAF_POSES_LEAVE_CLEAN makes the fused kernel reset every key it consumed, and
 * JAF_POSES_KEYS_CLEAN on the next call asserts that the first B*S*S keys are still empty (nothing else wrote the
 * workspace since a LEAVE_CLEAN call of at least that size). */
#define JAF_POSES_KEYS_CLEAN 1
#define JAF_POSES_LEAVE_CLEAN 2
typedef struct JafPoseFlowParams {
  const float* tgt_cam;      /* [B,3] (s,tx,ty)                                      */
  const float* tgt_verts;    /* [B,V,3]                                              */
  const float* src_cam;      /* [R,K,3]                                              */
  const float* src_verts;    /* [R,K,V,3]                                            */
  const int32_t* faces_idx;  /* [F,3]                                                */
  int32_t V, F;
  float eye_z, near_, far_;  /* as jaf_render_fim_wim                                */
  int32_t flags;             /* JAF_POSES_* (0 = clear the keys on entry, leave them as they fall) */
  float* T;                  /* [B,K,S,S,2] out, or NULL (then it never exists)      */
  int32_t* fim;              /* [B,S,S] out, or NULL                                 */
  void* workspace;           /* jaf_raster_workspace_bytes(B, S)                     */
} JafPoseFlowParams;

int jaf_warp_fuse_from_poses_supported(int C, int K, int feat_layout, int feat_dtype);
int jaf_warp_fuse_from_poses(const JafWarpFuseParams* p, const JafPoseFlowParams* q);

/* Same operation with every pointer in `p` a HOST pointer (pinned or pageable): the
 * library stages frames through device buffers it owns, overlapping H2D copies, the
 * kernel and D2H copies on its own streams, `frames_per_chunk` target frames at a time
 * (0 = pick).  Reference sets are uploaded once per distinct ref_index run.  Blocking. */
int jaf_warp_fuse_host(const JafWarpFuseParams* p, int frames_per_chunk);

/* The pose-driven operation (jaf_warp_fuse_from_poses) with HOST buffers: what the application really moves per target
 * frame is a pose (82 KB), its logits and mask up, and the fused RGB frame down; the K references and reference poses
 * of a video go up once (ref_index).  Every pointer of `p` and `q` is a HOST pointer (q->workspace, q->T, q->fim are
 * ignored / must be NULL).  out_feat_device (DEVICE pointer [B,H,W,C], or NULL): when given, the fused features are
 * written there and never cross PCIe (they feed the next device stage — ConvLSTM / generators); p->out_feat (host)
 * is then ignored.  One pipeline per device; blocking. */
int jaf_warp_fuse_from_poses_host(const JafWarpFuseParams* p, const JafPoseFlowParams* q, int frames_per_chunk,
                                  void* out_feat_device);

/* ---------------------------------------------------------------------------------
 * a10  float_estimate.warp_image (src/cal_flow.py:37-39): jaf_warp_fuse with K = 1.
 * src [N,C,Hs,Ws] f32 planar; grid [N,H,W,2]; out [N,C,H,W].
 * --------------------------------------------------------------------------------- */
int jaf_warp_image(const float* src, const float* grid, int N, int C, int Hs, int Ws, int H, int W,
                   int align_corners, float* out, void* stream);

/* ---------------------------------------------------------------------------------
 * a11  Propagation3DFlowNet.forward lines 91 and 98 (src/flow_net.py:87-99)
 * tsf, fake, pred, masked_out: [B,C,H,W]; mask [B,mask_c,H,W] (mask_c 1 or C) or NULL;
 * conf [B,1,H,W].  masked_out (nullable) = tsf*mask; pred (nullable) =
 * fake*conf + (tsf*mask)*(1-conf).
 * --------------------------------------------------------------------------------- */
int jaf_mask_blend(const float* fake, const float* tsf, const float* mask, int mask_c,
                   const float* conf, int B, int C, int H, int W, float* masked_out, float* pred,
                   void* stream);

/* ---------------------------------------------------------------------------------
 * a12  softmax-over-K reduction of Downsampler_mask.forward (src/networks.py:1264-1286)
 * feat [B,K*C,H,W] (channel-concatenated references); logits [B,K,H,W] = the mask conv
 * output BEFORE nn.Softmax(dim=1); out [B,C,H,W].
 * --------------------------------------------------------------------------------- */
int jaf_softmax_fuse(const float* feat, const float* logits, int B, int K, int C, int H, int W,
                     float* out, void* stream);

/* ---------------------------------------------------------------------------------
 * a13  ConvLSTMCell.forward (src/convLSTM.py:41-56), fp32 reference layout
 * x [B,Cin,H,W]; h,c [B,Ch,H,W]; weight [4Ch,Cin+Ch,kh,kw]; bias [4Ch] or NULL;
 * h_out,c_out [B,Ch,H,W].  Gate order in the 4Ch rows is i,f,o,g (:46).
 * CUDA-core direct convolution with the gate epilogue fused — the right tool for the
 * reference's own channel counts (12..96, src/networks.py:1304-1313).
 * --------------------------------------------------------------------------------- */
int jaf_convlstm_step_f32(const float* x, const float* h, const float* c, const float* weight,
                          const float* bias, int B, int Cin, int Ch, int H, int W, int kh, int kw,
                          float* h_out, float* c_out, void* stream);

/* ---------------------------------------------------------------------------------
 * row F  per-reference visibility (optional `vis` input of jaf_warp_fuse)
 * replaces: SMPLRenderer.get_vis_f2pts (src/nmr.py:507-546): faces absent from the SOURCE
 *           pose's face-index map are invisible (the reference gives them the sentinel -2).
 * fim_src [B,K,HW] i32 (face-index maps of the K reference poses), fim_tgt [B,HW] i32.
 * seen [B,K,F] u8 (out): 1 if face f shows in fim_src[b,k].  vis [B,K,HW] f32 (out, optional
 * together with fim_tgt): 1 where fim_tgt[b,p] >= 0 and that face is seen in reference k.
 * --------------------------------------------------------------------------------- */
int jaf_face_visibility(const int32_t* fim_src, const int32_t* fim_tgt, int B, int K, int HW, int F,
                        uint8_t* seen, float* vis, void* stream);

/* ---------------------------------------------------------------------------------
 * a13  the same cell as a tensor-core implicit GEMM (tcgen05 + TMEM + TMA), for wide
 * cells (BASELINE config 4: Cin = Ch = 256, 64x64, B = 16).  3x3 kernel, pad 1.
 * Activations are channels-last bf16: x [B,H,W,Cin], h [B,H,W,Ch]; the cell state stays
 * fp32: c, c_out [B,H,W,Ch]; h_out [B,H,W,Ch] bf16.  `wpack` is the weight repacked once
 * by jaf_convlstm_pack_weight (bf16, tap-major, gate-interleaved per 64 hidden channels);
 * bias [4Ch] f32 or NULL.  Requirements: Cin % 64 == 0, Ch % 64 == 0, W % 64 == 0 or
 * 128 % W == 0 (whole rows per 128-pixel tile).
 * --------------------------------------------------------------------------------- */
size_t jaf_convlstm_wpack_bytes(int Cin, int Ch);
int jaf_convlstm_pack_weight(const float* weight /* [4Ch,Cin+Ch,3,3] f32 */, int Cin, int Ch,
                             void* wpack, void* stream);
int jaf_convlstm_step_tc(const void* x, const void* h, const float* c, const void* wpack,
                         const float* bias, int B, int Cin, int Ch, int H, int W, void* h_out,
                         float* c_out, void* stream);

/* ---------------------------------------------------------------------------------
 * a13  G independent reference-sized cells in ONE launch, on the tensor cores with
 * fp32-grade accuracy (split-bf16: hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM).
 * replaces: the 24 part-specific cells of one pyramid level of Accumulate_LSTM_no_loss
 *           (src/networks.py:1304-1313,:1346-1355,:1641-1662 — 24 x ConvLSTMCell.forward,
 *           src/convLSTM.py:41-56, each its own cuDNN conv + 11 elementwise launches).
 * Reference layout and dtype throughout: x [G,B,Cin,H,W], h, c, h_out, c_out
 * [G,B,Ch,H,W] f32; weight [G,4Ch,Cin+Ch,3,3] f32 repacked once by
 * jaf_convlstm_gpack_weight into `wpack` (jaf_convlstm_gpack_bytes bytes); bias [G,4Ch]
 * f32 or NULL.  3x3 kernel, pad 1.  G = 1 is the plain cell.  Requirements: Ch % 4 == 0
 * (Ch % 8 == 0 above 64), Ch <= 128.
 * --------------------------------------------------------------------------------- */
size_t jaf_convlstm_gpack_bytes(int G, int Cin, int Ch);
/* 1 when jaf_convlstm_step_grouped can run this cell on the current device (channel counts supported AND the row
 * window + weight ring fit the SM's shared memory), else 0: callers fall back to jaf_convlstm_step_f32, which takes
 * every cell the reference constructor accepts. */
int jaf_convlstm_grouped_supported(int G, int B, int Cin, int Ch, int H, int W);
int jaf_convlstm_gpack_weight(const float* weight, int G, int Cin, int Ch, void* wpack,
                              void* stream);
int jaf_convlstm_step_grouped(const float* x, const float* h, const float* c, const void* wpack,
                              const float* bias, int G, int B, int Cin, int Ch, int H, int W,
                              float* h_out, float* c_out, void* stream);

/* The recurrence over the K references in ONE call (src/convLSTM.py:131-134 `for t in range(seq_len)`; driven per
 * pyramid level by src/networks.py:1346-1355): T grouped steps launched back to back, step t reading x_seq[:, :, t] and
 * h_seq[:, :, t-1] in place and writing h_seq[:, :, t] — no per-step slicing, stacking or host round trip.
 * x_seq [G,B,T,Cin,H,W]; h0, c0 [G,B,Ch,H,W] (zeros for the reference's default state); h_seq out [G,B,T,Ch,H,W]
 * (= the batch_first layer output of ConvLSTM.forward, one per cell); c_last out [G,B,Ch,H,W]; c_tmp: workspace of the
 * same size (unused for T == 1).  Results are bit-identical to T calls of jaf_convlstm_step_grouped. */
int jaf_convlstm_sequence_grouped(const float* x_seq, const float* h0, const float* c0, const void* wpack,
                                  const float* bias, int G, int B, int T, int Cin, int Ch, int H, int W,
                                  float* h_seq, float* c_last, float* c_tmp, void* stream);

/* ---------------------------------------------------------------------------------
 * SURVEY §8f rank 3  bidirectional multi-scale feature warp of SpatioTempoCRN
 * replaces, per pyramid level (src/crn_model.py:457-566):
 *   flow_s = F.interpolate(flow, size, mode='nearest')
 *   out_fwd = F.grid_sample(feat_fwd, (grid + flow_s).permute(0,2,3,1), padding_mode='border')
 *   out_bwd = F.grid_sample(feat_bwd, (grid - flow_s).permute(0,2,3,1), padding_mode='border')
 * feat_fwd / feat_bwd / out_* [B,C,h,w] f32 (either pair may be NULL); base_grid [B,2,h,w]
 * (channel 0 = x, 1 = y, the `grid_list[i]` input); flow [B,2,H,W] at full resolution.
 * --------------------------------------------------------------------------------- */
int jaf_flow_warp_pair(const float* feat_fwd, const float* feat_bwd, const float* base_grid,
                       const float* flow, int B, int C, int h, int w, int H, int W, int align_corners,
                       float* out_fwd, float* out_bwd, void* stream);

/* ---------------------------------------------------------------------------------
 * SURVEY §8f rank 1  IUV texture lookup
 * replaces: texture_warp_pytorch (test/conv_pro_test.py:41-74; train/4.convLSTM_flowpro_interval.py:43-76):
 *           24 x (torch.where x2, grid build, F.grid_sample of one part texture with zero padding,
 *           torch.where) per target frame -> one pass.
 * tex_parts [P,3,Ht,Wt] f32 (P = 24 body-part textures); iuv [B,H,W,3] uint8 (part index, U, V);
 * out [B,3,H,W]: bilinear sample of part iuv[...,0]-1 at x = ((255-V)/255 - .5)*2, y = (U/255 - .5)*2,
 * 0 where the part index is 0 or > P.
 * --------------------------------------------------------------------------------- */
int jaf_texture_warp(const float* tex_parts, int P, int Ht, int Wt, const uint8_t* iuv, int B, int H, int W,
                     int align_corners, float* out, void* stream);

/* ---------------------------------------------------------------------------------
 * SURVEY §8f rank 2  texture-space assembly around the ConvLSTM accumulation
 * replaces: the part / reference slicing loops of test/conv_pro_test.py:209-217
 *           (train/4.convLSTM_flowpro_interval.py:269-277), the common-area OR + masking of
 *           :221-236, and the atlas re-assembly of src/networks.py:1685-1691.
 * The atlas is rows x cols parts of ph x pw pixels (reference: 4 x 6 x 200 x 200 = 800 x 1200).
 *   gather : atlas [B,Kmax,C,rows*ph,cols*pw], ref_index [K] i32 (the `random_index` frames)
 *            -> out [rows*cols, K, B, C, ph, pw]  (per part = torch.cat(x_in[part], dim=0))
 *   common_mask : parts [rows*cols, B, C, ph, pw] *= float(OR_z uint8(mask[b, ref[z]])) in place;
 *            mask [B,Kmax,rows*ph,cols*pw] f32
 *   scatter: parts [rows*cols, B, C, ph, pw] -> atlas [B, C, rows*ph, cols*pw]
 * --------------------------------------------------------------------------------- */
int jaf_texture_parts_gather(const float* atlas, const int32_t* ref_index, int B, int Kmax, int K, int C,
                             int rows, int cols, int ph, int pw, float* out, void* stream);
int jaf_texture_parts_common_mask(float* parts, const float* mask, const int32_t* ref_index, int B,
                                  int Kmax, int K, int C, int rows, int cols, int ph, int pw,
                                  void* stream);
int jaf_texture_parts_scatter(const float* parts, int B, int C, int rows, int cols, int ph, int pw,
                              float* atlas, void* stream);

/* ---------------------------------------------------------------------------------
 * SURVEY §8f rank 4  per-frame IUV preprocessing of the data loader
 * jaf_transfer_texture replaces TransferTexture (src/utils.py:369-394; three calls per frame at
 *   src/data.py:102-113): nearest texel of the atlas at U = rint(IUV[1]/255.*(ps-1)), V likewise,
 *   row i*ps + U, column j*ps + (ps-1-V) of part 6*i+j+1; output channels that come out 0 take `im`
 *   when it is given.  tex [rows*ps, cols*ps, 3] u8 (or one per frame when tex_batched); iuv, im, out
 *   [B,H,W,3] u8.  The reference hard-codes rows=4, cols=6, ps=200, H=W=256.
 * jaf_iuv_part_stats feeds compute_angle (src/computer_angle.py:4-39; src/data.py:504):
 *   counts [B,32] i32 = pixels of each part id < 32, sumx [B,32] i64 = sum of their column index.
 * --------------------------------------------------------------------------------- */
int jaf_transfer_texture(const uint8_t* tex, int tex_batched, int rows, int cols, int part_size,
                         const uint8_t* iuv, const uint8_t* im, int B, int H, int W, uint8_t* out,
                         void* stream);
int jaf_iuv_part_stats(const uint8_t* iuv, int B, int H, int W, int32_t* counts, int64_t* sumx,
                       void* stream);

/* ---------------------------------------------------------------------------------
 * SURVEY §8f rank 4  DensePose texture extraction
 * replaces: get_texture(im, IUV, tex_size=32, final_size=200) (src/utils.py:232-255): per part 1..24 scatter the part's
 *   pixels into a tex_size^2 map (row = int((255-V)*(tex_size-1)/255.), col = int(U*(tex_size-1)/255.), last pixel in
 *   row-major order wins), cv2.resize(..., INTER_LINEAR) to final_size^2 on float64 data, [:, :, ::-1] / 255.
 * im, iuv [B,H,W,3] u8 (im in cv2.imread's BGR order); parts out [B,24,final_size,final_size,3] f64 (zeros for a part
 * without pixels); workspace: jaf_get_texture_workspace_bytes(B, tex_size).  The resize is the half-pixel-centre
 * bilinear of cv::resize evaluated in fp64 with exactly rounded coefficients; OpenCV builds differ in how they round
 * theirs (generic path: float; IPP path: double), so agreement with cv2 is <= 1e-12, not bitwise.
 * --------------------------------------------------------------------------------- */
size_t jaf_get_texture_workspace_bytes(int B, int tex_size);
int jaf_get_texture(const uint8_t* im, const uint8_t* iuv, int B, int H, int W, int tex_size, int final_size,
                    double* parts, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JAFPRO_B200_H_ */
