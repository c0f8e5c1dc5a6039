#!/usr/bin/env python
"""One synthetic video through the drop-ins, in the order test/conv_pro_test.py uses the reference functions.

    python examples/video_pipeline.py            # needs a B200 and the built library

Nothing here is a model: the reference's CNN stages (encoders, inpainting / refinement / background nets) are outside the
scope of this repository and are replaced by the cheapest stand-ins that keep the tensor shapes (a fixed 1x1 lift from
3 to 12 channels, identity elsewhere)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jafpro_b200 import synth  # noqa: E402
from jafpro_b200.computer_angle import compute_angles  # noqa: E402
from jafpro_b200.convLSTM import ConvLSTM, ConvLSTMGrouped  # noqa: E402
from jafpro_b200.flow_net import Propagation3DFlowNet  # noqa: E402
from jafpro_b200.fusion import warp_fuse_from_poses  # noqa: E402
from jafpro_b200.nmr import SMPLRenderer  # noqa: E402
from jafpro_b200.texture import assemble_atlas, gather_parts, mask_common_area_, texture_warp_pytorch  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
torch.set_grad_enabled(False)  # inference, like test/conv_pro_test.py:190 — the drop-ins are forward-only and say so
FRAMES, K, S, C = 30, 4, 256, 64

# ---- per video: pick K reference frames by view angle (src/data.py:504: compute_angle on every IUV map)
iuv = torch.zeros(FRAMES, S, S, 3, dtype=torch.uint8, device=dev)
iuv[:, 60:200, 90:170, 0] = torch.randint(1, 25, (FRAMES, 1, 1), device=dev, dtype=torch.uint8)
iuv[..., 1:] = torch.randint(0, 256, (FRAMES, S, S, 2), device=dev, dtype=torch.uint8)
angles = compute_angles(iuv)                                            # one launch for the whole video
ref_frames = np.argsort(np.abs(np.asarray(angles, dtype=np.float64)))[:K]
print("reference frames by |view angle|:", ref_frames.tolist())

# ---- texture space: 24 parts x K references -> per-part ConvLSTM accumulation -> common-area mask -> atlas
src_texture_im = torch.rand(1, 5, 3, 800, 1200, device=dev) * 2 - 1    # [B, Kmax, 3, 800, 1200]
src_mask_im = (torch.rand(1, 5, 800, 1200, device=dev) > 0.3).float()
parts = gather_parts(src_texture_im, list(range(K)))                    # [24, K, 1, 3, 200, 200]
lift = torch.randn(12, 3, device=dev) * 0.5                              # stand-in for Downsampler.enc1 (3 -> 12 channels)
x1_con = torch.einsum("oc,pkbchw->pbkohw", lift, parts).contiguous()     # [24, B, K, 12, 200, 200]
lstms = [ConvLSTM((200, 200), 12, [12], [(3, 3)], 1, batch_first=True, bias=True).to(dev) for _ in range(24)]
accu = ConvLSTMGrouped.from_lstms(lstms)                                 # 24 part cells, one launch per step
_, (h_last, _) = accu(x1_con)                                            # [24, B, 12, 200, 200]
accu_out = torch.einsum("co,pbohw->pbchw", torch.linalg.pinv(lift), h_last).contiguous()  # stand-in for the upsampler
mask_common_area_(accu_out, src_mask_im, list(range(K)))
atlas = assemble_atlas(accu_out)                                         # [1, 3, 800, 1200]
print("texture atlas:", tuple(atlas.shape), "non-zero fraction %.2f" % float((atlas != 0).float().mean()))

# ---- per frame: IUV texture lookup (conv_pro_test.py:262), then the warp-and-fuse hot path from poses
tex_parts = accu_out[:, 0]                                              # 24 x [3, 200, 200]
inpaint_warp = texture_warp_pytorch(tex_parts, iuv)                      # [FRAMES, 3, 256, 256], one launch
cam, verts = synth.smpl_poses(FRAMES + K, seed=3, device=dev)
renderer = SMPLRenderer(image_size=S).to(dev)
rgb, feat = synth.reference_sets(1, K, C, S, S, seed=1, device=dev, channels_last=True)   # K references of the video
ref_index = torch.zeros(FRAMES, dtype=torch.int32, device=dev)          # every frame uses the video's reference set
logits = torch.randn(FRAMES, K, S, S, device=dev)
src_cams = cam[FRAMES:].unsqueeze(0).expand(FRAMES, -1, -1).contiguous()
src_verts = verts[FRAMES:].unsqueeze(0).expand(FRAMES, -1, -1, -1).contiguous()
fused_rgb, fused_feat, T, fim = warp_fuse_from_poses(renderer, src_cams, src_verts, cam[:FRAMES].contiguous(),
                                                     verts[:FRAMES].contiguous(), rgb=rgb, feat=feat, logits=logits,
                                                     ref_index=ref_index, per_reference_visibility=True)
print("fused rgb", tuple(fused_rgb.shape), "fused features", tuple(fused_feat.shape), fused_feat.dtype,
      "foreground %.1f %%" % (100 * float((fim != -1).float().mean())))

# ---- confidence blend of the warped appearance with a generated frame (src/flow_net.py:87-99)
class _StandInUnet(torch.nn.Module):  # the reference's CompositeWeightUnet is a generic conv stack, out of scope here
    def forward(self, cated):
        return torch.sigmoid(cated[:, :1] - cated[:, 3:4])


net = Propagation3DFlowNet(_StandInUnet()).to(dev)
tgt_mask = (fim != -1).float().unsqueeze(1).repeat(1, 3, 1, 1)
out = net({"fake_tgt": inpaint_warp.contiguous(), "tsf_image": fused_rgb, "tgt_IUV": None, "use_IUV": False,
           "use_mask": True, "tgt_smpl_mask": tgt_mask})
pred = out["pred_target"]
print("blended prediction", tuple(pred.shape), "finite:", bool(torch.isfinite(pred).all()))
torch.cuda.synchronize()
print("ok")
