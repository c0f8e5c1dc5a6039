"""Multi-reference appearance fusion: the composite warp+fuse operator and the
standalone softmax-over-K reduction of ``Downsampler_mask.forward`` (src/networks.py:1259-1286)."""
from __future__ import annotations

import torch

from . import ops
from .ops import warp_fuse, warp_fuse_from_poses_host, warp_fuse_host  # noqa: F401  (re-exported: the public entry points)


def _eye_z(renderer):
    from .nmr import SMPLRenderer
    return SMPLRenderer._eye_z_f32(renderer)  # also serves the reference's own SMPLRenderer (it only has `.eye`)


def softmax_fuse(x_con, mask_logits):
    """x_con [B,K*C,h,w] (references concatenated on channels, :1259-1263), mask_logits [B,K,h,w]
    (the mask conv output before nn.Softmax(dim=1), :1230-1244) -> sum_k softmax_k * x_k [B,C,h,w]."""
    return ops.softmax_fuse(x_con.contiguous(), mask_logits.contiguous())


def to_channels_last_5d(feat):
    """[R,K,C,H,W] -> same logical tensor with channels-last strides (dense [R,K,H,W,C] memory)."""
    return feat.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


def reference_visibility(renderer, src_cams, src_vertices, fim_tgt):
    """Per-reference visibility maps for `warp_fuse(vis=...)`: rasterise the K reference poses and apply the
    reference's rule (get_vis_f2pts, src/nmr.py:507-546).  src_cams [B,K,3], src_vertices [B,K,V,3], fim_tgt [B,S,S]
    -> vis [B,K,S,S] f32."""
    B, K = src_cams.shape[:2]
    S = renderer.image_size
    _, fim_src, _ = ops.render_fim_wim(src_cams.reshape(B * K, 3).contiguous(),
                                       src_vertices.reshape(B * K, -1, 3).contiguous(), renderer.faces, S,
                                       eye_z=_eye_z(renderer), return_faces=False)
    _, vis = ops.face_visibility(fim_src.reshape(B, K, S, S), fim_tgt.contiguous(), renderer.faces.shape[-2])
    return vis


def warp_fuse_from_poses(renderer, src_cams, src_vertices, tgt_cam, tgt_vertices, rgb=None, feat=None, *,
                         logits=None, tgt_mask=None, ref_index=None, align_corners: bool = False,
                         per_reference_visibility: bool = False):
    """The whole hot path in two steps on the device: transfer flows of the K reference poses into the
    target pose (one raster of the target, K composes) and the fused warp + fusion with the default
    visibility `target pixel is on the body` (fim != -1), or with `per_reference_visibility` the reference's
    get_vis_f2pts rule (the face under the target pixel must be visible in reference k).  `renderer` is a jafpro_b200.nmr.SMPLRenderer.
    src_cams [B,K,3], src_vertices [B,K,V,3], tgt_cam [B,3], tgt_vertices [B,V,3]; references as in warp_fuse.
    Returns (out_rgb, out_feat, T, fim)."""
    if not per_reference_visibility and feat is not None:
        # one pass: the flows are composed inside the warp kernel (T and fim are emitted because this API returns them)
        return ops.warp_fuse_from_poses(src_cams.contiguous(), src_vertices.contiguous(), tgt_cam.contiguous(),
                                        tgt_vertices.contiguous(), renderer.faces, renderer.image_size, rgb=rgb, feat=feat,
                                        logits=logits, tgt_mask=tgt_mask, ref_index=ref_index, align_corners=align_corners,
                                        eye_z=_eye_z(renderer), return_flow=True)
    T, fim, _ = ops.cal_flow_multi(src_cams.contiguous(), src_vertices.contiguous(), tgt_cam.contiguous(),
                                   tgt_vertices.contiguous(), renderer.faces, renderer.image_size,
                                   eye_z=_eye_z(renderer), return_wim=False)
    vis = reference_visibility(renderer, src_cams, src_vertices, fim) if per_reference_visibility else None
    out_rgb, out_feat = ops.warp_fuse(T, rgb=rgb, feat=feat, logits=logits, vis=vis, fim=None if vis is not None else fim,
                                      tgt_mask=tgt_mask, ref_index=ref_index, align_corners=align_corners)
    return out_rgb, out_feat, T, fim
