"""Drop-in for ``src/cal_flow.py``: ``float_estimate`` (transfer flow + warp of one reference)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .nmr import SMPLRenderer


class float_estimate(nn.Module):
    """src/cal_flow.py:13-39.  The reference constructor also loads an HMR network and SMPL weights
    that ``forward`` never runs (:17-19); they are not part of this module.

    align_corners: the flows are pixel-centre NDC (align_corners=False, today's torch default for the
    reference's call text); torch 1.2.0 — the version the reference pins — sampled with True semantics.
    """

    def __init__(self, smpl_pkl=None, hmr_model_path=None, image_size=256, align_corners=False, fused=True):
        super().__init__()
        self.render = SMPLRenderer(image_size=image_size, tex_size=3, has_front=True, fill_back=False)
        self.align_corners = align_corners
        self.fused = fused

    def forward(self, src_img, src_smpl, tgt_smpl):
        src_cam, src_pose, src_vertices, src_shape = src_smpl
        tgt_cam, tgt_pose, tgt_vertices, tgt_shape = tgt_smpl
        flow = self.cal_flow(src_cam, src_pose, src_vertices, src_shape, tgt_cam, tgt_pose, tgt_vertices, tgt_shape)
        return self.warp_image(src_img, flow)

    def cal_flow(self, src_cam, src_pose, src_vertices, src_shape, tgt_cam, tgt_pose, tgt_vertices, tgt_shape):
        if self.fused:
            return self.render.cal_flow(src_cam, src_vertices, tgt_cam, tgt_vertices)
        # the reference's own sequence of calls (:29-34), each served by one kernel group
        src_f2verts, _, _ = self.render.render_fim_wim(src_cam, src_vertices)
        src_f2verts = src_f2verts[:, :, :, 0:2].contiguous()
        src_f2verts[:, :, :, 1] *= -1
        _, tsf_fim, tsf_wim = self.render.render_fim_wim(tgt_cam, tgt_vertices)
        return self.render.cal_bc_transform(src_f2verts, tsf_fim, tsf_wim)

    def warp_image(self, src_image, flow):
        return ops.grid_sample_border(src_image.contiguous(), flow.contiguous(), align_corners=self.align_corners)
