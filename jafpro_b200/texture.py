"""Drop-in for ``texture_warp_pytorch`` (test/conv_pro_test.py:41-74, train/4.convLSTM_flowpro_interval.py:43-76):
the per-frame IUV texture lookup of the reference's inference / training loops."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def texture_warp_pytorch(tex_parts, IUV, device=None, align_corners: bool = False):
    """Same call as the reference: ``tex_parts`` is a list (or tensor) of 24 part textures [3,Ht,Wt], ``IUV`` a numpy
    array or tensor [H,W,3] (part index, U, V; 0..255).  Returns the generated image [3,H,W].
    A batched IUV [B,H,W,3] returns [B,3,H,W].  One kernel launch instead of ~120."""
    if isinstance(tex_parts, (list, tuple)):
        tex_parts = torch.stack([t.float() for t in tex_parts], 0)
    dev = tex_parts.device if tex_parts.is_cuda else torch.device("cuda" if device is None else device)
    tex = tex_parts.to(dev).float().contiguous()
    if isinstance(IUV, np.ndarray):
        IUV = torch.from_numpy(IUV)
    iuv = IUV.to(dev)
    if iuv.dtype != torch.uint8:
        iuv = iuv.to(torch.uint8)
    single = iuv.dim() == 3
    out = ops.texture_warp(tex, (iuv[None] if single else iuv).contiguous(), align_corners)
    return out[0] if single else out
