"""Multi-reference appearance fusion: the composite warp+fuse operator and the
standalone softmax-over-K reduction of ``Downsampler_mask.forward`` (src/networks.py:1259-1286)."""
from __future__ import annotations

import torch

from . import ops
from .ops import warp_fuse, warp_fuse_host  # noqa: F401  (re-exported: the public entry points)


def softmax_fuse(x_con, mask_logits):
    """x_con [B,K*C,h,w] (references concatenated on channels, :1259-1263), mask_logits [B,K,h,w]
    (the mask conv output before nn.Softmax(dim=1), :1230-1244) -> sum_k softmax_k * x_k [B,C,h,w]."""
    return ops.softmax_fuse(x_con.contiguous(), mask_logits.contiguous())


def to_channels_last_5d(feat):
    """[R,K,C,H,W] -> same logical tensor with channels-last strides (dense [R,K,H,W,C] memory)."""
    return feat.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
