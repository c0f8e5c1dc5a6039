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


# ---- texture-space assembly around Accumulate_LSTM_no_loss -------------------------------
def gather_parts(src_texture_im, random_index, rows: int = 4, cols: int = 6):
    """test/conv_pro_test.py:209-217 in one launch: ``src_texture_im`` [B,Kmax,3,800,1200], ``random_index`` the selected
    reference frames -> tensor [24, K, B, 3, 200, 200]; ``out[p].flatten(0, 1)`` is what
    ``Downsampler_convLSTM.forward`` builds with ``torch.cat(x_in[p], dim=0)`` (src/networks.py:1316) and
    ``list(out[p])`` is the reference's ``src_texture_im_input[p]``."""
    idx = torch.as_tensor(np.asarray(random_index), dtype=torch.int32, device=src_texture_im.device)
    return ops.texture_parts_gather(src_texture_im.float().contiguous(), idx, rows, cols)


def mask_common_area_(parts, src_mask_im, random_index, rows: int = 4, cols: int = 6):
    """test/conv_pro_test.py:221-236 in one launch, in place: ``parts`` [24,B,3,200,200] (the stacked
    ``Accu_output_texture``) times the OR over the selected references of ``src_mask_im`` [B,Kmax,800,1200]."""
    idx = torch.as_tensor(np.asarray(random_index), dtype=torch.int32, device=parts.device)
    return ops.texture_parts_common_mask_(parts, src_mask_im.float().contiguous(), idx, rows, cols)


def assemble_atlas(parts, rows: int = 4, cols: int = 6):
    """src/networks.py:1685-1691: 24 part images [24,B,3,200,200] (or a list of [B,3,200,200]) -> [B,3,800,1200]."""
    if isinstance(parts, (list, tuple)):
        parts = torch.stack(list(parts), 0)
    return ops.texture_parts_scatter(parts.float().contiguous(), rows, cols)
