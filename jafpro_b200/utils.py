"""Drop-in for the IUV texture transfer of ``src/utils.py`` used by the data loader (src/data.py:102-113)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def TransferTexture(TextureIm, IUV, im=None):
    """Same call as the reference (src/utils.py:369-394): TextureIm (800,1200,3) 0..255, IUV (256,256,3), optional
    background ``im``; numpy in -> numpy out.  Tensors on the GPU may be passed instead, with a leading batch dimension
    on IUV / im (and optionally on TextureIm): then the whole batch is one launch and the result stays on the GPU."""
    as_numpy = isinstance(IUV, np.ndarray)

    def dev(a):
        if a is None:
            return None
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(device="cuda", dtype=torch.uint8).contiguous()

    tex, iuv, bg = dev(TextureIm), dev(IUV), dev(im)
    single = iuv.dim() == 3
    if single:
        iuv = iuv[None]
        bg = None if bg is None else bg[None]
    out = ops.transfer_texture(tex, iuv, bg)
    out = out[0] if single else out
    return out.cpu().numpy() if as_numpy else out


def get_texture(im, IUV, tex_size=32, final_size=200):
    """Same call as the reference (src/utils.py:232-255): im (H,W,3) uint8 as cv2.imread returns it, IUV (H,W,3) ->
    list of 24 arrays (final_size, final_size, 3) float64 in [0, 1], channels reversed.  GPU tensors with a leading
    batch dimension may be passed instead: the result is then one tensor [B,24,final_size,final_size,3] on the GPU."""
    as_numpy = isinstance(IUV, np.ndarray)

    def dev(a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(device="cuda", dtype=torch.uint8).contiguous()

    img, iuv = dev(im), dev(IUV)
    single = iuv.dim() == 3
    if single:
        img, iuv = img[None], iuv[None]
    parts = ops.get_texture(img, iuv, tex_size, final_size)
    if as_numpy:
        parts = parts.cpu().numpy()
        return [parts[0, p] for p in range(24)] if single else parts
    return parts[0] if single else parts
