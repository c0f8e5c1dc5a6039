"""Drop-in for the pybind11 extension ``neural_renderer.cuda.rasterize`` on the hot path
(NR/cuda/rasterize_cuda.cpp:70-95,194-200): the one export the flow path reaches,
``forward_face_index_map``, with the extension's exact signature and in-place contract.

    import jafpro_b200.cuda_rasterize as rasterize_cuda      # instead of neural_renderer.cuda.rasterize
    fim, wim, depth, finv = rasterize_cuda.forward_face_index_map(
        faces, face_index_map, weight_map, depth_map, face_inv_map, faces_inv,
        image_size, near, far, return_rgb, return_alpha, return_depth)          # NR/rasterize.py:161-169

The caller pre-fills the outputs (NR/rasterize.py:50-52,164), they are filled in place and the same tensor
objects are returned (rasterize_cuda_kernel.cu:650).  CPU or non-contiguous tensors raise ``RuntimeError``
like ``CHECK_INPUT`` does (rasterize_cuda.cpp:66-68,84-89).  Forward only: the backward exports of the extension
(backward_pixel_map, backward_textures, backward_depth_map) and the texture sampling are not on the path.
"""
from __future__ import annotations

import torch

from . import _lib, ops


def _check_input(t, name, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}")
    ops.forbid_grad(t, name=name)
    return t


def forward_face_index_map(faces, face_index_map, weight_map, depth_map, face_inv_map, faces_inv, image_size, near, far,
                           return_rgb, return_alpha, return_depth):
    """-> [face_index_map, weight_map, depth_map, face_inv_map] (the tensors passed in, filled in place)."""
    faces = _check_input(faces, "faces", torch.float32)
    fim = _check_input(face_index_map, "face_index_map", torch.int32)
    wim = _check_input(weight_map, "weight_map", torch.float32)
    depth = _check_input(depth_map, "depth_map", torch.float32)
    finv_map = _check_input(face_inv_map, "face_inv_map", torch.float32)
    finv = _check_input(faces_inv, "faces_inv", torch.float32)
    if faces.dim() != 4 or tuple(faces.shape[2:]) != (3, 3):
        raise RuntimeError("faces must be [batch, num_faces, 3, 3]")
    B, F = faces.shape[:2]
    S = int(image_size)
    if fim.numel() != B * S * S or wim.numel() != B * S * S * 3 or depth.numel() != B * S * S:
        raise RuntimeError("face_index_map / weight_map / depth_map must be [B,S,S], [B,S,S,3], [B,S,S]")
    if finv.numel() != faces.numel():
        raise RuntimeError("faces_inv must have the shape of faces")
    if return_depth and finv_map.numel() != B * S * S * 9:
        raise RuntimeError("return_depth needs face_inv_map [B,S,S,3,3]")
    dev = faces.device
    with ops._on(dev):
        ws = ops._workspace(dev, _lib.lib().jaf_raster_workspace_bytes(B, S))
        _lib.check(_lib.lib().jaf_forward_face_index_map(
            faces.data_ptr(), fim.data_ptr(), wim.data_ptr(), depth.data_ptr(),
            finv_map.data_ptr() if return_depth else None, finv.data_ptr(), B, F, S, float(near), float(far),
            int(bool(return_depth)), ws.data_ptr(), ops._stream()), "forward_face_index_map")
    return [face_index_map, weight_map, depth_map, face_inv_map]
