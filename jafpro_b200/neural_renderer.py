"""Drop-in for the functions of ``neural_renderer`` that the hot path reaches
(third_party/neural_renderer/neural_renderer, "NR").  Same names, argument meaning and
return shapes; CUDA tensors only."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

DEFAULT_IMAGE_SIZE = 256      # NR/rasterize.py:8-13
DEFAULT_ANTI_ALIASING = True
DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100
DEFAULT_EPS = 1e-4


def _as_vec(v, device):
    if isinstance(v, (list, tuple)):
        return torch.tensor(v, dtype=torch.float32, device=device)
    if isinstance(v, np.ndarray):
        return torch.from_numpy(v).to(device)
    return v.to(device)


def look_at(vertices, eye, at=(0, 0, 0), up=(0, 1, 0)):
    """NR/look_at.py:6-62.  (The reference calls torch.cross without dim=, which silently picks
    dim 0 when the batch size is exactly 3; this version always crosses along the last dim.)"""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    device = vertices.device
    at, up, eye = _as_vec(at, device), _as_vec(up, device), _as_vec(eye, device)
    bs = vertices.shape[0]
    if eye.ndimension() == 1:
        eye = eye[None, :].repeat(bs, 1)
    if at.ndimension() == 1:
        at = at[None, :].repeat(bs, 1)
    if up.ndimension() == 1:
        up = up[None, :].repeat(bs, 1)
    z_axis = F.normalize(at - eye, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    r = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    if vertices.shape != eye.shape:
        eye = eye[:, None, :]
    vertices = vertices - eye
    # exact fp32 products: with allow_tf32 the identity rotation would round the coordinates
    return (vertices[:, :, None, :] * r[:, None, :, :]).sum(-1)


def vertices_to_faces(vertices, faces):
    """NR/vertices_to_faces.py:4-22: [B,V,3], [B,F,3] -> [B,F,3,3]."""
    assert vertices.ndimension() == 3 and faces.ndimension() == 3
    assert vertices.shape[0] == faces.shape[0] and vertices.shape[2] == 3 and faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    faces = faces.long() + (torch.arange(bs, device=vertices.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[faces]


def rasterize_face_index_map_and_weight_map(faces, image_size=DEFAULT_IMAGE_SIZE,
                                            anti_aliasing=DEFAULT_ANTI_ALIASING, near=DEFAULT_NEAR,
                                            far=DEFAULT_FAR, eps=DEFAULT_EPS):
    """NR/rasterize.py:543-571.  Returns (face_index_map int32 [B,S,S], weight_map f32 [B,S,S,3]),
    rows already flipped (:334-338).  As in the reference, anti_aliasing=True rasterises at 2x and
    the index / weight maps are returned at that doubled size (:315-316, only rgb/alpha/depth are
    pooled back, :340-347)."""
    size = image_size * 2 if anti_aliasing else image_size
    return ops.raster_fim_wim(faces.contiguous(), size, near, far, flip_rows=True)


def rasterize_face_index_map(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                             near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    return rasterize_face_index_map_and_weight_map(faces, image_size, anti_aliasing, near, far, eps)[0]


def rasterize_weight_map(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                         near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    return rasterize_face_index_map_and_weight_map(faces, image_size, anti_aliasing, near, far, eps)[1]


def rasterize_silhouettes(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                          near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    """Coverage only (alpha = fim >= 0, NR/rasterize.py:118-127); used by the golden teapot test."""
    size = image_size * 2 if anti_aliasing else image_size
    fim, _ = ops.raster_fim_wim(faces.contiguous(), size, near, far, flip_rows=True)
    alpha = (fim >= 0).float()
    if anti_aliasing:
        alpha = F.avg_pool2d(alpha[:, None], kernel_size=(2, 2))[:, 0]
    return alpha
