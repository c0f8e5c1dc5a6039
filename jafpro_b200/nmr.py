"""Drop-in for the parts of ``src/nmr.py`` on the hot path: ``orthographic_proj_withz_idrot``
(:10-28) and ``SMPLRenderer.render_fim_wim`` (:263-278), ``render_fim`` (:246-261),
``cal_bc_transform`` (:617-659), plus the fused ``cal_flow`` the reference spreads over
src/cal_flow.py:28-35."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import neural_renderer as nr
from . import ops

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_smpl_template():
    """(verts [6890,3] f32, faces [13776,3] int32): the T-pose vertices of the reference's mapper.txt and
    the topology of smpl_faces.npy (tools/make_golden.py wrote the .npz)."""
    d = np.load(os.path.join(_DATA, "smpl_template.npz"))
    return d["verts"].astype(np.float32), d["faces"].astype(np.int32)


def orthographic_proj_withz_idrot(X, cam, offset_z=0.):
    """src/nmr.py:10-28: sc * (x + [tx; ty]) keeping z."""
    scale = cam[:, 0].contiguous().view(-1, 1, 1)
    trans = cam[:, 1:3].contiguous().view(cam.size(0), 1, -1)
    proj_xy = scale * (X[:, :, :2] + trans)
    proj_z = X[:, :, 2, None] + offset_z
    return torch.cat((proj_xy, proj_z), 2)


class SMPLRenderer(nn.Module):
    """Geometry half of the reference SMPLRenderer (src/nmr.py:104-177).  Texture, lighting and
    UV tables are outside the warp-and-fuse path and are not built."""

    def __init__(self, face_path=None, uv_map_path=None, map_name='uv_seg', tex_size=3, image_size=256,
                 anti_aliasing=True, fill_back=False, background_color=(0, 0, 0), viewing_angle=30,
                 near=0.1, far=25.0, has_front=False):
        super().__init__()
        faces = np.load(face_path) if face_path is not None else load_smpl_template()[1]
        if fill_back:
            faces = np.concatenate((faces, faces[:, ::-1]), axis=0)  # :129-130
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        self.fill_back = fill_back
        self.tex_size = tex_size
        self.base_nf = faces.shape[0] // (2 if fill_back else 1)
        self.nf = faces.shape[0]
        self.register_buffer('faces', torch.tensor(faces.astype(np.int32)).int())
        self.near, self.far = near, far  # stored but, as in the reference, not used by render_fim_wim
        self.proj_func = orthographic_proj_withz_idrot
        self.viewing_angle = viewing_angle
        self.eye = [0, 0, -(1. / np.tan(np.radians(self.viewing_angle)) + 1)]  # :177
        self._eye_z = float(np.float32(self.eye[2]))

    def _eye_z_f32(self):
        """float32 z of the look_at eye.  Derived from ``self.eye`` when the attribute cache is missing, so that these
        methods also work when they are bound onto the REFERENCE's SMPLRenderer, which only has
        ``eye`` (src/nmr.py:177)."""
        ez = getattr(self, "_eye_z", None)
        if ez is None:
            ez = float(np.float32(self.eye[2]))
            try:
                self._eye_z = ez
            except Exception:  # pragma: no cover - exotic __setattr__
                pass
        return ez

    # ---- src/nmr.py:263-278
    def render_fim_wim(self, cam, vertices, faces=None):
        if faces is None:
            f3, fim, wim = ops.render_fim_wim(cam.contiguous(), vertices.contiguous(), self.faces.int(), self.image_size,
                                              eye_z=SMPLRenderer._eye_z_f32(self))
            return f3, fim, wim
        proj_verts = self.proj_func(vertices, cam)
        proj_verts[:, :, 1] *= -1
        verts = nr.look_at(proj_verts, self.eye)
        f3 = nr.vertices_to_faces(verts, faces)
        fim, wim = nr.rasterize_face_index_map_and_weight_map(f3, self.image_size, False)
        return f3, fim, wim

    # ---- src/nmr.py:246-261
    def render_fim(self, cam, vertices, faces=None):
        return self.render_fim_wim(cam, vertices, faces)[1]

    # ---- src/nmr.py:617-659
    def cal_bc_transform(self, src_f2pts, dst_fims, dst_wims):
        return ops.flow_compose(src_f2pts.contiguous(), dst_fims.contiguous(), dst_wims.contiguous())

    # ---- src/nmr.py:507-546
    @staticmethod
    def get_vis_f2pts(f2pts, fims):
        """Faces absent from `fims` get coordinates -2.  f2pts [bs,f,3,c] with fims [bs,H,W], or [f,3,c] with [H,W].
        Like the reference, which drops the FIRST unique value of the fim assuming it is the background -1
        (`fim.unique()[1:]`), an item without any background pixel also loses its lowest visible face."""
        single = f2pts.dim() == 3
        f2 = f2pts.unsqueeze(0) if single else f2pts
        fm = (fims.unsqueeze(0) if single else fims).to(torch.int32).contiguous()
        bs, nf = f2.shape[0], f2.shape[1]
        seen, _ = ops.face_visibility(fm.unsqueeze(1), None, nf)
        seen = seen[:, 0].bool()
        no_bg = ~(fm == -1).flatten(1).any(1)
        first = seen.int().argmax(1)  # lowest visible face
        drop = no_bg & seen.any(1)
        seen[torch.arange(bs, device=seen.device)[drop], first[drop]] = False
        out = torch.where(seen[:, :, None, None], f2, torch.full_like(f2, -2.0))
        return out[0] if single else out

    # ---- src/cal_flow.py:28-35 as one fused call (no source raster, no [B,F,3,3] round trip)
    def cal_flow(self, src_cam, src_vertices, tgt_cam, tgt_vertices, return_maps=False):
        return ops.cal_flow(src_cam.contiguous(), src_vertices.contiguous(), tgt_cam.contiguous(),
                            tgt_vertices.contiguous(), self.faces.int(), self.image_size,
                            eye_z=SMPLRenderer._eye_z_f32(self), return_maps=return_maps)
