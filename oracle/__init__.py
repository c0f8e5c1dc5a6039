"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

numpy/ctypes front end of the CPU oracle (``oracle/jaf_oracle.c``), the plain-C
restatement of the reference's arithmetic for the appearance warp-and-fuse hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as
the checker or the CPU baseline.  Nothing under ``jafpro_b200/`` imports it.

Each wrapper names the reference file:line its C function follows; the full
citations live next to the C code.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libjaf_oracle.so")
_lib = None

# float32 value of the look_at eye z used by SMPLRenderer (src/nmr.py:177):
# eye = [0, 0, -(1/tan(radians(30)) + 1)]
EYE_Z = float(np.float32(-(1.0 / math.tan(math.radians(30.0)) + 1.0)))


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc only; no GPU, no reference needed)."""
    src = os.path.join(_HERE, "jaf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libjaf_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def build_ref() -> str | None:
    """Compile the reference's own rasteriser kernels (needs /root/reference + nvcc).

    Returns the path of oracle/_ref/libjaf_ref_raster.so, or None when the
    reference tree is not present (e.g. on the GPU box, where the prebuilt file
    travels with the snapshot)."""
    out = os.path.join(_HERE, "_ref", "libjaf_ref_raster.so")
    if os.path.isdir("/root/reference/third_party/neural_renderer"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return out if os.path.exists(out) else None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


# --------------------------------------------------------------------------- a1-a3
def project_gather(cam, verts, faces_idx, eye_z: float = EYE_Z):
    """src/nmr.py:10-28,:263-276; NR/look_at.py:6-62; NR/vertices_to_faces.py:4-22."""
    cam, verts, faces_idx = _f32(cam), _f32(verts), _i32(faces_idx)
    B, V, _ = verts.shape
    F = faces_idx.shape[0]
    out = np.empty((B, F, 3, 3), np.float32)
    lib().orc_project_gather(_p(cam), _p(verts), _p(faces_idx), B, V, F, C.c_float(eye_z), _p(out))
    return out


# --------------------------------------------------------------------------- a4-a6
def raster_fim_wim(faces_xyz, image_size: int, near: float = 0.1, far: float = 100.0,
                   flip_rows: bool = True, return_depth: bool = False):
    """NR/rasterize.py:543-571 -> NR/cuda/rasterize_cuda_kernel.cu:24-169 (+ flips :334-338)."""
    faces_xyz = _f32(faces_xyz)
    B, F = faces_xyz.shape[:2]
    fim = np.empty((B, image_size, image_size), np.int32)
    wim = np.empty((B, image_size, image_size, 3), np.float32)
    depth = np.empty((B, image_size, image_size), np.float32) if return_depth else None
    lib().orc_raster_fim_wim(_p(faces_xyz), B, F, image_size, C.c_float(near), C.c_float(far),
                             int(bool(flip_rows)), _p(fim), _p(wim), _p(depth))
    return (fim, wim, depth) if return_depth else (fim, wim)


def render_fim_wim(cam, verts, faces_idx, image_size: int = 256):
    """src/nmr.py:263-278 -> (faces_xyz, fim, wim)."""
    faces_xyz = project_gather(cam, verts, faces_idx)
    fim, wim = raster_fim_wim(faces_xyz, image_size)
    return faces_xyz, fim, wim


# --------------------------------------------------------------------------- a9
def flow_compose(src_pts, fim, wim, negate_y: bool = False):
    """src/nmr.py:617-659 (cal_bc_transform).  src_pts [B,F,3,2] or [B,F,3,3]."""
    src_pts, fim, wim = _f32(src_pts), _i32(fim), _f32(wim)
    B, F = src_pts.shape[:2]
    stride = src_pts.shape[3]
    H, W = fim.shape[1:]
    T = np.empty((B, H, W, 2), np.float32)
    lib().orc_flow_compose(_p(src_pts), stride, int(bool(negate_y)), _p(fim), _p(wim), B, F, H * W, _p(T))
    return T


def cal_flow(src_cam, src_verts, tgt_cam, tgt_verts, faces_idx, image_size: int = 256):
    """src/cal_flow.py:28-35 -> T [B,H,W,2] (plus the target fim/wim it was built from)."""
    src_faces = project_gather(src_cam, src_verts, faces_idx)
    # cal_flow.py:30-31: keep x,y and negate y
    tgt_faces = project_gather(tgt_cam, tgt_verts, faces_idx)
    fim, wim = raster_fim_wim(tgt_faces, image_size)
    T = flow_compose(src_faces, fim, wim, negate_y=True)
    return T, fim, wim


# --------------------------------------------------------------------------- a10
def grid_sample_border(src, grid, align_corners: bool = False):
    """src/cal_flow.py:37-39 == F.grid_sample(src, grid, 'bilinear', 'border', align_corners)."""
    src, grid = _f32(src), _f32(grid)
    N, Cc, Hs, Ws = src.shape
    H, W = grid.shape[1:3]
    out = np.empty((N, Cc, H, W), np.float32)
    lib().orc_grid_sample_border(_p(src), _p(grid), N, Cc, Hs, Ws, H, W, int(bool(align_corners)), _p(out))
    return out


# --------------------------------------------------------------------------- row F
class WarpFuseParams(C.Structure):
    """Mirror of ``orc_warp_fuse_params`` — and, field for field, of the product's
    ``JafWarpFuseParams`` in include/jafpro_b200.h."""
    _fields_ = [
        ("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Hs", C.c_int32), ("Ws", C.c_int32), ("C", C.c_int32),
        ("align_corners", C.c_int32), ("feat_layout", C.c_int32), ("feat_dtype", C.c_int32),
        ("mask_c", C.c_int32), ("reserved", C.c_int32),
        ("rgb", C.c_void_p), ("feat", C.c_void_p), ("ref_index", C.c_void_p),
        ("grid", C.c_void_p), ("logits", C.c_void_p), ("vis", C.c_void_p), ("fim", C.c_void_p),
        ("tgt_mask", C.c_void_p), ("fake", C.c_void_p), ("conf", C.c_void_p),
        ("out_rgb", C.c_void_p), ("out_feat", C.c_void_p), ("warped_rgb", C.c_void_p),
        ("stream", C.c_void_p),
    ]


def f32_to_bf16_bits(a):
    """Round-to-nearest-even float32 -> bfloat16 bit patterns (uint16)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16).astype(np.uint16)
    return r


def bf16_bits_to_f32(h):
    return (np.ascontiguousarray(h, np.uint16).astype(np.uint32) << 16).view(np.float32)


def warp_fuse(grid, rgb=None, feat=None, *, feat_layout="planar", feat_bf16=False, logits=None,
              vis=None, fim=None, tgt_mask=None, fake=None, conf=None, ref_index=None,
              align_corners=False, return_warped=False):
    """SURVEY.md §8a row F (composition of cal_flow.py:37-39, networks.py:1230-1286,
    flow_net.py:91,:98).

    grid [B,K,H,W,2]; rgb [R,K,3,Hs,Ws] f32; feat [R,K,C,Hs,Ws] (planar) or
    [R,K,Hs,Ws,C] (channels-last), float32 or — with feat_bf16 — uint16 bf16 bits.
    Returns dict(out_rgb, out_feat, warped_rgb)."""
    grid = _f32(grid)
    B, K, H, W, _ = grid.shape
    q = WarpFuseParams()
    keep = [grid]
    q.B, q.K, q.H, q.W = B, K, H, W
    q.align_corners = int(bool(align_corners))
    q.grid = _p(grid)
    out = {"out_rgb": None, "out_feat": None, "warped_rgb": None}
    Hs = Ws = None
    if rgb is not None:
        rgb = _f32(rgb)
        Hs, Ws = rgb.shape[-2:]
        q.rgb = _p(rgb)
        out["out_rgb"] = np.empty((B, 3, H, W), np.float32)
        q.out_rgb = _p(out["out_rgb"])
        if return_warped:
            out["warped_rgb"] = np.empty((B, K, 3, H, W), np.float32)
            q.warped_rgb = _p(out["warped_rgb"])
        keep.append(rgb)
    if feat is not None:
        feat = np.ascontiguousarray(feat, np.uint16 if feat_bf16 else np.float32)
        cl = feat_layout in ("nhwc", "channels_last", 1)
        if cl:
            Hs, Ws, Cc = feat.shape[-3:]
            oshape = (B, H, W, Cc)
        else:
            Cc, Hs, Ws = feat.shape[-3:]
            oshape = (B, Cc, H, W)
        q.C, q.feat_layout, q.feat_dtype = Cc, int(cl), int(bool(feat_bf16))
        q.feat = _p(feat)
        out["out_feat"] = np.empty(oshape, feat.dtype)
        q.out_feat = _p(out["out_feat"])
        keep.append(feat)
    q.Hs, q.Ws = Hs, Ws
    for name, arr, conv in (("logits", logits, _f32), ("vis", vis, _f32), ("fim", fim, _i32),
                            ("tgt_mask", tgt_mask, _f32), ("fake", fake, _f32), ("conf", conf, _f32),
                            ("ref_index", ref_index, _i32)):
        if arr is not None:
            a = conv(arr)
            keep.append(a)
            setattr(q, name, _p(a))
            if name == "tgt_mask":
                q.mask_c = a.shape[1]
    rc = lib().orc_warp_fuse(C.byref(q))
    if rc != 0:
        raise ValueError(f"orc_warp_fuse failed: {rc}")
    return out


# --------------------------------------------------------------------------- a11
def mask_blend(tsf, mask=None, fake=None, conf=None):
    """src/flow_net.py:91 (tsf*mask) and :98 (fake*w + tsf*(1-w)).  Returns (masked, pred|None)."""
    tsf = _f32(tsf)
    B, Cc, H, W = tsf.shape
    mask, fake, conf = _f32(mask), _f32(fake), _f32(conf)
    masked = np.empty_like(tsf)
    pred = np.empty_like(tsf) if conf is not None else None
    lib().orc_mask_blend(_p(fake), _p(tsf), _p(mask), 1 if mask is None else mask.shape[1], _p(conf),
                         B, Cc, C.c_long(H * W), _p(masked), _p(pred))
    return masked, pred


# --------------------------------------------------------------------------- a12
def softmax_fuse(feat, logits):
    """src/networks.py:1264-1286.  feat [B,K*C,h,w], logits [B,K,h,w] -> [B,C,h,w]."""
    feat, logits = _f32(feat), _f32(logits)
    B, KC, H, W = feat.shape
    K = logits.shape[1]
    Cc = KC // K
    out = np.empty((B, Cc, H, W), np.float32)
    lib().orc_softmax_fuse(_p(feat), _p(logits), B, K, Cc, C.c_long(H * W), _p(out))
    return out


# --------------------------------------------------------------------------- a13
def convlstm_step(x, h, c, weight, bias=None):
    """src/convLSTM.py:41-56.  Returns (h_next, c_next)."""
    x, h, c, weight, bias = _f32(x), _f32(h), _f32(c), _f32(weight), _f32(bias)
    B, Cin, H, W = x.shape
    Ch = h.shape[1]
    kh, kw = weight.shape[2:]
    h2, c2 = np.empty_like(h), np.empty_like(c)
    lib().orc_convlstm_step(_p(x), _p(h), _p(c), _p(weight), _p(bias), B, Cin, Ch, H, W, kh, kw,
                            _p(h2), _p(c2))
    return h2, c2


# --------------------------------------------------------------------------- O-gpu
class RefRaster:
    """The reference's own CUDA rasteriser kernels (oracle/_ref/, GPU only)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libjaf_ref_raster.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        import torch  # noqa: F401  (libtorch must be loaded first: the shim links c10/torch_cpu)
        self._lib = C.CDLL(path)
        self._lib.jaf_ref_forward_face_index_map.restype = C.c_int

    def raw(self, faces_xyz, image_size: int, near: float = 0.1, far: float = 100.0, return_depth: bool = False):
        """The extension call itself (NR/rasterize.py:161-169) on outputs pre-filled like NR/rasterize.py:50-69,164,
        rows NOT flipped -> (fim, wim, depth, face_inv_map, faces_inv)."""
        import torch
        faces = faces_xyz.contiguous().clone()
        B, F = faces.shape[:2]
        dev = faces.device
        fim = torch.full((B, image_size, image_size), -1, dtype=torch.int32, device=dev)
        wim = torch.zeros((B, image_size, image_size, 3), dtype=torch.float32, device=dev)
        depth = torch.full((B, image_size, image_size), float(far), dtype=torch.float32, device=dev)
        finv_map = torch.zeros((B, image_size, image_size, 3, 3) if return_depth else (1,), dtype=torch.float32, device=dev)
        faces_inv = torch.zeros_like(faces)
        torch.cuda.synchronize()
        rc = self._lib.jaf_ref_forward_face_index_map(
            C.c_void_p(faces.data_ptr()), C.c_void_p(faces_inv.data_ptr()), C.c_void_p(fim.data_ptr()),
            C.c_void_p(wim.data_ptr()), C.c_void_p(depth.data_ptr()), C.c_void_p(finv_map.data_ptr()),
            B, F, image_size, C.c_float(near), C.c_float(far), int(bool(return_depth)))
        if rc != 0:
            raise RuntimeError(f"reference rasteriser failed: cuda error {-rc}")
        return fim, wim, depth, finv_map, faces_inv

    def __call__(self, faces_xyz, image_size: int, near: float = 0.1, far: float = 100.0,
                 flip_rows: bool = True):
        """faces_xyz: CUDA float32 tensor [B,F,3,3].  Follows NR/rasterize.py:37-69,:161-169,:334-338."""
        import torch
        faces = faces_xyz.contiguous().clone()
        B, F = faces.shape[:2]
        dev = faces.device
        fim = torch.full((B, image_size, image_size), -1, dtype=torch.int32, device=dev)
        wim = torch.zeros((B, image_size, image_size, 3), dtype=torch.float32, device=dev)
        depth = torch.full((B, image_size, image_size), float(far), dtype=torch.float32, device=dev)
        finv_map = torch.zeros(1, dtype=torch.float32, device=dev)
        faces_inv = torch.zeros_like(faces)
        torch.cuda.synchronize()
        rc = self._lib.jaf_ref_forward_face_index_map(
            C.c_void_p(faces.data_ptr()), C.c_void_p(faces_inv.data_ptr()), C.c_void_p(fim.data_ptr()),
            C.c_void_p(wim.data_ptr()), C.c_void_p(depth.data_ptr()), C.c_void_p(finv_map.data_ptr()),
            B, F, image_size, C.c_float(near), C.c_float(far), 0)
        if rc != 0:
            raise RuntimeError(f"reference rasteriser failed: cuda error {-rc}")
        if flip_rows:
            fim, wim, depth = torch.flip(fim, dims=(1,)), torch.flip(wim, dims=(1,)), torch.flip(depth, dims=(1,))
        return fim, wim, depth


    def time_kernels(self, faces_xyz, image_size: int, near: float = 0.1, far: float = 100.0, reps: int = 3):
        """Median milliseconds of the reference's forward_face_index_map kernels (1 + 2) on `faces_xyz`, CUDA events
        on the legacy default stream the reference launches on; the fills / clone / flips around them are outside."""
        import torch
        faces = faces_xyz.contiguous().clone()
        B, F = faces.shape[:2]
        dev = faces.device
        fim = torch.full((B, image_size, image_size), -1, dtype=torch.int32, device=dev)
        wim = torch.zeros((B, image_size, image_size, 3), dtype=torch.float32, device=dev)
        depth = torch.full((B, image_size, image_size), float(far), dtype=torch.float32, device=dev)
        finv_map = torch.zeros(1, dtype=torch.float32, device=dev)
        faces_inv = torch.zeros_like(faces)
        ts = []
        legacy = torch.cuda.ExternalStream(0)
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(legacy)
            rc = self._lib.jaf_ref_forward_face_index_map(
                C.c_void_p(faces.data_ptr()), C.c_void_p(faces_inv.data_ptr()), C.c_void_p(fim.data_ptr()),
                C.c_void_p(wim.data_ptr()), C.c_void_p(depth.data_ptr()), C.c_void_p(finv_map.data_ptr()),
                B, F, image_size, C.c_float(near), C.c_float(far), 0)
            e1.record(legacy)
            e1.synchronize()
            if rc != 0:
                raise RuntimeError(f"reference rasteriser failed: cuda error {-rc}")
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[1:])
        return ts[len(ts) // 2]


# --------------------------------------------------------------------------- §8f rank 1
def texture_warp(tex_parts, iuv, align_corners: bool = False):
    """test/conv_pro_test.py:41-74 (texture_warp_pytorch).  tex_parts [P,3,Ht,Wt] f32, iuv [B,H,W,3] or
    [H,W,3] uint8 -> [B,3,H,W] (or [3,H,W])."""
    tex = _f32(tex_parts)
    iuv = np.ascontiguousarray(iuv, np.uint8)
    single = iuv.ndim == 3
    if single:
        iuv = iuv[None]
    B, H, W, _ = iuv.shape
    P, _, Ht, Wt = tex.shape
    out = np.empty((B, 3, H, W), np.float32)
    lib().orc_texture_warp(_p(tex), P, Ht, Wt, _p(iuv), B, H, W, int(bool(align_corners)), _p(out))
    return out[0] if single else out


def face_visibility(fim_src, fim_tgt, num_faces: int):
    """Row F per-reference visibility, numpy restatement of the rule behind SMPLRenderer.get_vis_f2pts
    (src/nmr.py:507-546): seen[b,k,f] = f in fim_src[b,k]; vis[b,k,p] = fim_tgt[b,p] >= 0 and seen[b,k,fim_tgt[b,p]]."""
    fim_src = np.asarray(fim_src)
    B, K = fim_src.shape[:2]
    seen = np.zeros((B, K, num_faces), np.uint8)
    for b in range(B):
        for k in range(K):
            ids = np.unique(fim_src[b, k])
            ids = ids[(ids >= 0) & (ids < num_faces)]
            seen[b, k, ids] = 1
    if fim_tgt is None:
        return seen, None
    fim_tgt = np.asarray(fim_tgt)
    vis = np.zeros(fim_src.shape, np.float32)
    for b in range(B):
        t = fim_tgt[b]
        ok = (t >= 0) & (t < num_faces)
        for k in range(K):
            vis[b, k][ok] = seen[b, k][t[ok]]
    return seen, vis


def get_vis_f2pts(f2pts, fims):
    """SMPLRenderer.get_vis_f2pts (src/nmr.py:507-546), including `fim.unique()[1:]` (:529): the FIRST unique value
    is dropped whatever it is."""
    f2pts, fims = np.asarray(f2pts, np.float32), np.asarray(fims)
    out = np.full_like(f2pts, -2.0)
    for b in range(f2pts.shape[0]):
        ids = np.unique(fims[b])[1:]
        out[b, ids] = f2pts[b, ids]
    return out


def flow_warp_pair(feat_fwd, feat_bwd, base_grid, flow, align_corners: bool = False):
    """One level of SpatioTempoCRN.forward's warps (src/crn_model.py:457-466): F.interpolate(flow, size, 'nearest')
    [source index min(floor(dst * float32(in/out)), in-1), ATen UpSample.h nearest_neighbor_compute_source_index],
    grid +/- flow in fp32, permute, grid_sample border."""
    base_grid, flow = np.asarray(base_grid, np.float32), np.asarray(flow, np.float32)
    B, _, h, w = base_grid.shape
    H, W = flow.shape[2:]
    sy, sx = np.float32(H) / np.float32(h), np.float32(W) / np.float32(w)
    ys = np.minimum(np.floor(np.arange(h, dtype=np.float32) * sy).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(w, dtype=np.float32) * sx).astype(np.int64), W - 1)
    flow_s = flow[:, :, ys][:, :, :, xs]
    outs = []
    for feat, sign in ((feat_fwd, 1.0), (feat_bwd, -1.0)):
        if feat is None:
            outs.append(None)
            continue
        g = (base_grid + np.float32(sign) * flow_s).astype(np.float32)
        outs.append(grid_sample_border(np.asarray(feat, np.float32), np.ascontiguousarray(g.transpose(0, 2, 3, 1)),
                                       align_corners))
    return outs[0], outs[1]


# --------------------------------------------------------------------------- §8f rank 4
def get_texture(im, iuv, tex_size: int = 32, final_size: int = 200):
    """src/utils.py:232-255 restated in numpy: scatter per part (np.where order, last write wins), bilinear resize with
    half-pixel centres (cv::resize INTER_LINEAR geometry, coefficients as exactly rounded doubles), [:, :, ::-1] / 255.
    im, iuv [H,W,3] uint8 -> [24, final, final, 3] float64."""
    im, iuv = np.asarray(im, np.uint8), np.asarray(iuv, np.uint8)
    sf = float(tex_size) - 1
    U, V = iuv[:, :, 1], iuv[:, :, 2]

    def coeffs(sn, dn):
        d = np.arange(dn, dtype=np.int64)
        f = ((2 * d + 1) * sn - dn).astype(np.float64) / np.float64(2 * dn)
        s = np.floor(f).astype(np.int64)
        f = f - s
        lo, hi = s < 0, s >= sn - 1
        s = np.where(lo, 0, np.where(hi, sn - 1, s))
        f = np.where(lo | hi, 0.0, f)
        return s, np.minimum(s + 1, sn - 1), f

    x0, x1, ax = coeffs(tex_size, final_size)
    out = np.zeros((24, final_size, final_size, 3), np.float64)
    for part in range(1, 25):
        x, y = np.where(iuv[:, :, 0] == part)
        if len(x) == 0:
            continue
        a = np.zeros((tex_size, tex_size, 3))
        rows = ((255 - V[x, y]) * sf / 255.).astype(int)
        cols = (U[x, y] * sf / 255.).astype(int)
        for c in range(3):
            a[rows, cols, c] = im[x, y, c]
        top = a[x0][:, x0] * (1.0 - ax)[None, :, None] + a[x0][:, x1] * ax[None, :, None]
        bot = a[x1][:, x0] * (1.0 - ax)[None, :, None] + a[x1][:, x1] * ax[None, :, None]
        r = top * (1.0 - ax)[:, None, None] + bot * ax[:, None, None]
        out[part - 1] = r[:, :, ::-1] / 255.
    return out
