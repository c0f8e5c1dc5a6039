"""Tensor-level wrappers over the C ABI: same tensor signatures as the reference call
sites, CUDA tensors in, CUDA tensors out, launched on torch's current stream.

Error behaviour follows the reference extension (NR/cuda/rasterize_cuda.cpp:66-68,84-89):
a CPU or non-contiguous tensor raises ``RuntimeError``.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib

# float32 value of the look_at eye used by SMPLRenderer (src/nmr.py:177)
EYE_Z = float(torch.tensor(-(1.0 / math.tan(math.radians(30.0)) + 1.0), dtype=torch.float32))
DEFAULT_NEAR, DEFAULT_FAR = 0.1, 100.0  # NR/rasterize.py:10-11 (what src/nmr.py:277 ends up using)

_workspaces: dict = {}


def forbid_grad(*tensors, name: str = "input"):
    """The kernels behind this package are FORWARD-ONLY (the reference runs the flow / warp / fusion path under
    ``torch.no_grad()`` at inference, test/conv_pro_test.py:190).  Outputs are plain tensors without a grad_fn, so a
    drop-in that silently accepted tensors which require grad would stop a training script from learning
    (train/4.convLSTM_flowpro_interval.py:278 trains exactly these cells).  Fail loudly instead."""
    if torch.is_grad_enabled():
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.requires_grad:
                raise RuntimeError(
                    f"jafpro_b200 is forward-only: {name} requires grad while autograd is enabled. Call it under "
                    "torch.no_grad() / torch.inference_mode() (inference), or keep the reference's torch "
                    "implementation for training.")


def _check(t, name, dtype=None):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    forbid_grad(t, name=name)
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _ptr(t):
    # plain ints: ctypes converts them for c_void_p parameters (argtypes are declared in _lib.SIGNATURES)
    return None if t is None else t.data_ptr()


try:  # the raw handle of torch's current stream without building a torch.cuda.Stream object (~3 us per call saved)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    _raw_stream = None


def _stream():
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _on:
    """Device guard that costs nothing when the tensor already lives on the current device."""
    __slots__ = ("dev", "prev")

    def __init__(self, dev):
        self.dev = dev

    def __enter__(self):
        cur = torch.cuda.current_device()
        idx = self.dev.index if self.dev.index is not None else cur
        self.prev = cur if idx != cur else -1
        if self.prev >= 0:
            torch.cuda.set_device(idx)

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


_clean_keys: dict = {}  # workspace key -> number of leading z-buffer keys known to be empty (self-cleaning fused kernel)


def _workspace(dev, nbytes, keep_clean: bool = False):
    """Grow-only scratch per device (z-buffer keys).  Stream-ordered use only.  The buffer never leaves this module, so
    the pose-driven kernel's "keys left empty" protocol (JAF_POSES_*) can be tracked here: every other user dirties it."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
        _clean_keys[key] = 0
    if not keep_clean:
        _clean_keys[key] = 0
    return buf


# ----------------------------------------------------------------------------- a1-a3
def project_gather(cam, vertices, faces_idx, eye_z: float = EYE_Z):
    """proj_func + y flip + look_at + vertices_to_faces (src/nmr.py:266-276) -> [B,F,3,3]."""
    cam, vertices = _check(cam, "cam", torch.float32), _check(vertices, "vertices", torch.float32)
    faces_idx = _check(faces_idx, "faces", torch.int32)
    B, V, _ = vertices.shape
    F = faces_idx.shape[-2]
    out = torch.empty((B, F, 3, 3), dtype=torch.float32, device=vertices.device)
    with _on(vertices.device):
        _lib.check(_lib.lib().jaf_project_gather(_ptr(cam), _ptr(vertices), _ptr(faces_idx), B, V, F, eye_z,
                                                 _ptr(out), _stream()), "project_gather")
    return out


# ----------------------------------------------------------------------------- a4-a6
def raster_fim_wim(faces, image_size: int, near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR,
                   flip_rows: bool = True, return_depth: bool = False):
    """nr.rasterize_face_index_map_and_weight_map (NR/rasterize.py:543-571), anti_aliasing=False."""
    faces = _check(faces, "faces", torch.float32)
    if faces.dim() != 4 or faces.shape[2:] != (3, 3):
        raise RuntimeError("faces must be [batch, num_faces, 3, 3]")
    B, F = faces.shape[:2]
    dev = faces.device
    fim = torch.empty((B, image_size, image_size), dtype=torch.int32, device=dev)
    wim = torch.empty((B, image_size, image_size, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((B, image_size, image_size), dtype=torch.float32, device=dev) if return_depth else None
    with _on(dev):
        ws = _workspace(dev, _lib.lib().jaf_raster_workspace_bytes(B, image_size))
        _lib.check(_lib.lib().jaf_raster_fim_wim(_ptr(faces), B, F, image_size, near, far, int(flip_rows), _ptr(fim),
                                                 _ptr(wim), _ptr(depth), _ptr(ws), _stream()), "raster_fim_wim")
    return (fim, wim, depth) if return_depth else (fim, wim)


# ----------------------------------------------------------------------------- a7
def render_fim_wim(cam, vertices, faces_idx, image_size: int, eye_z: float = EYE_Z, near: float = DEFAULT_NEAR,
                   far: float = DEFAULT_FAR, return_faces: bool = True):
    """SMPLRenderer.render_fim_wim (src/nmr.py:263-278) in one call -> (faces|None, fim, wim)."""
    cam, vertices = _check(cam, "cam", torch.float32), _check(vertices, "vertices", torch.float32)
    faces_idx = _check(faces_idx, "faces", torch.int32)
    B, V, _ = vertices.shape
    F = faces_idx.shape[-2]
    dev = vertices.device
    faces = torch.empty((B, F, 3, 3), dtype=torch.float32, device=dev) if return_faces else None
    fim = torch.empty((B, image_size, image_size), dtype=torch.int32, device=dev)
    wim = torch.empty((B, image_size, image_size, 3), dtype=torch.float32, device=dev)
    with _on(dev):
        ws = _workspace(dev, _lib.lib().jaf_raster_workspace_bytes(B, image_size))
        _lib.check(_lib.lib().jaf_render_fim_wim(_ptr(cam), _ptr(vertices), _ptr(faces_idx), B, V, F, image_size,
                                                 eye_z, near, far, _ptr(faces), _ptr(fim), _ptr(wim), _ptr(ws),
                                                 _stream()), "render_fim_wim")
    return faces, fim, wim


# ----------------------------------------------------------------------------- a9
def flow_compose(src_f2pts, dst_fims, dst_wims, negate_y: bool = False):
    """SMPLRenderer.cal_bc_transform (src/nmr.py:617-659).  src_f2pts [B,F,3,2] (or [B,F,3,3])."""
    src = _check(src_f2pts, "src_f2pts", torch.float32)
    fim, wim = _check(dst_fims, "dst_fims", torch.int32), _check(dst_wims, "dst_wims", torch.float32)
    B, F = src.shape[:2]
    H, W = fim.shape[1:]
    T = torch.empty((B, H, W, 2), dtype=torch.float32, device=src.device)
    with _on(src.device):
        _lib.check(_lib.lib().jaf_flow_compose(_ptr(src), src.shape[3], int(negate_y), _ptr(fim), _ptr(wim), B, F, H,
                                               W, _ptr(T), _stream()), "flow_compose")
    return T


# ----------------------------------------------------------------------------- a8
def cal_flow(src_cam, src_vertices, tgt_cam, tgt_vertices, faces_idx, image_size: int, eye_z: float = EYE_Z,
             near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR, return_maps: bool = False):
    """float_estimate.cal_flow (src/cal_flow.py:28-35) fused -> T [B,S,S,2] (and fim, wim)."""
    sc, sv = _check(src_cam, "src_cam", torch.float32), _check(src_vertices, "src_vertices", torch.float32)
    tc, tv = _check(tgt_cam, "tgt_cam", torch.float32), _check(tgt_vertices, "tgt_vertices", torch.float32)
    faces_idx = _check(faces_idx, "faces", torch.int32)
    B, V, _ = tv.shape
    F = faces_idx.shape[-2]
    dev = tv.device
    T = torch.empty((B, image_size, image_size, 2), dtype=torch.float32, device=dev)
    fim = torch.empty((B, image_size, image_size), dtype=torch.int32, device=dev) if return_maps else None
    wim = torch.empty((B, image_size, image_size, 3), dtype=torch.float32, device=dev) if return_maps else None
    with _on(dev):
        ws = _workspace(dev, _lib.lib().jaf_raster_workspace_bytes(B, image_size))
        _lib.check(_lib.lib().jaf_cal_flow(_ptr(sc), _ptr(sv), _ptr(tc), _ptr(tv), _ptr(faces_idx), B, V, F,
                                           image_size, eye_z, near, far, _ptr(T), _ptr(fim), _ptr(wim), _ptr(ws),
                                           _stream()), "cal_flow")
    return (T, fim, wim) if return_maps else T


# ----------------------------------------------------------------------------- row F
def _feat_layout(feat):
    """-> (layout, dtype code, C, Hs, Ws, dense tensor).  Accepts [R,K,C,H,W] contiguous (planar) or the
    same logical shape with channels-last strides (dense [R,K,H,W,C] memory)."""
    if feat.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("features must be float32 or bfloat16")
    if not feat.is_cuda:
        raise RuntimeError("feat must be a CUDA tensor")
    forbid_grad(feat, name="feat")
    dt = 0 if feat.dtype == torch.float32 else 1
    if feat.dim() != 5:
        raise RuntimeError("feat must be [R, K, C, H, W]")
    R, K, Cc, H, W = feat.shape
    if feat.is_contiguous():
        return 0, dt, Cc, H, W
    if feat.permute(0, 1, 3, 4, 2).is_contiguous():
        return 1, dt, Cc, H, W
    raise RuntimeError("feat must be contiguous [R,K,C,H,W] or channels-last (dense [R,K,H,W,C])")


def warp_fuse(grid, rgb=None, feat=None, *, logits=None, vis=None, fim=None, tgt_mask=None, fake=None, conf=None,
              ref_index=None, align_corners: bool = False, return_warped: bool = False, out_rgb=None, out_feat=None):
    """Fused K-reference warp + fusion (include/jafpro_b200.h).

    grid [B,K,H,W,2] f32; rgb [R,K,3,Hs,Ws] f32; feat [R,K,C,Hs,Ws] f32/bf16 (contiguous, or with
    channels-last strides — the fast path); logits/vis [B,K,H,W]; fim [B,H,W] int32;
    tgt_mask [B,1|3,H,W]; fake [B,3,H,W]; conf [B,1,H,W]; ref_index [B] int32.
    Returns (out_rgb|None, out_feat|None[, warped_rgb]).  out_feat has feat's dtype and memory format."""
    grid = _check(grid, "grid", torch.float32)
    if grid.dim() != 5 or grid.shape[-1] != 2:
        raise RuntimeError("grid must be [B, K, H, W, 2]")
    B, K, H, W, _ = grid.shape
    dev = grid.device
    q = _lib.WarpFuseParams()
    q.B, q.K, q.H, q.W = B, K, H, W
    q.align_corners = int(bool(align_corners))
    q.grid = grid.data_ptr()
    warped = None
    Hs = Ws = None
    if rgb is not None:
        rgb = _check(rgb, "rgb", torch.float32)
        if rgb.dim() != 5 or rgb.shape[1] != K or rgb.shape[2] != 3:
            raise RuntimeError("rgb must be [R, K, 3, Hs, Ws]")
        Hs, Ws = rgb.shape[-2:]
        if out_rgb is None:
            out_rgb = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        elif tuple(_check(out_rgb, "out_rgb", torch.float32).shape) != (B, 3, H, W):
            raise RuntimeError("out_rgb must be [B, 3, H, W]")
        q.rgb, q.out_rgb = rgb.data_ptr(), out_rgb.data_ptr()
        if return_warped:
            warped = torch.empty((B, K, 3, H, W), dtype=torch.float32, device=dev)
            q.warped_rgb = warped.data_ptr()
    if feat is not None:
        layout, dt, Cc, fh, fw = _feat_layout(feat)
        if feat.shape[1] != K:
            raise RuntimeError("feat must be [R, K, C, Hs, Ws]")
        if Hs is not None and (fh, fw) != (Hs, Ws):
            raise RuntimeError("rgb and feat must share the reference size")
        Hs, Ws = fh, fw
        if out_feat is not None:  # caller-allocated (the reference extension's convention): same layout as feat
            if out_feat.dtype != feat.dtype or tuple(out_feat.shape) != (B, Cc, H, W) or not out_feat.is_cuda or \
                    not (out_feat.permute(0, 2, 3, 1) if layout == 1 else out_feat).is_contiguous():
                raise RuntimeError("out_feat must be [B, C, H, W] with feat's dtype and memory format")
        elif layout == 1:
            out_feat = torch.empty((B, H, W, Cc), dtype=feat.dtype, device=dev).permute(0, 3, 1, 2)
        else:
            out_feat = torch.empty((B, Cc, H, W), dtype=feat.dtype, device=dev)
        q.C, q.feat_layout, q.feat_dtype = Cc, layout, dt
        q.feat, q.out_feat = feat.data_ptr(), out_feat.data_ptr()
    if Hs is None:
        raise RuntimeError("warp_fuse needs rgb and/or feat")
    q.Hs, q.Ws = Hs, Ws
    keep = []
    for name, t, dtype, shape in (("logits", logits, torch.float32, (B, K, H, W)),
                                  ("vis", vis, torch.float32, (B, K, H, W)),
                                  ("fim", fim, torch.int32, (B, H, W)),
                                  ("fake", fake, torch.float32, (B, 3, H, W)),
                                  ("conf", conf, torch.float32, (B, 1, H, W)),
                                  ("ref_index", ref_index, torch.int32, (B,))):
        if t is not None:
            t = _check(t, name, dtype)
            if tuple(t.shape) != shape:
                raise RuntimeError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
            setattr(q, name, t.data_ptr())
            keep.append(t)
    if tgt_mask is not None:
        tgt_mask = _check(tgt_mask, "tgt_mask", torch.float32)
        if tgt_mask.dim() != 4 or tgt_mask.shape[0] != B or tgt_mask.shape[1] not in (1, 3) or \
                tuple(tgt_mask.shape[2:]) != (H, W):
            raise RuntimeError("tgt_mask must be [B, 1 or 3, H, W]")
        q.tgt_mask, q.mask_c = tgt_mask.data_ptr(), tgt_mask.shape[1]
    if ref_index is None:
        for t in (rgb, feat):
            if t is not None and t.shape[0] != B:
                raise RuntimeError("without ref_index the reference tensors need one set per target frame")
    with _on(dev):
        q.stream = _stream()
        _lib.check(_lib.lib().jaf_warp_fuse(C.byref(q)), "warp_fuse")
    if return_warped:
        return out_rgb, out_feat, warped
    return out_rgb, out_feat


def warp_fuse_from_poses(src_cam, src_vertices, tgt_cam, tgt_vertices, faces_idx, image_size: int, rgb=None, feat=None, *,
                         logits=None, tgt_mask=None, fake=None, conf=None, ref_index=None, align_corners: bool = False,
                         eye_z: float = EYE_Z, near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR,
                         return_flow: bool = False, out_rgb=None, out_feat=None):
    """Rows a7-a12 from the poses in one pass (``jaf_warp_fuse_from_poses``): the transfer flows of the K reference
    poses into every target pose are composed inside the warp kernel, in shared memory; visibility is "the target
    pixel is on the body".  src_cam [R,K,3], src_vertices [R,K,V,3] (indexed like the reference sets: ref_index[b] or
    b), tgt_cam [B,3], tgt_vertices [B,V,3]; references / logits / masks as in `warp_fuse`.
    -> (out_rgb|None, out_feat) and, with return_flow, also (T [B,K,S,S,2], fim [B,S,S]) — bit-identical to
    `cal_flow_multi`.  Shapes the fused kernel does not serve take `cal_flow_multi` + `warp_fuse` (same results)."""
    sc, sv = _check(src_cam, "src_cam", torch.float32), _check(src_vertices, "src_vertices", torch.float32)
    tc, tv = _check(tgt_cam, "tgt_cam", torch.float32), _check(tgt_vertices, "tgt_vertices", torch.float32)
    faces_idx = _check(faces_idx, "faces", torch.int32)
    if sv.dim() != 4 or sc.dim() != 3 or sv.shape[:2] != sc.shape[:2] or tv.dim() != 3:
        raise RuntimeError("expected src_cam [R,K,3], src_vertices [R,K,V,3], tgt_cam [B,3], tgt_vertices [B,V,3]")
    B, V, _ = tv.shape
    R, K = sv.shape[:2]
    S = int(image_size)
    dev = tv.device
    if feat is None and rgb is None:
        raise RuntimeError("warp_fuse_from_poses needs rgb and/or feat")
    if feat is not None:
        layout, dt, Cc, Hs, Ws = _feat_layout(feat)
        if feat.shape[0] != R or feat.shape[1] != K:
            raise RuntimeError("feat must be [R, K, C, Hs, Ws] with the R, K of the source poses")
    else:  # RGB planes only
        layout, dt, Cc = 1, 1, 0
        Hs, Ws = rgb.shape[-2:]
    if ref_index is None and R != B:
        raise RuntimeError("without ref_index the source poses / reference sets need one entry per target frame")
    if not _lib.lib().jaf_warp_fuse_from_poses_supported(Cc, K, layout, dt):
        idx = ref_index.long() if ref_index is not None else None
        T, fim, _ = cal_flow_multi(sc if idx is None else sc[idx].contiguous(), sv if idx is None else sv[idx].contiguous(),
                                   tc, tv, faces_idx, S, eye_z, near, far, return_wim=False)
        o = warp_fuse(T, rgb=rgb, feat=feat, logits=logits, fim=fim, tgt_mask=tgt_mask, fake=fake, conf=conf,
                      ref_index=ref_index, align_corners=align_corners, out_rgb=out_rgb, out_feat=out_feat)
        return (o[0], o[1], T, fim) if return_flow else o
    q = _lib.WarpFuseParams()
    q.B, q.K, q.H, q.W, q.Hs, q.Ws, q.C = B, K, S, S, Hs, Ws, Cc
    q.align_corners = int(bool(align_corners))
    q.feat_layout, q.feat_dtype = layout, dt
    if rgb is not None:
        rgb = _check(rgb, "rgb", torch.float32)
        if tuple(rgb.shape) != (R, K, 3, Hs, Ws):
            raise RuntimeError("rgb must be [R, K, 3, Hs, Ws]")
        if out_rgb is None:
            out_rgb = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
        q.rgb, q.out_rgb = rgb.data_ptr(), _check(out_rgb, "out_rgb", torch.float32).data_ptr()
    if feat is not None:
        if out_feat is None:
            out_feat = torch.empty((B, S, S, Cc), dtype=feat.dtype, device=dev).permute(0, 3, 1, 2)
        q.feat, q.out_feat = feat.data_ptr(), out_feat.data_ptr()
    keep = []
    for name, t, dtype, shape in (("logits", logits, torch.float32, (B, K, S, S)), ("fake", fake, torch.float32, (B, 3, S, S)),
                                  ("conf", conf, torch.float32, (B, 1, S, S)), ("ref_index", ref_index, torch.int32, (B,))):
        if t is not None:
            t = _check(t, name, dtype)
            if tuple(t.shape) != shape:
                raise RuntimeError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
            setattr(q, name, t.data_ptr())
            keep.append(t)
    if tgt_mask is not None:
        tgt_mask = _check(tgt_mask, "tgt_mask", torch.float32)
        if tgt_mask.dim() != 4 or tgt_mask.shape[0] != B or tgt_mask.shape[1] not in (1, 3) or tuple(tgt_mask.shape[2:]) != (S, S):
            raise RuntimeError("tgt_mask must be [B, 1 or 3, S, S]")
        q.tgt_mask, q.mask_c = tgt_mask.data_ptr(), tgt_mask.shape[1]
    T = torch.empty((B, K, S, S, 2), dtype=torch.float32, device=dev) if return_flow else None
    fim = torch.empty((B, S, S), dtype=torch.int32, device=dev) if return_flow else None
    pq = _lib.PoseFlowParams()
    pq.tgt_cam, pq.tgt_verts, pq.src_cam, pq.src_verts = tc.data_ptr(), tv.data_ptr(), sc.data_ptr(), sv.data_ptr()
    pq.faces_idx, pq.V, pq.F = faces_idx.data_ptr(), V, faces_idx.shape[-2]
    pq.eye_z, pq.near_, pq.far_ = eye_z, near, far
    pq.T, pq.fim = _ptr(T), _ptr(fim)
    with _on(dev):
        ws = _workspace(dev, _lib.lib().jaf_raster_workspace_bytes(B, S), keep_clean=True)
        wkey = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
        # the kernel leaves every key it consumed empty: the next call on this (module-owned) workspace skips the clear.
        # Not during CUDA-graph capture: a replayed graph must not depend on what ran between the replays.
        capturing = torch.cuda.is_current_stream_capturing()
        clean = (not capturing) and _clean_keys.get(wkey, 0) >= B * S * S
        pq.flags = 0 if capturing else (2 | (1 if clean else 0))  # JAF_POSES_LEAVE_CLEAN | JAF_POSES_KEYS_CLEAN
        _clean_keys[wkey] = 0
        pq.workspace = ws.data_ptr()
        q.stream = _stream()
        _lib.check(_lib.lib().jaf_warp_fuse_from_poses(C.byref(q), C.byref(pq)), "warp_fuse_from_poses")
        if not capturing:
            _clean_keys[wkey] = B * S * S
    return (out_rgb, out_feat, T, fim) if return_flow else (out_rgb, out_feat)


_branch_streams: dict = {}


def run_concurrently(fns):
    """Run independent callables (each a sequence of ops of this module on tensors it owns) on one side stream each and
    join them back into the current stream -> list of their results.  The five pyramid levels of
    Accumulate_LSTM_no_loss (src/networks.py:1346-1355: five separate ConvLSTMs per step) are such callables: the tails
    of the big levels overlap the small ones (0.378 -> 0.369 ms per pyramid step, 0.360 ms captured in a FrameGraph).
    Capture-safe (the fork / join are event waits); workspaces are per stream."""
    cur = torch.cuda.current_stream()
    dev = cur.device
    pool = _branch_streams.setdefault(dev.index, [])
    while len(pool) < len(fns):
        pool.append(torch.cuda.Stream(device=dev))
    outs = []
    for f, s in zip(fns, pool):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(f())
    for s in pool[:len(fns)]:
        cur.wait_stream(s)

    def _handoff(o):  # results were allocated on a side stream and are consumed on the current one
        if isinstance(o, torch.Tensor):
            o.record_stream(cur)
        elif isinstance(o, (list, tuple)):
            for e in o:
                _handoff(e)
    _handoff(outs)
    return outs


class FrameGraph:
    """Capture a fixed call sequence of this module once into a CUDA graph and replay it with one launch.

    The reference's inference loop runs the path per frame at batch 1 (test/conv_pro_test.py:255-278: cal_flow ->
    warp_image -> mask / blend), where Python + launch overhead dominates the few microseconds of GPU work
    (SURVEY §7 hard part 4).  `fn` is any closure over ops of this module with fixed tensors (write new inputs
    into those tensors with copy_ before `replay`); every kernel behind the C ABI is capture-safe (no allocation,
    synchronisation or host round trip inside a call)."""

    def __init__(self, fn, warmup: int = 3, *, frames: int | None = None, lanes: int = 1):
        """`fn()` is captured as it is.  With ``frames=n`` the callable is ``fn(t)`` for t in range(n) — n INDEPENDENT
        frames (the reference's loop does not feed frame t-1 into frame t, conv_pro_test.py:255-278) — and frame t is
        captured on lane ``t % lanes``: the graph gets `lanes` parallel branches, so the few-microsecond kernels of
        several batch-1 frames overlap on the GPU instead of queueing behind each other.  Workspaces are per stream
        (ops._workspace), so the branches do not share scratch."""
        self.frames, self.lanes = frames, max(1, int(lanes))
        self._fn = fn
        self._lane_streams = [torch.cuda.Stream() for _ in range(self.lanes)] if (frames is not None and self.lanes > 1) else []
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up off the capture: one-time initialisation (workspaces, attributes)
            for _ in range(warmup):
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self):
        if self.frames is None:
            return self._fn()
        if not self._lane_streams:
            return [self._fn(t) for t in range(self.frames)]
        cur = torch.cuda.current_stream()
        outs = [None] * self.frames
        for s in self._lane_streams:   # fork
            s.wait_stream(cur)
        for t in range(self.frames):
            with torch.cuda.stream(self._lane_streams[t % self.lanes]):
                outs[t] = self._fn(t)
        for s in self._lane_streams:   # join
            cur.wait_stream(s)
        return outs

    def replay(self):
        self.graph.replay()
        return self.out


def _host(t, name, dtype=None, shape=None):
    """Host-side argument check of the *_host entry points: raw pointers go straight into pipelined memcpys, so a
    wrong dtype / stride / shape would read or write out of bounds."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or t.is_cuda:
        raise RuntimeError(f"{name} must be a host tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    forbid_grad(t, name=name)
    return t


def _host_refs(q, K, rgb, feat, feat_channels_last, B, ref_index):
    """Fills the reference-set fields of a host-side parameter block -> (R, C)."""
    R = Cc = None
    if rgb is not None:
        rgb = _host(rgb, "rgb", torch.float32)
        if rgb.dim() != 5 or rgb.shape[1] != K or rgb.shape[2] != 3:
            raise RuntimeError("rgb must be [R, K, 3, Hs, Ws]")
        R = rgb.shape[0]
        q.Hs, q.Ws = rgb.shape[-2:]
        q.rgb = rgb.data_ptr()
    if feat is not None:
        feat = _host(feat, "feat")
        if feat.dtype not in (torch.float32, torch.bfloat16) or feat.dim() != 5 or feat.shape[1] != K:
            raise RuntimeError("feat must be a float32 / bfloat16 [R,K,C,Hs,Ws] (or dense [R,K,Hs,Ws,C]) host tensor")
        if feat_channels_last:
            fh, fw, Cc = feat.shape[-3:]
        else:
            Cc, fh, fw = feat.shape[-3:]
        if rgb is not None and ((fh, fw) != tuple(rgb.shape[-2:]) or feat.shape[0] != R):
            raise RuntimeError("rgb and feat must share the reference count and size")
        R = feat.shape[0]
        q.Hs, q.Ws, q.C = fh, fw, Cc
        q.feat_layout = int(feat_channels_last)
        q.feat_dtype = 0 if feat.dtype == torch.float32 else 1
        q.feat = feat.data_ptr()
    if R is None:
        raise RuntimeError("need rgb and/or feat")
    if ref_index is not None:
        ref_index = _host(ref_index, "ref_index", torch.int32, (B,))
        if int(ref_index.min()) < 0 or int(ref_index.max()) >= R:
            raise RuntimeError("ref_index out of range")
        q.ref_index = ref_index.data_ptr()
    elif R != B:
        raise RuntimeError("without ref_index the reference tensors need one set per target frame")
    return R, Cc


def warp_fuse_host(grid, rgb=None, feat=None, *, feat_channels_last: bool = False, logits=None, vis=None, fim=None,
                   tgt_mask=None, fake=None, conf=None, ref_index=None, align_corners: bool = False,
                   frames_per_chunk: int = 0, out_rgb=None, out_feat=None):
    """``jaf_warp_fuse_host``: every tensor lives in HOST memory (pin it for full PCIe speed); the
    library pipelines H2D, the kernel and D2H itself.  feat is [R,K,C,Hs,Ws], or dense
    [R,K,Hs,Ws,C] when feat_channels_last.  Returns (out_rgb, out_feat) host tensors."""
    grid = _host(grid, "grid", torch.float32)
    if grid.dim() != 5 or grid.shape[-1] != 2:
        raise RuntimeError("grid must be [B, K, H, W, 2]")
    B, K, H, W, _ = grid.shape
    q = _lib.WarpFuseParams()
    q.B, q.K, q.H, q.W = B, K, H, W
    q.align_corners = int(bool(align_corners))
    q.grid = grid.data_ptr()
    _, Cc = _host_refs(q, K, rgb, feat, feat_channels_last, B, ref_index)
    if rgb is not None:
        if out_rgb is None:
            out_rgb = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
        q.out_rgb = _host(out_rgb, "out_rgb", torch.float32, (B, 3, H, W)).data_ptr()
    if feat is not None:
        shape = (B, H, W, Cc) if feat_channels_last else (B, Cc, H, W)
        if out_feat is None:
            out_feat = torch.empty(shape, dtype=feat.dtype).pin_memory()
        q.out_feat = _host(out_feat, "out_feat", feat.dtype, shape).data_ptr()
    for name, t, dtype, shape in (("logits", logits, torch.float32, (B, K, H, W)), ("vis", vis, torch.float32, (B, K, H, W)),
                                  ("fim", fim, torch.int32, (B, H, W)), ("fake", fake, torch.float32, (B, 3, H, W)),
                                  ("conf", conf, torch.float32, (B, 1, H, W))):
        if t is not None:
            setattr(q, name, _host(t, name, dtype, shape).data_ptr())
    if tgt_mask is not None:
        tgt_mask = _host(tgt_mask, "tgt_mask", torch.float32)
        if tgt_mask.dim() != 4 or tgt_mask.shape[0] != B or tgt_mask.shape[1] not in (1, 3) or tuple(tgt_mask.shape[2:]) != (H, W):
            raise RuntimeError("tgt_mask must be [B, 1 or 3, H, W]")
        q.tgt_mask, q.mask_c = tgt_mask.data_ptr(), tgt_mask.shape[1]
    _lib.check(_lib.lib().jaf_warp_fuse_host(C.byref(q), int(frames_per_chunk)), "warp_fuse_host")
    return out_rgb, out_feat


def warp_fuse_from_poses_host(src_cam, src_vertices, tgt_cam, tgt_vertices, faces_idx, image_size: int, rgb=None, feat=None, *,
                              feat_channels_last: bool = True, logits=None, tgt_mask=None, fake=None, conf=None,
                              ref_index=None, align_corners: bool = False, eye_z: float = EYE_Z,
                              near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR, frames_per_chunk: int = 0,
                              out_rgb=None, out_feat=None, out_feat_device=None):
    """``jaf_warp_fuse_from_poses_host``: the pose-driven operation with HOST tensors.  Per target frame only the pose
    (cam [B,3], vertices [B,V,3]), its logits and masks go up and the fused RGB frame comes down; the K references
    (rgb [R,K,3,Hs,Ws], feat dense [R,K,Hs,Ws,C] bf16) and reference poses (src_cam [R,K,3], src_vertices [R,K,V,3]) of a
    video are uploaded once (ref_index [B]).  out_feat_device: a CUDA tensor [B,S,S,C] that receives the fused features
    on the device (they then never cross PCIe); otherwise they are downloaded into out_feat (host).
    Returns (out_rgb host, out_feat host | out_feat_device)."""
    tc, tv = _host(tgt_cam, "tgt_cam", torch.float32), _host(tgt_vertices, "tgt_vertices", torch.float32)
    sc, sv = _host(src_cam, "src_cam", torch.float32), _host(src_vertices, "src_vertices", torch.float32)
    faces_idx = _host(faces_idx, "faces", torch.int32)
    if tv.dim() != 3 or sv.dim() != 4 or sc.dim() != 3 or sv.shape[:2] != sc.shape[:2] or tc.shape != (tv.shape[0], 3):
        raise RuntimeError("expected src_cam [R,K,3], src_vertices [R,K,V,3], tgt_cam [B,3], tgt_vertices [B,V,3]")
    B, V, _ = tv.shape
    R, K = sv.shape[:2]
    S = int(image_size)
    q = _lib.WarpFuseParams()
    q.B, q.K, q.H, q.W = B, K, S, S
    q.align_corners = int(bool(align_corners))
    Rr, Cc = _host_refs(q, K, rgb, feat, feat_channels_last, B, ref_index)
    if feat is None or Rr != R:
        raise RuntimeError("feat is required and the reference sets must match the source poses (R)")
    if not _lib.lib().jaf_warp_fuse_from_poses_supported(Cc, K, q.feat_layout, q.feat_dtype):
        raise RuntimeError("warp_fuse_from_poses_host serves channels-last bf16 features with C = 64, K <= 8")
    if rgb is not None:
        if out_rgb is None:
            out_rgb = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()
        q.out_rgb = _host(out_rgb, "out_rgb", torch.float32, (B, 3, S, S)).data_ptr()
    dptr = None
    if out_feat_device is not None:
        od = _check(out_feat_device, "out_feat_device", feat.dtype)
        if tuple(od.shape) != (B, S, S, Cc):
            raise RuntimeError("out_feat_device must be a CUDA tensor [B, S, S, C]")
        dptr, out_feat = od.data_ptr(), od
    else:
        if out_feat is None:
            out_feat = torch.empty((B, S, S, Cc), dtype=feat.dtype).pin_memory()
        q.out_feat = _host(out_feat, "out_feat", feat.dtype, (B, S, S, Cc)).data_ptr()
    for name, t, dtype, shape in (("logits", logits, torch.float32, (B, K, S, S)), ("fake", fake, torch.float32, (B, 3, S, S)),
                                  ("conf", conf, torch.float32, (B, 1, S, S))):
        if t is not None:
            setattr(q, name, _host(t, name, dtype, shape).data_ptr())
    if tgt_mask is not None:
        tgt_mask = _host(tgt_mask, "tgt_mask", torch.float32)
        if tgt_mask.dim() != 4 or tgt_mask.shape[0] != B or tgt_mask.shape[1] not in (1, 3) or tuple(tgt_mask.shape[2:]) != (S, S):
            raise RuntimeError("tgt_mask must be [B, 1 or 3, S, S]")
        q.tgt_mask, q.mask_c = tgt_mask.data_ptr(), tgt_mask.shape[1]
    pq = _lib.PoseFlowParams()
    pq.tgt_cam, pq.tgt_verts, pq.src_cam, pq.src_verts = tc.data_ptr(), tv.data_ptr(), sc.data_ptr(), sv.data_ptr()
    pq.faces_idx, pq.V, pq.F = faces_idx.data_ptr(), V, faces_idx.shape[-2]
    pq.eye_z, pq.near_, pq.far_ = eye_z, near, far
    if out_feat_device is not None:
        with _on(out_feat_device.device):
            _lib.check(_lib.lib().jaf_warp_fuse_from_poses_host(C.byref(q), C.byref(pq), int(frames_per_chunk), dptr),
                       "warp_fuse_from_poses_host")
    else:
        _lib.check(_lib.lib().jaf_warp_fuse_from_poses_host(C.byref(q), C.byref(pq), int(frames_per_chunk), None),
                   "warp_fuse_from_poses_host")
    return out_rgb, out_feat


# ----------------------------------------------------------------------------- a10
def grid_sample_border(src, grid, align_corners: bool = False):
    """F.grid_sample(src, grid, mode='bilinear', padding_mode='border') — float_estimate.warp_image
    (src/cal_flow.py:37-39) and the feature warps of src/crn_model.py:463-566."""
    src, grid = _check(src, "src"), _check(grid, "grid", torch.float32)
    if src.dim() != 4 or grid.dim() != 4 or grid.shape[0] != src.shape[0] or grid.shape[-1] != 2:
        raise RuntimeError("expected src [N,C,H,W] and grid [N,Ho,Wo,2]")
    if src.dtype == torch.float32:
        N, Cc, Hs, Ws = src.shape
        H, W = grid.shape[1:3]
        out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=src.device)
        with _on(src.device):
            _lib.check(_lib.lib().jaf_warp_image(_ptr(src), _ptr(grid), N, Cc, Hs, Ws, H, W, int(bool(align_corners)),
                                                 _ptr(out), _stream()), "warp_image")
        return out
    _, out = warp_fuse(grid[:, None], feat=src[:, None], align_corners=align_corners)
    return out


# ----------------------------------------------------------------------------- a11
def mask_blend(tsf_image, tgt_smpl_mask=None, fake_tgt=None, weight=None):
    """Propagation3DFlowNet.forward lines 91 / 98 (src/flow_net.py).  Returns (masked, pred|None)."""
    tsf = _check(tsf_image, "tsf_image", torch.float32)
    mask = _check(tgt_smpl_mask, "tgt_smpl_mask", torch.float32)
    fake, w = _check(fake_tgt, "fake_tgt", torch.float32), _check(weight, "weight", torch.float32)
    if tsf.dim() != 4:
        raise RuntimeError("tsf_image must be [B, C, H, W]")
    B, Cc, H, W = tsf.shape
    if mask is not None and (mask.dim() != 4 or mask.shape[0] != B or mask.shape[1] not in (1, Cc) or tuple(mask.shape[2:]) != (H, W)):
        raise RuntimeError("tgt_smpl_mask must be [B, 1 or C, H, W]")
    if (fake is None) != (w is None):
        raise RuntimeError("fake_tgt and weight go together")
    if w is not None and (tuple(w.shape) != (B, 1, H, W) or fake.shape != tsf.shape):
        raise RuntimeError("weight must be [B, 1, H, W] and fake_tgt must have tsf_image's shape")
    masked = torch.empty_like(tsf)
    pred = torch.empty_like(tsf) if w is not None else None
    mc = 1 if mask is None else mask.shape[1]
    with _on(tsf.device):
        _lib.check(_lib.lib().jaf_mask_blend(_ptr(fake), _ptr(tsf), _ptr(mask), mc, _ptr(w), B, Cc, H, W,
                                             _ptr(masked), _ptr(pred), _stream()), "mask_blend")
    return masked, pred


# ----------------------------------------------------------------------------- a12
def softmax_fuse(feat_cat, logits):
    """K-reduction of Downsampler_mask.forward (src/networks.py:1264-1286).
    feat_cat [B,K*C,H,W] f32, logits [B,K,H,W] (pre-softmax) -> [B,C,H,W]."""
    feat, logits = _check(feat_cat, "feat_cat", torch.float32), _check(logits, "logits", torch.float32)
    B, KC, H, W = feat.shape
    K = logits.shape[1]
    if KC % K != 0:
        raise RuntimeError("feat_cat channels must be K*C")
    out = torch.empty((B, KC // K, H, W), dtype=torch.float32, device=feat.device)
    with _on(feat.device):
        _lib.check(_lib.lib().jaf_softmax_fuse(_ptr(feat), _ptr(logits), B, K, KC // K, H, W, _ptr(out), _stream()),
                   "softmax_fuse")
    return out


# ----------------------------------------------------------------------------- a13
def convlstm_step(x, h, c, weight, bias=None):
    """ConvLSTMCell.forward (src/convLSTM.py:41-56), fp32 NCHW -> (h_next, c_next)."""
    x, h, c = _check(x, "input", torch.float32), _check(h, "h", torch.float32), _check(c, "c", torch.float32)
    weight, bias = _check(weight, "weight", torch.float32), _check(bias, "bias", torch.float32)
    B, Cin, H, W = x.shape
    Ch = h.shape[1]
    if weight.shape[0] != 4 * Ch or weight.shape[1] != Cin + Ch:
        raise RuntimeError("weight must be [4*Ch, Cin+Ch, kh, kw]")
    kh, kw = weight.shape[2:]
    h2, c2 = torch.empty_like(h), torch.empty_like(c)
    with _on(x.device):
        _lib.check(_lib.lib().jaf_convlstm_step_f32(_ptr(x), _ptr(h), _ptr(c), _ptr(weight), _ptr(bias), B, Cin, Ch,
                                                    H, W, kh, kw, _ptr(h2), _ptr(c2), _stream()), "convlstm_step")
    return h2, c2


def convlstm_pack_weight(weight, Cin: int, Ch: int):
    """Repack a [4Ch, Cin+Ch, 3, 3] f32 weight for the tensor-core cell (done once per model)."""
    weight = _check(weight, "weight", torch.float32)
    if tuple(weight.shape) != (4 * Ch, Cin + Ch, 3, 3):
        raise RuntimeError("weight must be [4*Ch, Cin+Ch, 3, 3]")
    n = _lib.lib().jaf_convlstm_wpack_bytes(Cin, Ch)
    wpack = torch.empty(n, dtype=torch.uint8, device=weight.device)
    with _on(weight.device):
        _lib.check(_lib.lib().jaf_convlstm_pack_weight(_ptr(weight), Cin, Ch, _ptr(wpack), _stream()), "pack_weight")
    return wpack


def convlstm_step_tc(x, h, c, wpack, bias, Cin: int, Ch: int):
    """Tensor-core ConvLSTM step (tcgen05): x [B,H,W,Cin] bf16, h [B,H,W,Ch] bf16, c [B,H,W,Ch] f32
    (all channels-last dense) -> (h_next bf16, c_next f32)."""
    x, h = _check(x, "x", torch.bfloat16), _check(h, "h", torch.bfloat16)
    c, bias = _check(c, "c", torch.float32), _check(bias, "bias", torch.float32)
    if x.dim() != 4 or x.shape[-1] != Cin:
        raise RuntimeError("x must be [B, H, W, Cin] (channels-last bf16)")
    B, H, W, _ = x.shape
    if tuple(h.shape) != (B, H, W, Ch) or tuple(c.shape) != (B, H, W, Ch):
        raise RuntimeError("h and c must be [B, H, W, Ch]")
    if bias is not None and tuple(bias.shape) != (4 * Ch,):
        raise RuntimeError("bias must be [4*Ch]")
    if wpack.numel() != _lib.lib().jaf_convlstm_wpack_bytes(Cin, Ch):
        raise RuntimeError("wpack was packed for other channel counts")
    h2, c2 = torch.empty_like(h), torch.empty_like(c)
    with _on(x.device):
        _lib.check(_lib.lib().jaf_convlstm_step_tc(_ptr(x), _ptr(h), _ptr(c), _ptr(wpack), _ptr(bias), B, Cin, Ch, H,
                                                   W, _ptr(h2), _ptr(c2), _stream()), "convlstm_step_tc")
    return h2, c2


def convlstm_gpack_weight(weight, Cin: int, Ch: int):
    """Repack G cell weights [G,4Ch,Cin+Ch,3,3] f32 (one-off) for `convlstm_step_grouped`."""
    weight = _check(weight, "weight", torch.float32)
    if weight.dim() != 5 or tuple(weight.shape[1:]) != (4 * Ch, Cin + Ch, 3, 3):
        raise RuntimeError("expected weight [G,4Ch,Cin+Ch,3,3]")
    G = weight.shape[0]
    nbytes = _lib.lib().jaf_convlstm_gpack_bytes(G, Cin, Ch)
    if nbytes == 0:
        raise RuntimeError("grouped ConvLSTM: unsupported channel counts (Ch % 4 == 0, Ch <= 128)")
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    with _on(weight.device):
        _lib.check(_lib.lib().jaf_convlstm_gpack_weight(_ptr(weight), G, Cin, Ch, _ptr(wpack), _stream()),
                   "convlstm_gpack_weight")
    return wpack


def convlstm_grouped_supported(G: int, B: int, Cin: int, Ch: int, H: int, W: int) -> bool:
    """True when `convlstm_step_grouped` can run this cell (channel counts AND the shared-memory plan fit)."""
    return bool(_lib.lib().jaf_convlstm_grouped_supported(G, B, Cin, Ch, H, W))


def convlstm_step_grouped(x, h, c, wpack, bias, Cin: int, Ch: int):
    """G independent ConvLSTM cells, one launch, tensor cores with split-bf16 (fp32-grade) arithmetic.
    x [G,B,Cin,H,W], h, c [G,B,Ch,H,W] f32 (reference NCHW layout), bias [G,4Ch] or None -> (h_next, c_next)."""
    x, h, c = _check(x, "x", torch.float32), _check(h, "h", torch.float32), _check(c, "c", torch.float32)
    if bias is not None:
        bias = _check(bias, "bias", torch.float32)
    if x.dim() != 5 or h.shape != c.shape or x.shape[1] != h.shape[1] or x.shape[2] != Cin or h.shape[2] != Ch:
        raise RuntimeError("expected x [G,B,Cin,H,W] and h, c [G,B,Ch,H,W]")
    G, B, _, H, W = x.shape
    h2, c2 = torch.empty_like(h), torch.empty_like(c)
    with _on(x.device):
        _lib.check(_lib.lib().jaf_convlstm_step_grouped(_ptr(x), _ptr(h), _ptr(c), _ptr(wpack), _ptr(bias), G, B, Cin,
                                                        Ch, H, W, _ptr(h2), _ptr(c2), _stream()),
                   "convlstm_step_grouped")
    return h2, c2


def convlstm_sequence_grouped(x_seq, h0, c0, wpack, bias, Cin: int, Ch: int):
    """T recurrent steps of G cells in one C call (src/convLSTM.py:131-134): x_seq [G,B,T,Cin,H,W], h0, c0 [G,B,Ch,H,W]
    -> (h_seq [G,B,T,Ch,H,W], c_last [G,B,Ch,H,W]).  Bit-identical to T calls of `convlstm_step_grouped`."""
    x, h0, c0 = _check(x_seq, "x_seq", torch.float32), _check(h0, "h0", torch.float32), _check(c0, "c0", torch.float32)
    if bias is not None:
        bias = _check(bias, "bias", torch.float32)
    if x.dim() != 6 or x.shape[3] != Cin or h0.shape != c0.shape or h0.dim() != 5 or h0.shape[2] != Ch or \
            x.shape[:2] != h0.shape[:2] or x.shape[4:] != h0.shape[3:]:
        raise RuntimeError("expected x_seq [G,B,T,Cin,H,W] and h0, c0 [G,B,Ch,H,W]")
    G, B, T, _, H, W = x.shape
    h_seq = torch.empty((G, B, T, Ch, H, W), dtype=torch.float32, device=x.device)
    c_last = torch.empty_like(c0)
    c_tmp = torch.empty_like(c0) if T > 1 else None
    with _on(x.device):
        _lib.check(_lib.lib().jaf_convlstm_sequence_grouped(_ptr(x), _ptr(h0), _ptr(c0), _ptr(wpack), _ptr(bias), G, B, T, Cin,
                                                            Ch, H, W, _ptr(h_seq), _ptr(c_last), _ptr(c_tmp), _stream()),
                   "convlstm_sequence_grouped")
    return h_seq, c_last


# ----------------------------------------------------------------------------- §8f rank 1
def texture_warp(tex_parts, iuv, align_corners: bool = False):
    """IUV texture lookup (test/conv_pro_test.py:41-74).  tex_parts [P,3,Ht,Wt] f32, iuv [B,H,W,3] uint8
    -> [B,3,H,W] f32."""
    tex, iuv = _check(tex_parts, "tex_parts", torch.float32), _check(iuv, "IUV", torch.uint8)
    if tex.dim() != 4 or tex.shape[1] != 3 or iuv.dim() != 4 or iuv.shape[-1] != 3:
        raise RuntimeError("expected tex_parts [P,3,Ht,Wt] and IUV [B,H,W,3]")
    P, _, Ht, Wt = tex.shape
    B, H, W, _ = iuv.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=iuv.device)
    with _on(iuv.device):
        _lib.check(_lib.lib().jaf_texture_warp(_ptr(tex), P, Ht, Wt, _ptr(iuv), B, H, W, int(bool(align_corners)),
                                               _ptr(out), _stream()), "texture_warp")
    return out


# ----------------------------------------------------------------------------- a8 for K references
def cal_flow_multi(src_cam, src_vertices, tgt_cam, tgt_vertices, faces_idx, image_size: int, eye_z: float = EYE_Z,
                   near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR, return_maps: bool = True,
                   return_wim: bool = True):
    """Transfer flows from K source poses into one target pose per frame: the target is rasterised once,
    composed K times.  src_cam [B,K,3], src_vertices [B,K,V,3], tgt_cam [B,3], tgt_vertices [B,V,3]
    -> T [B,K,S,S,2] (+ fim [B,S,S], wim [B,S,S,3]; wim is None with return_wim=False: it then never touches HBM)."""
    sc, sv = _check(src_cam, "src_cam", torch.float32), _check(src_vertices, "src_vertices", torch.float32)
    tc, tv = _check(tgt_cam, "tgt_cam", torch.float32), _check(tgt_vertices, "tgt_vertices", torch.float32)
    faces_idx = _check(faces_idx, "faces", torch.int32)
    if sv.dim() != 4 or sc.dim() != 3 or sv.shape[:2] != sc.shape[:2] or sv.shape[0] != tv.shape[0]:
        raise RuntimeError("expected src_cam [B,K,3], src_vertices [B,K,V,3], tgt_cam [B,3], tgt_vertices [B,V,3]")
    B, K, V, _ = sv.shape
    F = faces_idx.shape[-2]
    dev = tv.device
    T = torch.empty((B, K, image_size, image_size, 2), dtype=torch.float32, device=dev)
    fim = torch.empty((B, image_size, image_size), dtype=torch.int32, device=dev) if return_maps else None
    wim = torch.empty((B, image_size, image_size, 3), dtype=torch.float32, device=dev) if (return_maps and return_wim) else None
    with _on(dev):
        ws = _workspace(dev, _lib.lib().jaf_raster_workspace_bytes(B, image_size))
        _lib.check(_lib.lib().jaf_cal_flow_multi(_ptr(sc), _ptr(sv), _ptr(tc), _ptr(tv), _ptr(faces_idx), B, K, V, F,
                                                 image_size, eye_z, near, far, _ptr(T), _ptr(fim), _ptr(wim), _ptr(ws),
                                                 _stream()), "cal_flow_multi")
    return (T, fim, wim) if return_maps else T


# ----------------------------------------------------------------------------- row F: per-reference visibility
def face_visibility(fim_src, fim_tgt, num_faces: int):
    """The reference's visibility rule (SMPLRenderer.get_vis_f2pts, src/nmr.py:507-546) as data for `warp_fuse`:
    fim_src [B,K,H,W] int32 (face-index maps of the K reference poses), fim_tgt [B,H,W] int32 or None
    -> (seen [B,K,F] uint8, vis [B,K,H,W] f32 or None) with vis = 1 where the face shown by the target pixel is
    visible in reference k."""
    fim_src = _check(fim_src, "fim_src", torch.int32)
    if fim_src.dim() != 4:
        raise RuntimeError("expected fim_src [B,K,H,W]")
    B, K, H, W = fim_src.shape
    vis = None
    if fim_tgt is not None:
        fim_tgt = _check(fim_tgt, "fim_tgt", torch.int32)
        if tuple(fim_tgt.shape) != (B, H, W):
            raise RuntimeError("expected fim_tgt [B,H,W]")
        vis = torch.empty((B, K, H, W), dtype=torch.float32, device=fim_src.device)
    seen = torch.empty((B, K, num_faces), dtype=torch.uint8, device=fim_src.device)
    with _on(fim_src.device):
        _lib.check(_lib.lib().jaf_face_visibility(_ptr(fim_src), _ptr(fim_tgt), B, K, H * W, num_faces, _ptr(seen),
                                                  _ptr(vis), _stream()), "face_visibility")
    return seen, vis


# ----------------------------------------------------------------------------- §8f rank 3
def flow_warp_pair(feat_fwd, feat_bwd, base_grid, flow, align_corners: bool = False):
    """One pyramid level of SpatioTempoCRN.forward (src/crn_model.py:457-566): nearest-downsample `flow` [B,2,H,W]
    to the feature resolution, warp `feat_fwd` with (grid + flow) and `feat_bwd` with (grid - flow), border padding.
    feats [B,C,h,w] f32 (either may be None), base_grid [B,2,h,w] -> (out_fwd, out_bwd)."""
    base_grid, flow = _check(base_grid, "grid", torch.float32), _check(flow, "flow", torch.float32)
    ref = feat_fwd if feat_fwd is not None else feat_bwd
    if ref is None:
        raise RuntimeError("nothing to warp")
    B, C, h, w = ref.shape
    if tuple(base_grid.shape) != (B, 2, h, w) or flow.dim() != 4 or flow.shape[:2] != (B, 2):
        raise RuntimeError("expected base_grid [B,2,h,w] and flow [B,2,H,W]")
    outs = []
    for f in (feat_fwd, feat_bwd):
        if f is not None:
            f = _check(f, "feat", torch.float32)
            if tuple(f.shape) != (B, C, h, w):
                raise RuntimeError("feature maps must share their shape")
        outs.append(None if f is None else torch.empty_like(f))
    H, W = flow.shape[2:]
    with _on(ref.device):
        _lib.check(_lib.lib().jaf_flow_warp_pair(_ptr(feat_fwd), _ptr(feat_bwd), _ptr(base_grid), _ptr(flow), B, C, h, w,
                                                 H, W, int(bool(align_corners)), _ptr(outs[0]), _ptr(outs[1]),
                                                 _stream()), "flow_warp_pair")
    return outs[0], outs[1]


# ----------------------------------------------------------------------------- §8f rank 2
def _atlas_geom(AH, AW, rows, cols):
    if AH % rows or AW % cols:
        raise RuntimeError("atlas size must be a multiple of the part grid")
    return AH // rows, AW // cols


def texture_parts_gather(atlas, ref_index, rows: int = 4, cols: int = 6):
    """atlas [B,Kmax,C,AH,AW] f32, ref_index [K] int32 -> [rows*cols, K, B, C, ph, pw] (test/conv_pro_test.py:209-217)."""
    atlas, ref_index = _check(atlas, "atlas", torch.float32), _check(ref_index, "ref_index", torch.int32)
    B, Kmax, C, AH, AW = atlas.shape
    ph, pw = _atlas_geom(AH, AW, rows, cols)
    K = ref_index.numel()
    out = torch.empty((rows * cols, K, B, C, ph, pw), dtype=torch.float32, device=atlas.device)
    with _on(atlas.device):
        _lib.check(_lib.lib().jaf_texture_parts_gather(_ptr(atlas), _ptr(ref_index), B, Kmax, K, C, rows, cols, ph, pw,
                                                       _ptr(out), _stream()), "texture_parts_gather")
    return out


def texture_parts_common_mask_(parts, mask, ref_index, rows: int = 4, cols: int = 6):
    """In place: parts [rows*cols,B,C,ph,pw] *= float(OR_z uint8(mask[:, ref_index[z]])) (conv_pro_test.py:221-236)."""
    parts, mask = _check(parts, "parts", torch.float32), _check(mask, "mask", torch.float32)
    ref_index = _check(ref_index, "ref_index", torch.int32)
    P, B, C, ph, pw = parts.shape
    if P != rows * cols or tuple(mask.shape) != (B, mask.shape[1], rows * ph, cols * pw):
        raise RuntimeError("expected parts [rows*cols,B,C,ph,pw] and mask [B,Kmax,rows*ph,cols*pw]")
    with _on(parts.device):
        _lib.check(_lib.lib().jaf_texture_parts_common_mask(_ptr(parts), _ptr(mask), _ptr(ref_index), B, mask.shape[1],
                                                            ref_index.numel(), C, rows, cols, ph, pw, _stream()),
                   "texture_parts_common_mask")
    return parts


def texture_parts_scatter(parts, rows: int = 4, cols: int = 6):
    """parts [rows*cols,B,C,ph,pw] -> atlas [B,C,rows*ph,cols*pw] (src/networks.py:1685-1691)."""
    parts = _check(parts, "parts", torch.float32)
    P, B, C, ph, pw = parts.shape
    if P != rows * cols:
        raise RuntimeError("expected rows*cols parts")
    atlas = torch.empty((B, C, rows * ph, cols * pw), dtype=torch.float32, device=parts.device)
    with _on(parts.device):
        _lib.check(_lib.lib().jaf_texture_parts_scatter(_ptr(parts), B, C, rows, cols, ph, pw, _ptr(atlas), _stream()),
                   "texture_parts_scatter")
    return atlas


# ----------------------------------------------------------------------------- §8f rank 4
def transfer_texture(tex, iuv, im=None, rows: int = 4, cols: int = 6):
    """TransferTexture (src/utils.py:369-394) for a batch: tex [rows*ps, cols*ps, 3] uint8 (or [B,...] one atlas per
    frame), iuv [B,H,W,3] uint8, im [B,H,W,3] uint8 or None -> [B,H,W,3] uint8."""
    tex, iuv = _check(tex, "TextureIm", torch.uint8), _check(iuv, "IUV", torch.uint8)
    if im is not None:
        im = _check(im, "im", torch.uint8)
    batched = tex.dim() == 4
    AH, AW = tex.shape[-3], tex.shape[-2]
    if tex.shape[-1] != 3 or AH % rows or AW % cols or AH // rows != AW // cols or iuv.dim() != 4 or iuv.shape[-1] != 3:
        raise RuntimeError("expected tex [rows*ps, cols*ps, 3] and IUV [B,H,W,3]")
    B, H, W, _ = iuv.shape
    if (batched and tex.shape[0] != B) or (im is not None and im.shape != iuv.shape):
        raise RuntimeError("batch / image shapes do not match")
    out = torch.empty_like(iuv)
    with _on(iuv.device):
        _lib.check(_lib.lib().jaf_transfer_texture(_ptr(tex), int(batched), rows, cols, AH // rows, _ptr(iuv), _ptr(im), B, H,
                                                   W, _ptr(out), _stream()), "transfer_texture")
    return out


def get_texture(im, iuv, tex_size: int = 32, final_size: int = 200):
    """get_texture (src/utils.py:232-255) for a batch: im, iuv [B,H,W,3] uint8 -> parts [B,24,final,final,3] float64."""
    im, iuv = _check(im, "im", torch.uint8), _check(iuv, "IUV", torch.uint8)
    if iuv.dim() != 4 or iuv.shape[-1] != 3 or im.shape != iuv.shape:
        raise RuntimeError("expected im and IUV [B,H,W,3]")
    B, H, W, _ = iuv.shape
    parts = torch.empty((B, 24, final_size, final_size, 3), dtype=torch.float64, device=iuv.device)
    with _on(iuv.device):
        ws = torch.empty(_lib.lib().jaf_get_texture_workspace_bytes(B, tex_size), dtype=torch.uint8, device=iuv.device)
        _lib.check(_lib.lib().jaf_get_texture(_ptr(im), _ptr(iuv), B, H, W, int(tex_size), int(final_size), _ptr(parts),
                                              _ptr(ws), _stream()), "get_texture")
    return parts


def iuv_part_stats(iuv):
    """iuv [B,H,W,3] uint8 -> (counts [B,32] int32, sumx [B,32] int64): pixels per part id and the sum of their x."""
    iuv = _check(iuv, "IUV", torch.uint8)
    B, H, W, _ = iuv.shape
    counts = torch.empty((B, 32), dtype=torch.int32, device=iuv.device)
    sumx = torch.empty((B, 32), dtype=torch.int64, device=iuv.device)
    with _on(iuv.device):
        _lib.check(_lib.lib().jaf_iuv_part_stats(_ptr(iuv), B, H, W, _ptr(counts), _ptr(sumx), _stream()), "iuv_part_stats")
    return counts, sumx
