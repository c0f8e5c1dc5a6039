"""ctypes binding of ``libjafpro_b200.so`` (the C ABI in include/jafpro_b200.h).

There is no fallback: if the library is missing the import fails loudly, and every
non-zero status becomes a ``RuntimeError`` carrying ``jaf_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JAFPRO_B200_LIB") or os.path.join(_HERE, "libjafpro_b200.so")  # override: A/B builds
CSRC = os.path.join(_HERE, "csrc")

_vp, _i, _f, _sz, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_uint64


class WarpFuseParams(C.Structure):
    """``JafWarpFuseParams`` (include/jafpro_b200.h)."""
    _fields_ = [
        ("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Hs", C.c_int32), ("Ws", C.c_int32), ("C", C.c_int32),
        ("align_corners", C.c_int32), ("feat_layout", C.c_int32), ("feat_dtype", C.c_int32),
        ("mask_c", C.c_int32), ("reserved", C.c_int32),
        ("rgb", _vp), ("feat", _vp), ("ref_index", _vp), ("grid", _vp), ("logits", _vp),
        ("vis", _vp), ("fim", _vp), ("tgt_mask", _vp), ("fake", _vp), ("conf", _vp),
        ("out_rgb", _vp), ("out_feat", _vp), ("warped_rgb", _vp), ("stream", _vp),
    ]


class PoseFlowParams(C.Structure):
    """``JafPoseFlowParams`` (include/jafpro_b200.h)."""
    _fields_ = [
        ("tgt_cam", _vp), ("tgt_verts", _vp), ("src_cam", _vp), ("src_verts", _vp), ("faces_idx", _vp),
        ("V", C.c_int32), ("F", C.c_int32), ("eye_z", _f), ("near_", _f), ("far_", _f), ("flags", C.c_int32),
        ("T", _vp), ("fim", _vp), ("workspace", _vp),
    ]


# name -> (restype, argtypes); kept in the order of include/jafpro_b200.h
SIGNATURES = {
    "jaf_version": (_i, []),
    "jaf_last_error": (C.c_char_p, []),
    "jaf_launch_count": (_u64, []),
    "jaf_last_kernel": (C.c_char_p, []),
    "jaf_tuning_info": (_i, [C.c_char_p, _i]),
    "jaf_project_gather": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "jaf_raster_workspace_bytes": (_sz, [_i, _i]),
    "jaf_raster_fim_wim": (_i, [_vp, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "jaf_forward_face_index_map": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _vp]),
    "jaf_render_fim_wim": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "jaf_flow_compose": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "jaf_cal_flow": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "jaf_cal_flow_multi": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "jaf_warp_fuse": (_i, [C.POINTER(WarpFuseParams)]),
    "jaf_warp_fuse_from_poses_supported": (_i, [_i, _i, _i, _i]),
    "jaf_warp_fuse_from_poses": (_i, [C.POINTER(WarpFuseParams), C.POINTER(PoseFlowParams)]),
    "jaf_warp_fuse_host": (_i, [C.POINTER(WarpFuseParams), _i]),
    "jaf_warp_fuse_from_poses_host": (_i, [C.POINTER(WarpFuseParams), C.POINTER(PoseFlowParams), _i, _vp]),
    "jaf_warp_image": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "jaf_mask_blend": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "jaf_softmax_fuse": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "jaf_convlstm_step_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "jaf_face_visibility": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "jaf_convlstm_wpack_bytes": (_sz, [_i, _i]),
    "jaf_convlstm_pack_weight": (_i, [_vp, _i, _i, _vp, _vp]),
    "jaf_convlstm_step_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "jaf_convlstm_gpack_bytes": (_sz, [_i, _i, _i]),
    "jaf_convlstm_grouped_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "jaf_convlstm_gpack_weight": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "jaf_convlstm_step_grouped": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "jaf_convlstm_sequence_grouped": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "jaf_flow_warp_pair": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "jaf_texture_warp": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "jaf_texture_parts_gather": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "jaf_texture_parts_common_mask": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "jaf_texture_parts_scatter": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "jaf_transfer_texture": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "jaf_iuv_part_stats": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "jaf_get_texture_workspace_bytes": (_sz, [_i, _i]),
    "jaf_get_texture": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``libjafpro_b200.so`` (nvcc cross-compiles
    without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libjafpro_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C jafpro_b200/csrc`). "
                "jafpro_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().jaf_last_error().decode("utf-8", "replace")


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise RuntimeError(f"jafpro_b200 {what} failed ({status}): {last_error()}")


def launch_count() -> int:
    return int(lib().jaf_launch_count())


def last_kernel() -> str:
    """Kernel variant the calling thread launched last through jaf_warp_fuse (measurement aid)."""
    return lib().jaf_last_kernel().decode("utf-8", "replace")


def tuning_info() -> dict:
    """Effective JAF_* tuning knobs of the loaded library (environment, read once per process)."""
    buf = C.create_string_buffer(1024)
    lib().jaf_tuning_info(buf, len(buf))
    out = {}
    for tok in buf.value.decode().split():
        k, _, v = tok.partition("=")
        out[k] = int(v) if v.lstrip("-").isdigit() else v
    return out
