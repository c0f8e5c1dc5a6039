"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/jafpro_b200.h declares, argument validation works without a GPU, and the host mirror of the
reference interface imports.  No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

from jafpro_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "jafpro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jaf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 18
    handle = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/jafpro_b200.h but not exported"
    # and the Python binding table covers the header exactly
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_reporting_without_gpu():
    lib = _lib.lib()
    assert lib.jaf_version() >= 100
    # NULL params is an argument error, reported before any CUDA call
    assert lib.jaf_warp_fuse(None) == -1
    assert "null params" in _lib.last_error()
    with pytest.raises(RuntimeError, match="null params"):
        _lib.check(lib.jaf_warp_fuse(None), "warp_fuse")
    assert lib.jaf_raster_workspace_bytes(2, 256) >= 2 * 256 * 256 * 8  # z-buffer keys + the deferred-box queue
    assert lib.jaf_convlstm_wpack_bytes(256, 256) == 9 * 1024 * 512 * 2


def test_warp_fuse_params_struct_matches_header_layout():
    # 12 int32 + 14 pointers, no padding surprises (the oracle mirrors the same struct)
    assert C.sizeof(_lib.WarpFuseParams) == 12 * 4 + 14 * 8
    import oracle
    assert C.sizeof(oracle.WarpFuseParams) == C.sizeof(_lib.WarpFuseParams)
    assert [f[0] for f in oracle.WarpFuseParams._fields_] == [f[0] for f in _lib.WarpFuseParams._fields_]


def test_ops_reject_cpu_tensors_like_the_reference_extension():
    """NR/cuda/rasterize_cuda.cpp:66-68: CHECK_CUDA -> RuntimeError.  No silent CPU fallback."""
    from jafpro_b200 import ops
    faces = torch.zeros(1, 4, 3, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.raster_fim_wim(faces, 32)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.grid_sample_border(torch.zeros(1, 3, 8, 8), torch.zeros(1, 8, 8, 2))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.convlstm_step(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4),
                          torch.zeros(8, 4, 3, 3))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under jafpro_b200/ may reference it."""
    pkg = os.path.join(ROOT, "jafpro_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "jaf_oracle" not in src and "orc_" not in src, f


def test_host_mirror_keeps_reference_names():
    from jafpro_b200 import cal_flow, convLSTM, flow_net, fusion, neural_renderer, nmr
    for mod, names in ((neural_renderer, ["look_at", "vertices_to_faces", "rasterize_face_index_map_and_weight_map"]),
                       (nmr, ["SMPLRenderer", "orthographic_proj_withz_idrot"]),
                       (cal_flow, ["float_estimate"]), (convLSTM, ["ConvLSTMCell", "ConvLSTM"]),
                       (flow_net, ["Propagation3DFlowNet"]), (fusion, ["warp_fuse", "softmax_fuse"])):
        for n in names:
            assert hasattr(mod, n)
    r = nmr.SMPLRenderer(image_size=64)
    assert tuple(r.faces.shape) == (13776, 3) and r.faces.dtype == torch.int32
    for m in ("render_fim_wim", "cal_bc_transform", "render_fim"):
        assert callable(getattr(r, m))
    cell = convLSTM.ConvLSTMCell((8, 8), 3, 4, (3, 3), True)
    assert sorted(cell.state_dict()) == ["conv.bias", "conv.weight"]
    assert tuple(cell.conv.weight.shape) == (16, 7, 3, 3)


def test_host_look_at_and_gather_match_reference_fixtures(golden_dir):
    """The torch-level helpers (not kernels) against the reference-generated fixtures."""
    import numpy as np
    from jafpro_b200 import neural_renderer as nr
    d = np.load(os.path.join(golden_dir, "look_at.npz"))
    for e, out in zip(d["known_eyes"], d["known_out"]):
        got = nr.look_at(torch.from_numpy(d["known_in"]), e).numpy().squeeze()
        assert np.allclose(got, out, atol=1e-6)
    eye = [0, 0, -(1. / np.tan(np.radians(30)) + 1)]
    got = nr.look_at(torch.from_numpy(d["verts"]), eye).numpy()
    assert np.array_equal(got, d["smpl_eye_out"])
