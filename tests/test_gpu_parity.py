"""GPU parity tests (run on the B200 box with ``-m gpu``): every call goes through the C ABI
(jafpro_b200 -> libjafpro_b200.so) and is compared with the CPU oracle on the same seeded inputs,
with the committed golden fixtures, with the reference's own CUDA rasteriser (oracle/_ref, built
from the unmodified reference source), and — for the third-party grid_sample arithmetic — with the
installed torch.

Bars: bit-exact for face-index maps, barycentric weights, flows and every fp32 gather that has no
transcendental in it; <= 1e-4 max-abs (stated per test, usually far tighter) where expf or a
different-but-equivalent summation order is involved; <= 1 bf16 ulp for bf16 feature outputs.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from jafpro_b200 import _lib, ops, synth
from jafpro_b200.nmr import SMPLRenderer, load_smpl_template

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


@pytest.fixture(autouse=True)
def _inference_only():
    """The path is forward-only and the drop-ins refuse tensors that require grad while autograd is on (the
    reference runs it under torch.no_grad(), test/conv_pro_test.py:190)."""
    with torch.no_grad():
        yield


def _bits(t):
    return t.detach().cpu().contiguous().view(torch.int32).numpy()


def _np(t):
    return t.detach().cpu().numpy()


def _bf16_bits(t):
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def test_native_library_is_loaded_and_counts_launches():
    assert torch.cuda.is_available()
    n0 = _lib.launch_count()
    ops.grid_sample_border(torch.zeros(1, 1, 4, 4, device=DEV), torch.zeros(1, 4, 4, 2, device=DEV))
    assert _lib.launch_count() == n0 + 1
    maps = open("/proc/self/maps").read()
    assert "libjafpro_b200.so" in maps


# ------------------------------------------------------------------ rasteriser (a4-a6)
def test_teapot_golden_silhouette(golden_dir):
    d = _load(golden_dir, "teapot_faces.npz")
    shape = tuple(d["shape"])
    ref = np.unpackbits(d["silhouette_bits"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
    faces = _cu(d["faces"][None])
    fim, wim, depth = ops.raster_fim_wim(faces, 256, return_depth=True)
    assert np.array_equal(_np(fim[0]) != -1, ref)
    # NR tests/utils.py:11-27: the sample sits at index 2 of a batch of all-zero meshes
    batch = torch.zeros((4,) + tuple(faces.shape[1:]), device=DEV)
    batch[2] = faces[0]
    fim4, wim4 = ops.raster_fim_wim(batch, 256)
    assert np.array_equal(_np(fim4[2]) != -1, ref)
    for i in (0, 1, 3):
        assert int((fim4[i] != -1).sum()) == 0 and float(wim4[i].abs().max()) == 0.0
    # against the oracle: every value bit for bit (medium / large faces, warp-cooperative path)
    ofim, owim, odepth = oracle.raster_fim_wim(d["faces"][None], 256, return_depth=True)
    assert np.array_equal(_np(fim), ofim)
    assert np.array_equal(_bits(wim), owim.view(np.int32))
    assert np.array_equal(_bits(depth), odepth.view(np.int32))


def test_raster_matches_reference_cuda_kernels_bit_exact():
    """fim / wim / depth against the reference's own kernels (rasterize_cuda_kernel.cu:24-169) compiled
    unmodified for sm_100a — the parity gate of the north star (face-index maps bit-exact)."""
    try:
        ref = oracle.RefRaster()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libjaf_ref_raster.so not built (needs /root/reference at build time)")
    rend = SMPLRenderer(image_size=256).to(DEV)
    cam, verts = synth.smpl_poses(4, seed=11, device=DEV)
    faces, fim, wim = rend.render_fim_wim(cam, verts)
    rfim, rwim, rdepth = ref(faces, 256)
    assert 0.05 < float((rfim != -1).float().mean()) < 0.3
    assert torch.equal(fim, rfim)
    assert np.array_equal(_bits(wim), _bits(rwim))
    _, _, depth = ops.raster_fim_wim(faces, 256, return_depth=True)
    assert np.array_equal(_bits(depth), _bits(rdepth))
    # other sizes and a random triangle soup with big, overlapping, partly off-screen faces
    g = torch.Generator().manual_seed(5)
    soup = (torch.rand((2, 300, 3, 3), generator=g) * 2.6 - 1.3)
    soup[..., 2] = torch.rand((2, 300, 3), generator=g) * 3 + 0.5
    soup = soup.to(DEV)
    for size in (64, 200, 512):
        a = ops.raster_fim_wim(soup, size, return_depth=True)
        b = ref(soup, size)
        assert torch.equal(a[0], b[0]), size
        assert np.array_equal(_bits(a[1]), _bits(b[1])), size
        assert np.array_equal(_bits(a[2]), _bits(b[2])), size
    # SMPL mesh at 512 (BASELINE config 5)
    f512 = ops.raster_fim_wim(faces[:2].contiguous(), 512)
    r512 = ref(faces[:2].contiguous(), 512)
    assert torch.equal(f512[0], r512[0]) and np.array_equal(_bits(f512[1]), _bits(r512[1]))


def test_raster_matches_oracle_on_smpl_and_edge_cases():
    _, faces_idx = load_smpl_template()
    cam, verts = synth.smpl_poses(2, seed=2)
    ofaces, ofim, owim = oracle.render_fim_wim(cam.numpy(), verts.numpy(), faces_idx, 128)
    faces, fim, wim = ops.render_fim_wim(cam.to(DEV), verts.to(DEV), _cu(faces_idx), 128)
    assert np.array_equal(_bits(faces), ofaces.view(np.int32))
    assert np.array_equal(_np(fim), ofim)
    assert np.array_equal(_bits(wim), owim.view(np.int32))
    # two-step path (explicit faces tensor) == fused path
    fim2, wim2 = ops.raster_fim_wim(faces, 128)
    assert torch.equal(fim, fim2) and torch.equal(wim, wim2)
    # depth ties resolve to the lowest face index; exact-edge pixels belong to both neighbours
    quad = np.array([[[-1, -1, 2], [1, -1, 2], [1, 1, 2]], [[-1, -1, 2], [1, 1, 2], [-1, 1, 2]],
                     [[-1, -1, 2], [1, -1, 2], [1, 1, 2]]], np.float32)[None]
    for q in (quad, quad[:, ::-1].copy()):
        of, ow = oracle.raster_fim_wim(q, 16)
        gf, gw = ops.raster_fim_wim(_cu(q), 16)
        assert np.array_equal(_np(gf), of) and np.array_equal(_bits(gw), ow.view(np.int32))
    # degenerate inputs: zero-area, collinear, NaN / inf vertices, behind near / beyond far, empty
    rng = np.random.default_rng(0)
    deg = rng.uniform(-1, 1, (1, 40, 3, 3)).astype(np.float32)
    deg[..., 2] = rng.uniform(0.5, 3, (1, 40, 3))
    deg[0, 0] = 0
    deg[0, 1, 1] = deg[0, 1, 0]
    deg[0, 2, 2] = 0.5 * (deg[0, 2, 0] + deg[0, 2, 1])
    deg[0, 3, 0, 0] = np.nan
    deg[0, 4, 1, 1] = np.inf
    deg[0, 5, :, 2] = 0.05
    deg[0, 6, :, 2] = 500.0
    deg[0, 7] = [[-0.5, 0.03125, 1], [0.5, 0.03125, 1], [0.0, 0.03125, 1]]  # collinear through pixel centres
    of, ow = oracle.raster_fim_wim(deg, 32)
    gf, gw = ops.raster_fim_wim(_cu(deg), 32)
    assert np.array_equal(_np(gf), of)
    assert np.array_equal(_bits(gw), ow.view(np.int32))
    ef, ew = ops.raster_fim_wim(torch.zeros((1, 0, 3, 3), device=DEV), 8)
    assert int((ef != -1).sum()) == 0 and float(ew.abs().max()) == 0


def test_project_gather_matches_reference_fixture(golden_dir):
    d = _load(golden_dir, "render_faces.npz")
    _, faces_idx = load_smpl_template()
    out = ops.project_gather(_cu(d["cam"]), _cu(d["verts"]), _cu(faces_idx))
    assert np.array_equal(_np(out)[:, d["face_subset"]], d["faces_xyz_subset"])


# ------------------------------------------------------------------ flow (a8, a9)
def test_flow_compose_matches_reference_fixture_bit_exact(golden_dir):
    d = _load(golden_dir, "bc_transform.npz")
    T = ops.flow_compose(_cu(d["src"]), _cu(d["fim"]), _cu(d["wim"]))
    assert np.array_equal(_bits(T), d["T"].view(np.int32))


def test_cal_flow_fused_equals_stepwise_and_oracle():
    from jafpro_b200.cal_flow import float_estimate
    _, faces_idx = load_smpl_template()
    cam, verts = synth.smpl_poses(4, seed=4)
    sc, sv, tc, tv = cam[:2], verts[:2], cam[2:], verts[2:]
    oT, ofim, owim = oracle.cal_flow(sc.numpy(), sv.numpy(), tc.numpy(), tv.numpy(), faces_idx, 128)
    fe = float_estimate(image_size=128).to(DEV)
    args = (sc.to(DEV), None, sv.to(DEV), None, tc.to(DEV), None, tv.to(DEV), None)
    T = fe.cal_flow(*args)
    assert np.array_equal(_bits(T), oT.view(np.int32))
    fe.fused = False
    T2 = fe.cal_flow(*args)
    assert torch.equal(T, T2)
    T3, fim, wim = fe.render.cal_flow(sc.to(DEV), sv.to(DEV), tc.to(DEV), tv.to(DEV), return_maps=True)
    assert torch.equal(T, T3) and np.array_equal(_np(fim), ofim) and np.array_equal(_bits(wim), owim.view(np.int32))
    assert bool((T[fim == -1] == -2).all())
    # forward(): warp a source frame into the target pose == torch grid_sample on the same flow
    img = torch.rand(2, 3, 128, 128, device=DEV) * 2 - 1
    out = fe(img, [sc.to(DEV), None, sv.to(DEV), None], [tc.to(DEV), None, tv.to(DEV), None])
    ref = F.grid_sample(img, T, mode="bilinear", padding_mode="border", align_corners=False)
    assert float((out - ref).abs().max()) <= 1e-5


# ------------------------------------------------------------------ warp (a10)
@pytest.mark.parametrize("align_corners", [False, True])
def test_grid_sample_matches_torch_cuda_and_oracle(align_corners):
    g = torch.Generator().manual_seed(0)
    N, C, Hs, Ws, H, W = 3, 5, 37, 29, 41, 23
    src = torch.randn((N, C, Hs, Ws), generator=g)
    grid = torch.rand((N, H, W, 2), generator=g) * 2.6 - 1.3
    grid[0, 0, :4] = -2.0                      # background sentinel (src/nmr.py:627)
    grid[0, 1, 0] = torch.tensor([1.0, -1.0])  # exact corners
    grid[0, 1, 1] = torch.tensor([-1.0, 1.0])
    grid[1, 2, 2] = torch.tensor([55.0, -97.0])
    out = ops.grid_sample_border(src.to(DEV), grid.to(DEV), align_corners)
    ref = F.grid_sample(src.to(DEV), grid.to(DEV), mode="bilinear", padding_mode="border", align_corners=align_corners)
    assert float((out - ref).abs().max()) <= 1e-5  # tolerance of the north star is 1e-4
    orc = oracle.grid_sample_border(src.numpy(), grid.numpy(), align_corners)
    assert np.array_equal(_bits(out), orc.view(np.int32))


def test_config1_single_reference_frame():
    """BASELINE config 1: one 256x256 RGB reference, batch 1, synthetic transfer flow."""
    rgb, _ = synth.reference_sets(1, 1, 0, 256, 256, seed=1)
    grid = synth.dense_flows(1, 1, 256, 256, seed=1)
    out = ops.grid_sample_border(rgb[:, 0].to(DEV), grid[:, 0].to(DEV))
    ref = F.grid_sample(rgb[:, 0], grid[:, 0], mode="bilinear", padding_mode="border", align_corners=False)
    assert float((out.cpu() - ref).abs().max()) <= 1e-5
    assert np.array_equal(_bits(out), oracle.grid_sample_border(rgb[:, 0].numpy(), grid[:, 0].numpy()).view(np.int32))


# ------------------------------------------------------------------ row F
def _rand_case(B, K, C, H, W, seed, Hs=None, Ws=None, R=None):
    rng = np.random.default_rng(seed)
    Hs, Ws, R = Hs or H, Ws or W, R or B
    return dict(
        rgb=rng.normal(size=(R, K, 3, Hs, Ws)).astype(np.float32),
        feat=rng.normal(size=(R, K, C, Hs, Ws)).astype(np.float32),
        grid=rng.uniform(-1.15, 1.15, size=(B, K, H, W, 2)).astype(np.float32),
        logits=rng.normal(size=(B, K, H, W)).astype(np.float32),
        vis=(rng.random((B, K, H, W)) > 0.25).astype(np.float32),
        mask=(rng.random((B, 1, H, W)) > 0.2).astype(np.float32),
        fim=rng.integers(-1, 3, size=(B, H, W)).astype(np.int32),
        fake=rng.normal(size=(B, 3, H, W)).astype(np.float32),
        conf=rng.random((B, 1, H, W)).astype(np.float32))


def _bf16_close(got_bits, ref_bits):
    """bf16 outputs agree with the oracle's single-rounded result to within one bf16 ulp of the value
    (2^-7 relative) plus 1e-6 absolute for results that cancel to ~0; returns (ok, mismatch fraction)."""
    a = oracle.bf16_bits_to_f32(np.ascontiguousarray(got_bits))
    b = oracle.bf16_bits_to_f32(np.ascontiguousarray(ref_bits))
    ok = bool(np.all(np.abs(a - b) <= np.abs(b) * 2.0 ** -7 + 1e-6))
    return ok, float((a != b).mean())  # value comparison: -0.0 == +0.0 (masked pixels)


@pytest.mark.parametrize("K,C", [(1, 64), (4, 64), (8, 64), (3, 32), (4, 128), (2, 256)])
def test_warp_fuse_hot_kernel_matches_oracle(K, C):
    B, H, W = 2, 41, 52  # ragged: W is not a multiple of the CTA strip, H is odd and not a multiple of the row chunk
    c = _rand_case(B, K, C, H, W, seed=K * 100 + C, Hs=33, Ws=47)
    feat_bits = oracle.f32_to_bf16_bits(c["feat"].transpose(0, 1, 3, 4, 2))        # dense [R,K,Hs,Ws,C]
    o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=feat_bits, feat_layout="nhwc", feat_bf16=True,
                         logits=c["logits"], vis=c["vis"], tgt_mask=c["mask"])
    feat = _cu(feat_bits.view(np.int16)).view(torch.bfloat16).permute(0, 1, 4, 2, 3)  # channels-last strides
    n0 = _lib.launch_count()
    out_rgb, out_feat = ops.warp_fuse(_cu(c["grid"]), rgb=_cu(c["rgb"]), feat=feat, logits=_cu(c["logits"]),
                                      vis=_cu(c["vis"]), tgt_mask=_cu(c["mask"]))
    assert _lib.launch_count() == n0 + 1, "RGB + features must be ONE fused launch"
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 2e-6
    got = _bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous())
    ok, frac = _bf16_close(got, o["out_feat"])
    assert ok and frac < 2e-3, frac
    # and against the same operation written with torch primitives in bf16 (the reference's path): <= 2e-2
    ft = feat.float()
    warped = torch.stack([F.grid_sample(ft[:, k], _cu(c["grid"])[:, k], padding_mode="border", align_corners=False)
                          for k in range(K)], 1)
    a = torch.softmax(_cu(c["logits"]), 1) * _cu(c["vis"])
    ref = (warped * a[:, :, None]).sum(1) * _cu(c["mask"])
    assert float((out_feat.float() - ref).abs().max()) <= 2e-2


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6, 7, 8])
def test_warp_fuse_wide_lane_kernel_matches_oracle(K):
    """C = 64 without a visibility input takes the wide-lane kernel (4 lanes x 32 B per pixel, 256-bit gathers;
    two references per lane beyond K = 4).  Ragged sizes: W is not a multiple of the 64-column tile, H is odd."""
    B, H, W, C = 2, 37, 75, 64
    c = _rand_case(B, K, C, H, W, seed=900 + K, Hs=29, Ws=58)
    feat_bits = oracle.f32_to_bf16_bits(c["feat"].transpose(0, 1, 3, 4, 2))
    o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=feat_bits, feat_layout="nhwc", feat_bf16=True,
                         logits=c["logits"], tgt_mask=c["mask"])
    feat = _cu(feat_bits.view(np.int16)).view(torch.bfloat16).permute(0, 1, 4, 2, 3)
    n0 = _lib.launch_count()
    out_rgb, out_feat = ops.warp_fuse(_cu(c["grid"]), rgb=_cu(c["rgb"]), feat=feat, logits=_cu(c["logits"]),
                                      tgt_mask=_cu(c["mask"]))
    assert _lib.launch_count() == n0 + 1
    # the RGB planes are produced inside the feature row loop from the feature weights (softmax * mask folded into
    # the taps, approximate softmax division): <= 1e-5 of the oracle, north-star tolerance 1e-4
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 1e-5
    ok, frac = _bf16_close(_bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous()), o["out_feat"])
    assert ok and frac < 2e-3, frac
    # the same call with an all-ones visibility map runs the 8-lane kernel: the two kernels agree
    ones = torch.ones(B, K, H, W, device=DEV)
    rgb2, feat2 = ops.warp_fuse(_cu(c["grid"]), rgb=_cu(c["rgb"]), feat=feat, logits=_cu(c["logits"]), vis=ones,
                                tgt_mask=_cu(c["mask"]))
    assert float((rgb2 - out_rgb).abs().max()) <= 1e-5
    ok2, frac2 = _bf16_close(_bf16_bits(feat2.permute(0, 2, 3, 1).contiguous()), _bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous()))
    assert ok2 and frac2 < 2e-3


@pytest.mark.parametrize("seed", list(range(24)))
def test_warp_fuse_random_configurations_match_the_oracle(seed):
    """Randomised sweep over the argument space of row F: batch, K = 1..8, channel count / layout / dtype of the features,
    ragged target and source sizes, reference sets shared through ref_index, and every optional input (logits, per-reference
    visibility or a face-index map, 1- or 3-channel target mask, confidence blend, align_corners) — whichever kernel the
    dispatcher picks must agree with the CPU oracle within the north star's bounds."""
    rng = np.random.default_rng(4242 + seed)
    B, K = int(rng.integers(1, 4)), int(rng.integers(1, 9))
    H, W = int(rng.integers(3, 70)), int(rng.integers(3, 90))
    Hs, Ws = int(rng.integers(2, 60)), int(rng.integers(2, 75))
    nhwc = bool(rng.integers(0, 2))
    C = int(rng.choice([32, 64, 128])) if nhwc else int(rng.integers(1, 20))
    R = int(rng.integers(1, B + 1))
    c = _rand_case(B, K, C, H, W, seed=9000 + seed, Hs=Hs, Ws=Ws, R=R)
    ref_index = rng.integers(0, R, size=B).astype(np.int32) if (R != B or rng.integers(0, 2)) else None
    use_logits, use_mask, use_blend = bool(rng.integers(0, 2)), int(rng.integers(0, 3)), bool(rng.integers(0, 2))
    vis_kind = int(rng.integers(0, 3))  # 0 none, 1 per-reference visibility, 2 face-index map
    ac = bool(rng.integers(0, 2))
    mask = None if use_mask == 0 else (c["mask"] if use_mask == 1 else np.repeat(c["mask"], 3, axis=1) * (rng.random((B, 3, H, W)) > 0.1))
    mask = None if mask is None else np.ascontiguousarray(mask, np.float32)
    kw_o = dict(logits=c["logits"] if use_logits else None, vis=c["vis"] if vis_kind == 1 else None,
                fim=c["fim"] if vis_kind == 2 else None, tgt_mask=mask, fake=c["fake"] if use_blend else None,
                conf=c["conf"] if use_blend else None, ref_index=ref_index, align_corners=ac)
    kw_g = {k: (None if v is None else (v if isinstance(v, bool) else _cu(v))) for k, v in kw_o.items()}
    if nhwc:
        fb = oracle.f32_to_bf16_bits(c["feat"].transpose(0, 1, 3, 4, 2))
        o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=fb, feat_layout="nhwc", feat_bf16=True, **kw_o)
        feat = _cu(fb.view(np.int16)).view(torch.bfloat16).permute(0, 1, 4, 2, 3)
    else:
        o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=c["feat"], **kw_o)
        feat = _cu(c["feat"])
    out_rgb, out_feat = ops.warp_fuse(_cu(c["grid"]), rgb=_cu(c["rgb"]), feat=feat, **kw_g)
    what = (seed, B, K, C, H, W, Hs, Ws, nhwc, R, use_logits, use_mask, use_blend, vis_kind, ac, _lib.last_kernel())
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 1e-5, what
    if nhwc:
        ok, frac = _bf16_close(_bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous()), o["out_feat"])
        assert ok and frac < 5e-3, (what, frac)
    else:
        assert float(np.abs(_np(out_feat) - o["out_feat"]).max()) <= 1e-4, what


@pytest.mark.parametrize("knob,ks,label", [("JAF_WF_WIDE8_SPLITK", (5, 6, 7, 8), "widesk<K=%d"),
                                           ("JAF_WF_PAIR", (1, 2, 3, 4), "pair<K=%d"),
                                           ("JAF_WF_WIDE_NOSHFL", (1, 2, 3, 4), "wide_ns<K=%d")])
def test_ab_flavours_of_the_wide_kernel_match_the_oracle_in_a_subprocess(knob, ks, label):
    """The A/B flavours kept behind environment knobs (read once per process, so each runs in a child process): K = 5..8 on
    8-lane groups whose halves split the references; 8-lane groups that own a PAIR of adjacent pixels (shared tap column in
    registers); the no-shuffle kernel in which every lane prepares all K references.  Ragged sizes, against the oracle with
    the bounds of the default kernel."""
    import subprocess
    import sys
    code = r"""
import sys, numpy as np, torch
root = %r
sys.path[:0] = [root, root + "/tests"]
import oracle
from test_gpu_parity import _rand_case, _cu, _np, _bf16_bits, _bf16_close
from jafpro_b200 import ops, _lib
torch.set_grad_enabled(False)
for K in %r:
    B, H, W, C = 2, 37, 75, 64
    c = _rand_case(B, K, C, H, W, seed=700 + K, Hs=29, Ws=58)
    fb = oracle.f32_to_bf16_bits(c["feat"].transpose(0, 1, 3, 4, 2))
    o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=fb, feat_layout="nhwc", feat_bf16=True, logits=c["logits"], tgt_mask=c["mask"])
    feat = _cu(fb.view(np.int16)).view(torch.bfloat16).permute(0, 1, 4, 2, 3)
    out_rgb, out_feat = ops.warp_fuse(_cu(c["grid"]), rgb=_cu(c["rgb"]), feat=feat, logits=_cu(c["logits"]), tgt_mask=_cu(c["mask"]))
    assert (%r %% K) in _lib.last_kernel(), _lib.last_kernel()
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 1e-5
    ok, frac = _bf16_close(_bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous()), o["out_feat"])
    assert ok and frac < 2e-3, (K, frac)
print("flavour ok")
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), tuple(ks), label)
    env = dict(os.environ)
    env[knob] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "flavour ok" in r.stdout, r.stdout + r.stderr


def test_warp_fuse_k1_is_exactly_warp_image_times_mask():
    c = _rand_case(2, 1, 64, 32, 32, seed=9)
    mask3 = np.repeat(c["mask"], 3, axis=1)
    rgb, grid = _cu(c["rgb"]), _cu(c["grid"])
    out_rgb, _ = ops.warp_fuse(grid, rgb=rgb, tgt_mask=_cu(mask3))
    ws = ops.grid_sample_border(rgb[:, 0].contiguous(), grid[:, 0].contiguous())
    assert torch.equal(out_rgb, ws * _cu(mask3))
    ref = oracle.warp_fuse(c["grid"], rgb=c["rgb"], tgt_mask=mask3)["out_rgb"]
    assert np.array_equal(_bits(out_rgb), ref.view(np.int32))
    # hot kernel, K = 1, no logits: bf16 features bit-exact too
    fb = oracle.f32_to_bf16_bits(c["feat"].transpose(0, 1, 3, 4, 2))
    o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=fb, feat_layout="nhwc", feat_bf16=True, tgt_mask=c["mask"])
    feat = _cu(fb.view(np.int16)).view(torch.bfloat16).permute(0, 1, 4, 2, 3)
    r2, f2 = ops.warp_fuse(grid, rgb=rgb, feat=feat, tgt_mask=_cu(c["mask"]))
    # (the fused kernel sums the four RGB taps across lanes, i.e. in a different order than ATen: 1 ulp)
    assert float(np.abs(_np(r2) - o["out_rgb"]).max()) <= 5e-7
    assert np.array_equal(oracle.bf16_bits_to_f32(_bf16_bits(f2.permute(0, 2, 3, 1).contiguous())),
                          oracle.bf16_bits_to_f32(o["out_feat"]))  # as values: masked pixels are +0 here, -0/+0 there


@pytest.mark.parametrize("layout,dtype", [("planar", "f32"), ("nhwc", "f32"), ("planar", "bf16"), ("nhwc", "bf16")])
def test_warp_fuse_generic_kernel_all_layouts(layout, dtype):
    B, K, C, H, W = 2, 3, 12, 19, 27  # C = 12: one of the reference's own channel counts, not hot-path eligible
    c = _rand_case(B, K, C, H, W, seed=21, R=3)
    ref_index = np.array([2, 0], np.int32)
    src = c["feat"] if layout == "planar" else c["feat"].transpose(0, 1, 3, 4, 2)
    if dtype == "bf16":
        src_o = oracle.f32_to_bf16_bits(src)
        t = _cu(src_o.view(np.int16)).view(torch.bfloat16)
    else:
        src_o = src
        t = _cu(src)
    if layout == "nhwc":
        t = t.permute(0, 1, 4, 2, 3)
    o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=src_o, feat_layout=layout, feat_bf16=(dtype == "bf16"),
                         logits=c["logits"], fim=c["fim"], tgt_mask=c["mask"], fake=c["fake"], conf=c["conf"],
                         ref_index=ref_index, return_warped=True)
    out_rgb, out_feat, warped = ops.warp_fuse(_cu(c["grid"]), rgb=_cu(c["rgb"]), feat=t, logits=_cu(c["logits"]),
                                              fim=_cu(c["fim"]), tgt_mask=_cu(c["mask"]), fake=_cu(c["fake"]),
                                              conf=_cu(c["conf"]), ref_index=_cu(ref_index), return_warped=True)
    assert np.array_equal(_bits(warped), o["warped_rgb"].view(np.int32))  # pure gathers: bit-exact
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 2e-6       # expf ulps only
    of = out_feat.permute(0, 2, 3, 1).contiguous() if layout == "nhwc" else out_feat
    if dtype == "bf16":
        ok, frac = _bf16_close(_bf16_bits(of), o["out_feat"])
        assert ok and frac < 2e-3, frac
    else:
        assert float(np.abs(_np(of) - o["out_feat"]).max()) <= 2e-6


def test_warp_fuse_default_visibility_and_background_skip():
    """vis defaults to fim != -1 (the -2 sentinel of src/nmr.py:627,644): background pixels come out 0
    whatever the references hold there, foreground pixels ignore the sentinel."""
    c = _rand_case(1, 4, 64, 32, 32, seed=5)
    c["grid"][0, :, :8] = -2.0
    c["fim"][0, :8] = -1
    c["fim"][0, 8:] = 7
    fb = oracle.f32_to_bf16_bits(c["feat"].transpose(0, 1, 3, 4, 2))
    feat = _cu(fb.view(np.int16)).view(torch.bfloat16).permute(0, 1, 4, 2, 3)
    out_rgb, out_feat = ops.warp_fuse(_cu(c["grid"]), rgb=_cu(c["rgb"]), feat=feat, logits=_cu(c["logits"]),
                                      fim=_cu(c["fim"]))
    assert float(out_rgb[0, :, :8].abs().max()) == 0 and float(out_feat[0, :, :8].float().abs().max()) == 0
    o = oracle.warp_fuse(c["grid"], rgb=c["rgb"], feat=fb, feat_layout="nhwc", feat_bf16=True, logits=c["logits"],
                         fim=c["fim"])
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 2e-6


def test_warp_fuse_rejects_bad_arguments():
    g = torch.zeros(1, 1, 4, 4, 2, device=DEV)
    with pytest.raises(RuntimeError):
        ops.warp_fuse(g)                                           # nothing to warp
    with pytest.raises(RuntimeError):
        ops.warp_fuse(g, rgb=torch.zeros(1, 2, 3, 4, 4, device=DEV))  # K mismatch
    with pytest.raises(RuntimeError):
        ops.warp_fuse(g, rgb=torch.zeros(1, 1, 3, 4, 4, device=DEV), logits=torch.zeros(1, 1, 4, 5, device=DEV))
    with pytest.raises(RuntimeError):
        ops.warp_fuse(g[..., :1], rgb=torch.zeros(1, 1, 3, 4, 4, device=DEV))


def test_full_size_properties_256_k4_c64():
    """BASELINE config 2 shape (one video's worth): size-independent properties at full resolution."""
    B, K, C, H, W = 6, 4, 64, 256, 256
    rgb, feat = synth.reference_sets(B, K, C, H, W, seed=3, device=DEV)
    logits = torch.randn(B, K, H, W, device=DEV)
    ident = synth.identity_grid(H, W, DEV)[None, None].expand(B, K, H, W, 2).contiguous()
    # (1) identity flow: the warp is the identity, so fused == softmax-weighted sum of the references
    out_rgb, out_feat = ops.warp_fuse(ident, rgb=rgb, feat=feat, logits=logits)
    a = torch.softmax(logits, 1)
    assert float((out_rgb - (rgb * a[:, :, None]).sum(1)).abs().max()) <= 1e-5
    assert float((out_feat.float() - (feat.float() * a[:, :, None]).sum(1)).abs().max()) <= 2e-2
    # (2) one-hot logits select one reference: fused == warp of that reference alone (bit-exact)
    grid = synth.dense_flows(B, K, H, W, seed=4, device=DEV)
    onehot = torch.full((B, K, H, W), -1e30, device=DEV)
    onehot[:, 2] = 0
    r_sel, f_sel = ops.warp_fuse(grid, rgb=rgb, feat=feat, logits=onehot)
    r_one, f_one = ops.warp_fuse(grid[:, 2:3].contiguous(), rgb=rgb[:, 2:3].contiguous(),
                                 feat=feat[:, 2:3].permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3))
    assert torch.equal(r_sel, r_one) and torch.equal(f_sel, f_one)
    # (3) linearity in the references (fp32 RGB): F(a*x + y) == a*F(x) + F(y) within rounding
    rgb2 = torch.randn_like(rgb)
    lhs, _ = ops.warp_fuse(grid, rgb=2.0 * rgb + rgb2, logits=logits)
    rhs = 2.0 * ops.warp_fuse(grid, rgb=rgb, logits=logits)[0] + ops.warp_fuse(grid, rgb=rgb2, logits=logits)[0]
    assert float((lhs - rhs).abs().max()) <= 2e-5
    # (4) against torch's grid_sample composition at full size: <= 1e-4 fp32, <= 2e-2 bf16 (north star)
    warped = torch.stack([F.grid_sample(rgb[:, k], grid[:, k], padding_mode="border", align_corners=False)
                          for k in range(K)], 1)
    ref = (warped * a[:, :, None]).sum(1)
    got, gotf = ops.warp_fuse(grid, rgb=rgb, feat=feat, logits=logits)
    assert float((got - ref).abs().max()) <= 1e-4
    wf = torch.stack([F.grid_sample(feat[:, k].float(), grid[:, k], padding_mode="border", align_corners=False)
                      for k in range(K)], 1)
    assert float((gotf.float() - (wf * a[:, :, None]).sum(1)).abs().max()) <= 2e-2
    # (5) a sub-batch of the oracle at full size (seconds on CPU)
    o = oracle.warp_fuse(_np(grid[:1]), rgb=_np(rgb[:1]), logits=_np(logits[:1]))
    assert float(np.abs(_np(got[:1]) - o["out_rgb"]).max()) <= 1e-5


@pytest.mark.parametrize("S,K,flow", [(256, 4, "dense"), (256, 4, "hard"), (512, 8, "dense"), (512, 8, "hard")])
def test_full_size_frames_match_the_oracle(S, K, flow):
    """BASELINE config 2 and config 5 frame shapes (256^2 K=4, 512^2 K=8; C=64) compared with the CPU oracle pixel for
    pixel at FULL size — RGB planes and bf16 features — on smooth and on hard (piecewise-affine, +-64 px) flows."""
    B, C = 1, 64
    rgb, feat = synth.reference_sets(B, K, C, S, S, seed=S + K, device=DEV)
    gen = synth.hard_flows if flow == "hard" else synth.dense_flows
    grid = gen(B, K, S, S, seed=17, device=DEV)
    logits = torch.randn(B, K, S, S, device=DEV)
    mask = (torch.rand(B, 1, S, S, device=DEV) > 0.05).float()
    out_rgb, out_feat = ops.warp_fuse(grid, rgb=rgb, feat=feat, logits=logits, tgt_mask=mask)
    assert ("wide<K=4" if K == 4 else "wide2<K=8") in _lib.last_kernel()
    fb = _bf16_bits(feat.permute(0, 1, 3, 4, 2).contiguous())
    o = oracle.warp_fuse(_np(grid), rgb=_np(rgb), feat=fb, feat_layout="nhwc", feat_bf16=True, logits=_np(logits),
                         tgt_mask=_np(mask))
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 2e-6
    ok, frac = _bf16_close(_bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous()), o["out_feat"])
    assert ok and frac < 2e-3, frac


def test_warp_fuse_host_pipeline_matches_device_path():
    """e2e entry point: host buffers in, host buffers out, chunked + pipelined inside the library."""
    B, K, C, H, W = 7, 4, 64, 64, 64
    rgb, feat = synth.reference_sets(2, K, C, H, W, seed=8)
    feat_dense = feat.permute(0, 1, 3, 4, 2).contiguous()  # [R,K,H,W,C]
    grid = synth.dense_flows(B, K, H, W, seed=8)
    logits = torch.randn(B, K, H, W)
    ref_index = torch.tensor([0, 0, 0, 1, 1, 1, 1], dtype=torch.int32)
    pin = lambda t: t.contiguous().pin_memory()
    o_rgb, o_feat = ops.warp_fuse_host(pin(grid), rgb=pin(rgb), feat=pin(feat_dense), feat_channels_last=True,
                                       logits=pin(logits), ref_index=ref_index, frames_per_chunk=2)
    d_rgb, d_feat = ops.warp_fuse(grid.to(DEV), rgb=rgb.to(DEV), feat=feat.to(DEV), logits=logits.to(DEV),
                                  ref_index=ref_index.to(DEV))
    assert torch.equal(o_rgb, d_rgb.cpu())
    assert torch.equal(o_feat, d_feat.permute(0, 2, 3, 1).contiguous().cpu())
    # per-frame reference sets (no ref_index), pageable memory, odd chunking
    rgb7, _ = synth.reference_sets(B, K, 0, H, W, seed=9)
    o2, _ = ops.warp_fuse_host(grid, rgb=rgb7, logits=logits, frames_per_chunk=3)
    d2, _ = ops.warp_fuse(grid.to(DEV), rgb=rgb7.to(DEV), logits=logits.to(DEV))
    assert torch.equal(o2, d2.cpu())


# ------------------------------------------------------------------ a11, a12
def test_mask_blend_matches_reference_fixture(golden_dir):
    d = _load(golden_dir, "mask_blend.npz")
    masked, pred = ops.mask_blend(_cu(d["tsf"]), _cu(d["mask"]), _cu(d["fake"]), _cu(d["weight"]))
    assert np.array_equal(_np(masked), d["tsf"] * d["mask"])
    assert float(np.abs(_np(pred) - d["pred"]).max()) <= 1e-6
    om, op_ = oracle.mask_blend(d["tsf"], d["mask"], d["fake"], d["weight"])
    assert np.array_equal(_bits(pred), op_.view(np.int32))
    # drop-in module with an injected confidence net
    from jafpro_b200.flow_net import Propagation3DFlowNet
    w = _cu(d["weight"])
    net = Propagation3DFlowNet(lambda x: w)
    out = net({'fake_tgt': _cu(d["fake"]), 'tsf_image': _cu(d["tsf"]), 'tgt_IUV': None, 'use_IUV': False,
               'use_mask': True, 'tgt_smpl_mask': _cu(d["mask"])})
    assert float(np.abs(_np(out['pred_target']) - d["pred"]).max()) <= 1e-6
    # ragged plane size (scalar path)
    t = torch.randn(1, 3, 5, 7, device=DEV)
    m = (torch.rand(1, 1, 5, 7, device=DEV) > 0.5).float()
    assert torch.equal(ops.mask_blend(t, m)[0], t * m)


def test_softmax_fuse_matches_reference_fixture(golden_dir):
    d = _load(golden_dir, "softmax_fuse.npz")
    out = ops.softmax_fuse(_cu(d["feat"]), _cu(d["logits"]))
    assert float(np.abs(_np(out) - d["out"]).max()) <= 1e-6
    assert float(np.abs(_np(out) - oracle.softmax_fuse(d["feat"], d["logits"])).max()) <= 1e-6
    f = torch.randn(2, 5 * 6, 7, 9, device=DEV)  # K=5, ragged plane
    l = torch.randn(2, 5, 7, 9, device=DEV)
    ref = (f.view(2, 5, 6, 7, 9) * torch.softmax(l, 1)[:, :, None]).sum(1)
    assert float((ops.softmax_fuse(f, l) - ref).abs().max()) <= 1e-5


# ------------------------------------------------------------------ a13
def test_convlstm_cell_matches_reference_fixture(golden_dir):
    d = _load(golden_dir, "convlstm.npz")
    h2, c2 = ops.convlstm_step(_cu(d["x"]), _cu(d["h"]), _cu(d["c"]), _cu(d["weight"]), _cu(d["bias"]))
    assert float(np.abs(_np(h2) - d["h_out"]).max()) <= 1e-5
    assert float(np.abs(_np(c2) - d["c_out"]).max()) <= 1e-5
    from jafpro_b200.convLSTM import ConvLSTM
    B, K, Cin, H, W = d["seq_x"].shape
    Ch = d["seq_h"].shape[1]
    lstm = ConvLSTM((H, W), Cin, Ch, (3, 3), 1, batch_first=True, bias=True).to(DEV)
    lstm.load_state_dict({"cell_list.0.conv.weight": torch.from_numpy(d["seq_weight"]),
                          "cell_list.0.conv.bias": torch.from_numpy(d["seq_bias"])})
    lstm.cell_list[0].tensor_cores = False   # exact fp32 kernel
    out, last = lstm(_cu(d["seq_x"]))  # default zero state on the module's device
    assert float(np.abs(_np(out) - d["seq_out"]).max()) <= 2e-5
    assert float(np.abs(_np(last[0][1]) - d["seq_c"]).max()) <= 2e-5
    lstm.cell_list[0].tensor_cores = True    # the default: split-bf16 on tensor cores, inside the 1e-4 fp32 bound
    out, last = lstm(_cu(d["seq_x"]))
    assert float(np.abs(_np(out) - d["seq_out"]).max()) <= 1e-4
    assert float(np.abs(_np(last[0][1]) - d["seq_c"]).max()) <= 1e-4


@pytest.mark.parametrize("Cin,Ch,H,W,k", [(12, 12, 50, 50, 3), (24, 48, 25, 25, 3), (7, 5, 13, 13, 5), (3, 2, 9, 20, 7)])
def test_convlstm_cell_reference_sizes_vs_torch(Cin, Ch, H, W, k):
    """The reference's own cell sizes (src/networks.py:1304-1313) against torch conv2d + gates in fp32."""
    torch.manual_seed(Cin)
    torch.backends.cudnn.allow_tf32 = False
    B = 3
    x, h, c = (torch.randn(B, n, H, W, device=DEV) for n in (Cin, Ch, Ch))
    wgt = torch.randn(4 * Ch, Cin + Ch, k, k, device=DEV) * 0.1
    bias = torch.randn(4 * Ch, device=DEV)
    def ref(b_):  # float64 so the checker itself carries no algorithm-dependent conv error
        cc = F.conv2d(torch.cat((x, h), 1).double(), wgt.double(), None if b_ is None else b_.double(), padding=k // 2)
        i, f, o, g = torch.split(cc, Ch, dim=1)
        c_r = torch.sigmoid(f) * c.double() + torch.sigmoid(i) * torch.tanh(g)
        return (torch.sigmoid(o) * torch.tanh(c_r)).float(), c_r.float()
    h_ref, c_ref = ref(bias)
    h2, c2 = ops.convlstm_step(x, h, c, wgt, bias)
    assert float((c2 - c_ref).abs().max()) <= 5e-5 and float((h2 - h_ref).abs().max()) <= 5e-5
    h3, c3 = ops.convlstm_step(x, h, c, wgt, None)
    h_ref3, c_ref3 = ref(None)
    assert float((c3 - c_ref3).abs().max()) <= 5e-5 and float((h3 - h_ref3).abs().max()) <= 5e-5


# ------------------------------------------------------------------ a13 on tensor cores (tcgen05)
def _convlstm_ref_nhwc(x, h, c, wgt, bias):
    """fp64 torch reference of src/convLSTM.py:41-56 on the bf16-rounded operands (NHWC in / out)."""
    Ch = h.shape[-1]
    xin = torch.cat((x.float(), h.float()), -1).permute(0, 3, 1, 2).double()
    cc = F.conv2d(xin, wgt.to(torch.bfloat16).double(), None if bias is None else bias.double(), padding=1)
    i, f, o, g = torch.split(cc, Ch, dim=1)
    c_r = torch.sigmoid(f) * c.permute(0, 3, 1, 2).double() + torch.sigmoid(i) * torch.tanh(g)
    h_r = torch.sigmoid(o) * torch.tanh(c_r)
    return h_r.permute(0, 2, 3, 1).float(), c_r.permute(0, 2, 3, 1).float()


@pytest.mark.parametrize("B,Cin,Ch,H,W,use_bias", [(2, 64, 64, 8, 64, True), (1, 128, 64, 4, 32, False),
                                                  (3, 64, 128, 2, 128, True), (2, 256, 256, 64, 64, True)])
def test_convlstm_tensor_core_cell(B, Cin, Ch, H, W, use_bias):
    """tcgen05 implicit-GEMM cell (bf16 operands, fp32 accumulate and state) vs an fp64 reference on the
    same bf16-rounded operands: c within 2e-3, h within 2e-2 (north star: <= 2e-2 for bf16 features).
    The last case is the BASELINE config-4 layer shape (256+256 -> 1024 channels, 64x64) at batch 2."""
    torch.manual_seed(B * 1000 + Cin)
    x = torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16)
    h = torch.randn(B, H, W, Ch, device=DEV).to(torch.bfloat16)
    c = torch.randn(B, H, W, Ch, device=DEV)
    wgt = torch.randn(4 * Ch, Cin + Ch, 3, 3, device=DEV) * (1.0 / (3.0 * (Cin + Ch) ** 0.5))
    bias = torch.randn(4 * Ch, device=DEV) if use_bias else None
    wpack = ops.convlstm_pack_weight(wgt, Cin, Ch)
    h2, c2 = ops.convlstm_step_tc(x, h, c, wpack, bias, Cin, Ch)
    h_ref, c_ref = _convlstm_ref_nhwc(x, h, c, wgt, bias)
    assert float((c2 - c_ref).abs().max()) <= 2e-3
    assert float((h2.float() - h_ref).abs().max()) <= 2e-2
    # the fp32 CUDA-core cell on the same (bf16-rounded) operands agrees as well
    h3, c3 = ops.convlstm_step(x.float().permute(0, 3, 1, 2).contiguous(), h.float().permute(0, 3, 1, 2).contiguous(),
                               c.permute(0, 3, 1, 2).contiguous(), wgt.to(torch.bfloat16).float(), bias)
    assert float((c3.permute(0, 2, 3, 1) - c2).abs().max()) <= 2e-3
    # K = 4 sequential steps from zero state through the module wrapper (config 4: K = 4 references)
    from jafpro_b200.convLSTM import ConvLSTMCellTC
    cell = ConvLSTMCellTC(Cin, Ch, wgt, bias)
    hs = torch.zeros(B, H, W, Ch, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(B, H, W, Ch, device=DEV)
    hr, cr = hs, cs
    for t in range(4):
        xt = torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16)
        hs, cs = cell(xt, (hs, cs))
        hr, cr = _convlstm_ref_nhwc(xt, hr, cr, wgt, bias)
        hr = hr.to(torch.bfloat16)
    assert float((cs - cr).abs().max()) <= 2e-2 and float((hs.float() - hr.float()).abs().max()) <= 3e-2


def test_convlstm_tensor_core_rejects_unsupported_shapes():
    x = torch.zeros(1, 4, 48, 64, device=DEV, dtype=torch.bfloat16)  # W = 48 does not divide 128
    c = torch.zeros(1, 4, 48, 64, device=DEV)
    wp = torch.zeros(_lib.lib().jaf_convlstm_wpack_bytes(64, 64), dtype=torch.uint8, device=DEV)
    with pytest.raises(RuntimeError, match="W must divide 128"):
        ops.convlstm_step_tc(x, x, c, wp, None, 64, 64)
    with pytest.raises(RuntimeError, match="multiples of 64"):
        ops.convlstm_pack_weight(torch.zeros(4 * 24, 36, 3, 3, device=DEV), 12, 24)


# ------------------------------------------------------------------ §8f rank 1: IUV texture lookup
@pytest.mark.parametrize("align_corners", [False, True])
def test_texture_warp_matches_reference_and_oracle(golden_dir, align_corners):
    from jafpro_b200.texture import texture_warp_pytorch
    d = _load(golden_dir, "texture_warp.npz")
    got = texture_warp_pytorch([torch.from_numpy(t) for t in d["tex"]], d["iuv"], "cuda", align_corners=align_corners)
    orc = oracle.texture_warp(d["tex"], d["iuv"], align_corners=align_corners)
    assert np.array_equal(_bits(got), orc.view(np.int32))  # bit-exact vs the oracle
    if not align_corners:
        assert float(np.abs(_np(got) - d["out"]).max()) <= 1e-6  # the reference's own output
    # a batch of DanceVideo-sized frames with 200x200 part textures (the reference's texture size)
    rng = np.random.default_rng(5)
    tex = rng.normal(size=(24, 3, 200, 200)).astype(np.float32)
    iuv = rng.integers(0, 256, (4, 256, 256, 3)).astype(np.uint8)
    iuv[..., 0] = rng.integers(0, 26, (4, 256, 256))  # includes an out-of-range part id (25)
    out = ops.texture_warp(_cu(tex), _cu(iuv), align_corners)
    assert np.array_equal(_bits(out), oracle.texture_warp(tex, iuv, align_corners).view(np.int32))


def test_cal_flow_multi_equals_per_pair_cal_flow_and_feeds_warp_fuse():
    """K source poses -> one target raster: bit-identical to the reference-shaped per-pair call, and the
    two-step device pipeline (flows + fused warp) matches the oracle end to end."""
    from jafpro_b200.fusion import warp_fuse_from_poses
    _, faces_idx = load_smpl_template()
    B, K, S = 2, 3, 96
    cam, verts = synth.smpl_poses(B * (K + 1), seed=21)
    tcam, tverts = cam[:B].contiguous(), verts[:B].contiguous()
    scam = cam[B:].reshape(B, K, 3).contiguous()
    sverts = verts[B:].reshape(B, K, -1, 3).contiguous()
    f_idx = _cu(faces_idx)
    T, fim, wim = ops.cal_flow_multi(scam.to(DEV), sverts.to(DEV), tcam.to(DEV), tverts.to(DEV), f_idx, S)
    for k in range(K):
        Tk, fk, wk = ops.cal_flow(scam[:, k].contiguous().to(DEV), sverts[:, k].contiguous().to(DEV), tcam.to(DEV),
                                  tverts.to(DEV), f_idx, S, return_maps=True)
        assert torch.equal(T[:, k], Tk) and torch.equal(fim, fk) and torch.equal(wim, wk)
        oT, ofim, _ = oracle.cal_flow(scam[:, k].numpy(), sverts[:, k].numpy(), tcam.numpy(), tverts.numpy(), faces_idx, S)
        assert np.array_equal(_bits(T[:, k]), oT.view(np.int32)) and np.array_equal(_np(fim), ofim)
    rend = SMPLRenderer(image_size=S).to(DEV)
    rgb, feat = synth.reference_sets(B, K, 64, S, S, seed=4, device=DEV)
    logits = torch.randn(B, K, S, S, device=DEV)
    out_rgb, out_feat, T2, fim2 = warp_fuse_from_poses(rend, scam.to(DEV), sverts.to(DEV), tcam.to(DEV), tverts.to(DEV),
                                                       rgb=rgb, feat=feat, logits=logits)
    assert torch.equal(T2, T) and torch.equal(fim2, fim)
    fb = _bf16_bits(feat.permute(0, 1, 3, 4, 2).contiguous())
    o = oracle.warp_fuse(_np(T), rgb=_np(rgb), feat=fb, feat_layout="nhwc", feat_bf16=True, logits=_np(logits),
                         fim=_np(fim))
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 2e-6
    ok, frac = _bf16_close(_bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous()), o["out_feat"])
    assert ok and frac < 2e-3
    bg = _np(fim) == -1
    assert float(np.abs(_np(out_rgb).transpose(0, 2, 3, 1)[bg]).max()) == 0.0  # background stays empty


# ------------------------------------------------------------------ a13: G reference-sized cells per launch (tcgen05, split-bf16)
@pytest.mark.parametrize("G,B,Cin,Ch,H,W", [(1, 1, 12, 12, 20, 20), (3, 2, 12, 12, 50, 37), (2, 1, 24, 24, 100, 100),
                                            (24, 1, 48, 48, 25, 25), (2, 3, 96, 96, 13, 13), (2, 1, 3, 12, 200, 200),
                                            (1, 2, 20, 8, 9, 140), (2, 2, 8, 32, 30, 41), (1, 1, 4, 4, 7, 5),
                                            (2, 1, 16, 32, 38, 37), (160, 1, 12, 12, 16, 16)])
def test_convlstm_grouped_cells_vs_fp64(G, B, Cin, Ch, H, W):
    """Accumulate_LSTM_no_loss' per-part cells (src/networks.py:1304-1313: 12@200^2, 24@100^2, 24@50^2, 48@25^2,
    96@13^2; 24 parts with their own weights) in one launch.  Tolerance: the fp32 bound of the north star, 1e-4."""
    torch.manual_seed(G * 100 + Ch)
    x, h, c = (torch.randn(G, B, n, H, W, device=DEV) for n in (Cin, Ch, Ch))
    wgt = torch.randn(G, 4 * Ch, Cin + Ch, 3, 3, device=DEV) * (1.5 / (9 * (Cin + Ch)) ** 0.5)
    bias = torch.randn(G, 4 * Ch, device=DEV)
    wpack = ops.convlstm_gpack_weight(wgt, Cin, Ch)
    for b_ in (bias, None):
        h2, c2 = ops.convlstm_step_grouped(x, h, c, wpack, b_, Cin, Ch)
        for g in range(G):
            cc = F.conv2d(torch.cat((x[g], h[g]), 1).double(), wgt[g].double(), None if b_ is None else b_[g].double(), padding=1)
            i, f, o, gg = torch.split(cc, Ch, dim=1)
            c_r = torch.sigmoid(f) * c[g].double() + torch.sigmoid(i) * torch.tanh(gg)
            h_r = torch.sigmoid(o) * torch.tanh(c_r)
            ec, eh = float((c2[g].double() - c_r).abs().max()), float((h2[g].double() - h_r).abs().max())
            assert ec <= 1e-4 and eh <= 1e-4, (g, ec, eh)
    # the exact-fp32 CUDA-core cell agrees as well (same interface, G = 1 per call)
    h3, c3 = ops.convlstm_step(x[0], h[0], c[0], wgt[0], None)
    assert float((c3 - c2[0]).abs().max()) <= 1e-4 and float((h3 - h2[0]).abs().max()) <= 1e-4


@pytest.mark.parametrize("seed", list(range(16)))
def test_convlstm_grouped_random_shapes_vs_fp64(seed):
    """Randomised sweep over what the planner has to place: narrow cells (operand-swapped persistent kernel), wide ones
    (un-swapped kernel: full / half launch shapes, hidden-channel slices, weight-ring depths), odd map sizes, several
    bands, batch > 1, Cin != Ch.  Unsupported combinations must say so; supported ones meet 1e-4 against fp64."""
    rng = np.random.default_rng(77 + seed)
    G, B = int(rng.integers(1, 7)), int(rng.integers(1, 4))
    Ch = int(rng.choice([4, 8, 12, 16, 24, 32, 40, 48, 64, 96, 128]))
    Cin = int(rng.integers(1, 2 * Ch + 1))
    H, W = int(rng.integers(3, 64)), int(rng.integers(3, 110))
    if not ops.convlstm_grouped_supported(G, B, Cin, Ch, H, W):
        with pytest.raises(RuntimeError):
            ops.convlstm_step_grouped(torch.zeros(G, B, Cin, H, W, device=DEV), torch.zeros(G, B, Ch, H, W, device=DEV),
                                      torch.zeros(G, B, Ch, H, W, device=DEV),
                                      ops.convlstm_gpack_weight(torch.zeros(G, 4 * Ch, Cin + Ch, 3, 3, device=DEV), Cin, Ch),
                                      None, Cin, Ch)
        return
    torch.manual_seed(seed)
    x, h, c = (torch.randn(G, B, n, H, W, device=DEV) for n in (Cin, Ch, Ch))
    wgt = torch.randn(G, 4 * Ch, Cin + Ch, 3, 3, device=DEV) * (1.5 / (9 * (Cin + Ch)) ** 0.5)
    bias = torch.randn(G, 4 * Ch, device=DEV)
    h2, c2 = ops.convlstm_step_grouped(x, h, c, ops.convlstm_gpack_weight(wgt, Cin, Ch), bias, Cin, Ch)
    for g in range(G):
        cc = F.conv2d(torch.cat((x[g], h[g]), 1).double(), wgt[g].double(), bias[g].double(), padding=1)
        i, f, o, gg = torch.split(cc, Ch, dim=1)
        c_r = torch.sigmoid(f) * c[g].double() + torch.sigmoid(i) * torch.tanh(gg)
        h_r = torch.sigmoid(o) * torch.tanh(c_r)
        ec, eh = float((c2[g].double() - c_r).abs().max()), float((h2[g].double() - h_r).abs().max())
        assert ec <= 1e-4 and eh <= 1e-4, (seed, G, B, Cin, Ch, H, W, g, ec, eh)


def test_run_concurrently_gives_the_serial_results():
    """The pyramid levels of a recurrent step are independent: one side stream per level (fork / join by events), eager
    and captured in a CUDA graph, gives the bits of the serial loop."""
    levels = []
    for Ch, S in [(12, 40), (24, 20), (48, 9)]:
        torch.manual_seed(Ch)
        x, h, c = (torch.randn(3, 1, Ch, S, S, device=DEV) for _ in range(3))
        w = torch.randn(3, 4 * Ch, 2 * Ch, 3, 3, device=DEV) * 0.05
        levels.append((x, h, c, ops.convlstm_gpack_weight(w, Ch, Ch), Ch))
    step = lambda lv: ops.convlstm_step_grouped(lv[0], lv[1], lv[2], lv[3], None, lv[4], lv[4])
    serial = [step(lv) for lv in levels]
    fns = [(lambda lv=lv: step(lv)) for lv in levels]
    forked = ops.run_concurrently(fns)
    torch.cuda.synchronize()
    for (h1, c1), (h2, c2) in zip(serial, forked):
        assert torch.equal(h1, h2) and torch.equal(c1, c2)
    g = ops.FrameGraph(lambda: ops.run_concurrently(fns))
    res = g.replay()
    torch.cuda.synchronize()
    for (h1, c1), (h2, c2) in zip(serial, res):
        assert torch.equal(h1, h2) and torch.equal(c1, c2)


def test_convlstm_grouped_rejects_unsupported_shapes():
    assert _lib.lib().jaf_convlstm_gpack_bytes(1, 12, 10) == 0   # Ch % 4
    assert _lib.lib().jaf_convlstm_gpack_bytes(1, 12, 256) == 0  # 4*Ch > 512 TMEM columns
    with pytest.raises(RuntimeError):
        ops.convlstm_gpack_weight(torch.zeros(1, 40, 22, 3, 3, device=DEV), 12, 10)


def test_convlstm_grouped_module_matches_per_part_lstms():
    """24-part usage: ConvLSTMGrouped.from_lstms == running each part's ConvLSTM on its own (K = 3 steps)."""
    from jafpro_b200.convLSTM import ConvLSTM, ConvLSTMGrouped
    torch.manual_seed(5)
    G, B, T, Ch, S = 4, 1, 3, 24, 50
    lstms = [ConvLSTM((S, S), Ch, [Ch], [(3, 3)], 1, batch_first=True, bias=True).to(DEV) for _ in range(G)]
    grouped = ConvLSTMGrouped.from_lstms(lstms)
    x = torch.randn(G, B, T, Ch, S, S, device=DEV)
    out, (h, c) = grouped(x)
    for g in range(G):
        o_ref, last = lstms[g](x[g])
        assert float((out[g] - o_ref).abs().max()) <= 1e-4
        assert float((h[g] - last[0][0]).abs().max()) <= 1e-4 and float((c[g] - last[0][1]).abs().max()) <= 1e-4
    # the single cell's opt-in tensor-core path (G = 1) gives the grouped result bit for bit; weights are re-packed
    # when they change
    cell = lstms[1].cell_list[0]
    assert cell.tensor_cores
    o_tc, _ = lstms[1](x[1])
    assert torch.equal(o_tc, out[1])
    with torch.no_grad():
        cell.conv.weight.mul_(0.5)
    o_half, _ = lstms[1](x[1])
    cell.tensor_cores = False
    o_half_exact, _ = lstms[1](x[1])
    assert float((o_half - o_half_exact).abs().max()) <= 1e-4 and not torch.equal(o_half, o_tc)


@pytest.mark.parametrize("K,S,with_rgb", [(4, 64, True), (2, 96, True), (8, 64, False), (1, 40, True)])
def test_warp_fuse_from_poses_is_the_two_call_path_bit_for_bit(K, S, with_rgb):
    """jaf_warp_fuse_from_poses composes the transfer flows per tile in shared memory: outputs, and the optional T / fim,
    must equal jaf_cal_flow_multi + jaf_warp_fuse(fim) exactly (same pinned arithmetic, no flow round trip)."""
    B, C = 3, 64
    _, faces_idx = load_smpl_template()
    f_idx = _cu(faces_idx)
    cam, verts = synth.smpl_poses(B * (K + 1), seed=40 + K, device=DEV)
    tc, tv = cam[:B].contiguous(), verts[:B].contiguous()
    sc, sv = cam[B:].reshape(B, K, 3).contiguous(), verts[B:].reshape(B, K, -1, 3).contiguous()
    rgb, feat = synth.reference_sets(B, K, C, S, S, seed=41, device=DEV)
    logits = torch.randn(B, K, S, S, device=DEV)
    mask = (torch.rand(B, 1, S, S, device=DEV) > 0.1).float()
    T, fim, _ = ops.cal_flow_multi(sc, sv, tc, tv, f_idx, S, return_wim=False)
    assert int((fim != -1).sum()) > 0
    r2, f2 = ops.warp_fuse(T, rgb=rgb if with_rgb else None, feat=feat, logits=logits, fim=fim, tgt_mask=mask)
    n0 = _lib.launch_count()
    r1, f1, T1, fim1 = ops.warp_fuse_from_poses(sc, sv, tc, tv, f_idx, S, rgb=rgb if with_rgb else None, feat=feat,
                                                logits=logits, tgt_mask=mask, return_flow=True)
    assert _lib.launch_count() - n0 == 3, "scatter + huge + ONE fused warp kernel (no resolve / compose pass)"
    assert "POSES=1" in _lib.last_kernel()
    assert torch.equal(fim1, fim) and np.array_equal(_bits(T1), _bits(T))
    assert np.array_equal(_bf16_bits(f1.permute(0, 2, 3, 1).contiguous()), _bf16_bits(f2.permute(0, 2, 3, 1).contiguous()))
    if with_rgb:
        assert np.array_equal(_bits(r1), _bits(r2))
    # without the optional outputs nothing changes; shared reference sets / poses through ref_index; confidence blend
    r3, f3 = ops.warp_fuse_from_poses(sc, sv, tc, tv, f_idx, S, rgb=rgb if with_rgb else None, feat=feat, logits=logits,
                                      tgt_mask=mask)
    assert torch.equal(f3, f1) and (not with_rgb or torch.equal(r3, r1))
    if with_rgb:
        ridx = torch.tensor([1, 1, 0], dtype=torch.int32, device=DEV)
        fake, conf = torch.randn(B, 3, S, S, device=DEV), torch.rand(B, 1, S, S, device=DEV)
        r4, f4 = ops.warp_fuse_from_poses(sc[:2].contiguous(), sv[:2].contiguous(), tc, tv, f_idx, S, rgb=rgb[:2].contiguous(),
                                          feat=feat[:2], logits=logits, fake=fake, conf=conf, ref_index=ridx)
        idx = ridx.long()
        T5, fim5, _ = ops.cal_flow_multi(sc[idx].contiguous(), sv[idx].contiguous(), tc, tv, f_idx, S, return_wim=False)
        r5, f5 = ops.warp_fuse(T5, rgb=rgb[:2].contiguous(), feat=feat[:2], logits=logits, fim=fim5, fake=fake, conf=conf,
                               ref_index=ridx)
        assert torch.equal(r4, r5) and torch.equal(f4, f5)


def test_warp_fuse_from_poses_host_matches_the_device_path():
    """e2e entry point of the pose-driven operation: poses / references / logits in HOST memory, fused RGB back to the
    host, fused features either downloaded or left device-resident; chunked and pipelined inside the library."""
    B, K, C, S = 7, 4, 64, 64
    _, faces_idx = load_smpl_template()
    cam, verts = synth.smpl_poses(B + 2 * K, seed=77)
    tc, tv = cam[:B].contiguous(), verts[:B].contiguous()
    sc, sv = cam[B:].reshape(2, K, 3).contiguous(), verts[B:].reshape(2, K, -1, 3).contiguous()
    rgb, feat = synth.reference_sets(2, K, C, S, S, seed=78)
    feat_dense = feat.permute(0, 1, 3, 4, 2).contiguous()
    logits = torch.randn(B, K, S, S)
    mask = (torch.rand(B, 1, S, S) > 0.1).float()
    ridx = torch.tensor([0, 0, 0, 1, 1, 1, 1], dtype=torch.int32)
    f_idx = torch.from_numpy(faces_idx)
    pin = lambda t: t.contiguous().pin_memory()
    d_rgb, d_feat = ops.warp_fuse_from_poses(sc.to(DEV), sv.to(DEV), tc.to(DEV), tv.to(DEV), f_idx.to(DEV), S, rgb=rgb.to(DEV),
                                             feat=feat.to(DEV), logits=logits.to(DEV), tgt_mask=mask.to(DEV),
                                             ref_index=ridx.to(DEV))
    o_rgb, o_feat = ops.warp_fuse_from_poses_host(pin(sc), pin(sv), pin(tc), pin(tv), f_idx, S, rgb=pin(rgb), feat=pin(feat_dense),
                                                  logits=pin(logits), tgt_mask=pin(mask), ref_index=ridx, frames_per_chunk=2)
    assert torch.equal(o_rgb, d_rgb.cpu())
    assert torch.equal(o_feat, d_feat.permute(0, 2, 3, 1).contiguous().cpu())
    # device-resident feature output, pageable inputs, automatic chunking
    dev_feat = torch.empty(B, S, S, C, dtype=torch.bfloat16, device=DEV)
    o2, f2 = ops.warp_fuse_from_poses_host(sc, sv, tc, tv, f_idx, S, rgb=rgb, feat=feat_dense, logits=logits, tgt_mask=mask,
                                           ref_index=ridx, out_feat_device=dev_feat)
    assert f2 is dev_feat and torch.equal(o2, d_rgb.cpu())
    assert torch.equal(dev_feat, d_feat.permute(0, 2, 3, 1).contiguous())
    # one set of references / reference poses per target frame (no ref_index)
    cam2, verts2 = synth.smpl_poses(3 * (K + 1), seed=79)
    rgb3, feat3 = synth.reference_sets(3, K, C, S, S, seed=80)
    a = dict(sc=cam2[3:].reshape(3, K, 3).contiguous(), sv=verts2[3:].reshape(3, K, -1, 3).contiguous(), tc=cam2[:3].contiguous(),
             tv=verts2[:3].contiguous())
    o3, f3 = ops.warp_fuse_from_poses_host(a["sc"], a["sv"], a["tc"], a["tv"], f_idx, S, rgb=rgb3,
                                           feat=feat3.permute(0, 1, 3, 4, 2).contiguous(), frames_per_chunk=2)
    d3, df3 = ops.warp_fuse_from_poses(a["sc"].to(DEV), a["sv"].to(DEV), a["tc"].to(DEV), a["tv"].to(DEV), f_idx.to(DEV), S,
                                       rgb=rgb3.to(DEV), feat=feat3.to(DEV))
    assert torch.equal(o3, d3.cpu()) and torch.equal(f3, df3.permute(0, 2, 3, 1).contiguous().cpu())
    # argument validation of the host entry points (raw pointers feed pipelined memcpys)
    with pytest.raises(RuntimeError):
        ops.warp_fuse_host(torch.zeros(1, 1, 8, 8, 2), rgb=torch.zeros(1, 1, 3, 8, 8, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        ops.warp_fuse_host(torch.zeros(1, 1, 8, 8, 2), rgb=torch.zeros(1, 1, 3, 8, 16)[..., ::2])
    with pytest.raises(RuntimeError):
        ops.warp_fuse_host(torch.zeros(2, 1, 8, 8, 2), rgb=torch.zeros(1, 1, 3, 8, 8))   # one set for two frames, no ref_index
    with pytest.raises(RuntimeError):
        ops.warp_fuse_host(torch.zeros(1, 1, 8, 8, 2), rgb=torch.zeros(1, 1, 3, 8, 8), logits=torch.zeros(1, 1, 8, 9))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_host_pipeline_runs_on_every_device_of_the_process():
    """One pipeline per device (the reference wraps float_estimate in nn.DataParallel, test/conv_pro_test.py:140-141)."""
    grid = synth.dense_flows(3, 2, 32, 32, seed=1)
    rgb, _ = synth.reference_sets(3, 2, 0, 32, 32, seed=2)
    outs = []
    for d in (0, 1):
        with torch.cuda.device(d):
            outs.append(ops.warp_fuse_host(grid, rgb=rgb)[0].clone())
    assert torch.equal(outs[0], outs[1])


def test_frame_graph_replays_the_batch_1_sequence_bit_identically():
    """SURVEY §7 hard part 4: the per-frame loop of test/conv_pro_test.py:255-278 (cal_flow -> warp_image -> mask / blend)
    at batch 1, captured once into a CUDA graph — same bits as the eager calls, one launch per replay."""
    S, T = 64, 5
    _, faces_idx = load_smpl_template()
    f_idx = _cu(faces_idx)
    cam, verts = synth.smpl_poses(T + 1, seed=31, device=DEV)
    src = torch.randn(1, 1, 3, S, S, device=DEV)
    fake, conf = torch.randn(T, 3, S, S, device=DEV), torch.rand(T, 1, S, S, device=DEV)
    outs = [torch.empty(1, 3, S, S, device=DEV) for _ in range(T)]

    def sequence():
        for t in range(T):
            flow, fim, _ = ops.cal_flow(cam[T:], verts[T:], cam[t:t + 1], verts[t:t + 1], f_idx, S, return_maps=True)
            ops.warp_fuse(flow[:, None], rgb=src, fim=fim, fake=fake[t:t + 1], conf=conf[t:t + 1], out_rgb=outs[t])
        return outs

    eager = [o.clone() for o in sequence()]
    for o in outs:
        o.zero_()
    g = ops.FrameGraph(sequence)
    for o in outs:
        o.zero_()
    n0 = _lib.launch_count()
    res = g.replay()
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0, "a replay goes through cudaGraphLaunch, not through the library's launch paths"
    for a, b in zip(eager, res):
        assert torch.equal(a, b)
    # new inputs are written into the captured tensors
    verts[0].add_(0.01)
    e2 = [o.clone() for o in sequence()]
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(e2, outs)) and not torch.equal(e2[0], eager[0])


@pytest.mark.parametrize("lanes", [1, 4])
def test_frame_graph_with_parallel_lanes_matches_eager(lanes):
    """The frames of the reference's loop are independent (conv_pro_test.py:255-278 never feeds frame t-1 into frame t):
    captured on `lanes` parallel graph branches (per-stream workspaces) they give the same bits as the eager loop."""
    S, T = 64, 9
    _, faces_idx = load_smpl_template()
    f_idx = _cu(faces_idx)
    cam, verts = synth.smpl_poses(T + 1, seed=37, device=DEV)
    src = torch.randn(1, 1, 3, S, S, device=DEV)
    fake, conf = torch.randn(T, 3, S, S, device=DEV), torch.rand(T, 1, S, S, device=DEV)
    outs = [torch.empty(1, 3, S, S, device=DEV) for _ in range(T)]
    sc, sv = cam[T:].reshape(1, 1, 3).contiguous(), verts[T:].reshape(1, 1, -1, 3).contiguous()

    def frame(t):
        ops.warp_fuse_from_poses(sc, sv, cam[t:t + 1], verts[t:t + 1], f_idx, S, rgb=src, fake=fake[t:t + 1],
                                 conf=conf[t:t + 1], out_rgb=outs[t])
        return outs[t]

    eager = [frame(t).clone() for t in range(T)]
    g = ops.FrameGraph(frame, frames=T, lanes=lanes)
    for rep in range(3):
        for o in outs:
            o.zero_()
        res = g.replay()
        torch.cuda.synchronize()
        for a, b in zip(eager, res):
            assert torch.equal(a, b), (lanes, rep)


@pytest.mark.parametrize("K", [1, 3])
def test_warp_fuse_from_poses_rgb_only(K):
    """The reference's per-frame chain at its own shapes (cal_flow -> warp_image -> mask -> confidence blend,
    test/conv_pro_test.py:255-278; no feature tensor) through the pose-driven call: raster pass + ONE fused kernel."""
    B, S = 2, 64
    _, faces_idx = load_smpl_template()
    f_idx = _cu(faces_idx)
    cam, verts = synth.smpl_poses(B * (K + 1), seed=60 + K, device=DEV)
    tc, tv = cam[:B].contiguous(), verts[:B].contiguous()
    sc, sv = cam[B:].reshape(B, K, 3).contiguous(), verts[B:].reshape(B, K, -1, 3).contiguous()
    rgb = torch.randn(B, K, 3, S, S, device=DEV)
    logits = torch.randn(B, K, S, S, device=DEV)
    fake, conf = torch.randn(B, 3, S, S, device=DEV), torch.rand(B, 1, S, S, device=DEV)
    n0 = _lib.launch_count()
    r1, f1 = ops.warp_fuse_from_poses(sc, sv, tc, tv, f_idx, S, rgb=rgb, logits=logits, fake=fake, conf=conf)
    assert f1 is None and _lib.launch_count() - n0 == 3
    T, fim, _ = ops.cal_flow_multi(sc, sv, tc, tv, f_idx, S, return_wim=False)
    r2, _ = ops.warp_fuse(T, rgb=rgb, logits=logits, fim=fim, fake=fake, conf=conf)
    assert float((r1 - r2).abs().max()) <= 1e-6      # (the RGB-only kernel of the two-call path divides the softmax once)
    o = oracle.warp_fuse(_np(T), rgb=_np(rgb), logits=_np(logits), fim=_np(fim), fake=_np(fake), conf=_np(conf))
    assert float(np.abs(_np(r1) - o["out_rgb"]).max()) <= 2e-6


def test_warp_fuse_from_poses_self_cleaning_workspace():
    """The module-owned raster workspace is left empty by the fused kernel and the next call skips the 8 B / pixel clear
    (JAF_POSES_LEAVE_CLEAN / JAF_POSES_KEYS_CLEAN); any other user of the workspace dirties it.  Every call in any order
    must equal the two-call path."""
    K, C, S = 4, 64, 64
    _, faces_idx = load_smpl_template()
    f_idx = _cu(faces_idx)
    rgb, feat = synth.reference_sets(4, K, C, S, S, seed=5, device=DEV)

    def case(B, seed):
        cam, verts = synth.smpl_poses(B * (K + 1), seed=seed, device=DEV)
        return (cam[B:].reshape(B, K, 3).contiguous(), verts[B:].reshape(B, K, -1, 3).contiguous(), cam[:B].contiguous(),
                verts[:B].contiguous())

    def check(B, seed):
        sc, sv, tc, tv = case(B, seed)
        r1, f1 = ops.warp_fuse_from_poses(sc, sv, tc, tv, f_idx, S, rgb=rgb[:B].contiguous(), feat=feat[:B])
        T, fim, _ = ops.cal_flow_multi(sc, sv, tc, tv, f_idx, S, return_wim=False)   # dirties the shared workspace
        r2, f2 = ops.warp_fuse(T, rgb=rgb[:B].contiguous(), feat=feat[:B], fim=fim)
        assert torch.equal(r1, r2) and torch.equal(f1, f2), (B, seed)

    check(4, 1)
    sc, sv, tc, tv = case(4, 2)
    a1 = ops.warp_fuse_from_poses(sc, sv, tc, tv, f_idx, S, rgb=rgb, feat=feat)       # clears, leaves clean
    sc3, sv3, tc3, tv3 = case(2, 3)
    b1 = ops.warp_fuse_from_poses(sc3, sv3, tc3, tv3, f_idx, S, rgb=rgb[:2].contiguous(), feat=feat[:2])   # skips the clear
    a2 = ops.warp_fuse_from_poses(sc, sv, tc, tv, f_idx, S, rgb=rgb, feat=feat)       # skips the clear (4 frames were clean)
    assert torch.equal(a1[0], a2[0]) and torch.equal(a1[1], a2[1])
    ops.render_fim_wim(tc, tv, f_idx, S)                                               # another user: keys are dirty now
    a3 = ops.warp_fuse_from_poses(sc, sv, tc, tv, f_idx, S, rgb=rgb, feat=feat)       # must clear again
    assert torch.equal(a1[0], a3[0]) and torch.equal(a1[1], a3[1])
    T, fim, _ = ops.cal_flow_multi(sc3, sv3, tc3, tv3, f_idx, S, return_wim=False)
    b2 = ops.warp_fuse(T, rgb=rgb[:2].contiguous(), feat=feat[:2], fim=fim)
    assert torch.equal(b1[0], b2[0]) and torch.equal(b1[1], b2[1])
    check(3, 4)


def test_warp_fuse_from_poses_other_shapes_take_the_two_call_path():
    B, K, C, S = 2, 2, 32, 48   # C = 32: not served by the fused kernel
    _, faces_idx = load_smpl_template()
    cam, verts = synth.smpl_poses(B * (K + 1), seed=7, device=DEV)
    tc, tv = cam[:B].contiguous(), verts[:B].contiguous()
    sc, sv = cam[B:].reshape(B, K, 3).contiguous(), verts[B:].reshape(B, K, -1, 3).contiguous()
    rgb, feat = synth.reference_sets(B, K, C, S, S, seed=8, device=DEV)
    r1, f1, T1, fim1 = ops.warp_fuse_from_poses(sc, sv, tc, tv, _cu(faces_idx), S, rgb=rgb, feat=feat, return_flow=True)
    T, fim, _ = ops.cal_flow_multi(sc, sv, tc, tv, _cu(faces_idx), S, return_wim=False)
    r2, f2 = ops.warp_fuse(T, rgb=rgb, feat=feat, fim=fim)
    assert torch.equal(T1, T) and torch.equal(fim1, fim) and torch.equal(r1, r2) and torch.equal(f1, f2)


def test_convlstm_sequence_grouped_is_T_steps_in_one_call():
    """The recurrence over the K references (src/convLSTM.py:131-134) as ONE C call: bit-identical to T single steps,
    T launches, no slicing / stacking on the host."""
    torch.manual_seed(11)
    G, B, T, Cin, Ch, S = 3, 2, 4, 24, 24, 50
    x = torch.randn(G, B, T, Cin, S, S, device=DEV)
    h0, c0 = torch.randn(G, B, Ch, S, S, device=DEV), torch.randn(G, B, Ch, S, S, device=DEV)
    wgt = torch.randn(G, 4 * Ch, Cin + Ch, 3, 3, device=DEV) * 0.05
    bias = torch.randn(G, 4 * Ch, device=DEV)
    wpack = ops.convlstm_gpack_weight(wgt, Cin, Ch)
    n0 = _lib.launch_count()
    h_seq, c_last = ops.convlstm_sequence_grouped(x, h0, c0, wpack, bias, Cin, Ch)
    assert _lib.launch_count() - n0 == T
    h, c = h0, c0
    for t in range(T):
        h, c = ops.convlstm_step_grouped(x[:, :, t].contiguous(), h, c, wpack, bias, Cin, Ch)
        assert torch.equal(h_seq[:, :, t], h), t
    assert torch.equal(c_last, c)
    for T1 in (1, 3):   # odd lengths end in the other ping-pong buffer
        hs, cl = ops.convlstm_sequence_grouped(x[:, :, :T1].contiguous(), h0, c0, wpack, None, Cin, Ch)
        h, c = h0, c0
        for t in range(T1):
            h, c = ops.convlstm_step_grouped(x[:, :, t].contiguous(), h, c, wpack, None, Cin, Ch)
        assert torch.equal(hs[:, :, -1], h) and torch.equal(cl, c)


@pytest.mark.parametrize("S,grouped", [(16, True), (100, False)])
def test_convlstm_cell_falls_back_when_the_tensor_core_plan_does_not_fit(S, grouped):
    """Cin = Ch = 128 passes the channel-count rule of the grouped kernel.  On a small map the planner slices the hidden
    channels over CTAs and shrinks the weight stages until the cell fits; on a 100 x 100 map (too many tiles to slice)
    the row window + weight ring exceed the SM's shared memory and the cell must take the exact-fp32 kernel instead of
    raising (the reference constructor accepts any cell)."""
    from jafpro_b200.convLSTM import ConvLSTMCell
    assert bool(ops.convlstm_grouped_supported(1, 1, 128, 128, S, S)) == grouped
    assert ops.convlstm_grouped_supported(24, 1, 24, 24, 100, 100)
    torch.manual_seed(3)
    cell = ConvLSTMCell((S, S), 128, 128, (3, 3), True).to(DEV)
    x, h, c = (torch.randn(1, 128, S, S, device=DEV) for _ in range(3))
    h2, c2 = cell(x, (h, c))
    cc = F.conv2d(torch.cat((x, h), 1).double(), cell.conv.weight.double(), cell.conv.bias.double(), padding=1)
    i, f, o, g = torch.split(cc, 128, dim=1)
    c_r = torch.sigmoid(f) * c.double() + torch.sigmoid(i) * torch.tanh(g)
    h_r = torch.sigmoid(o) * torch.tanh(c_r)
    assert float((c2.double() - c_r).abs().max()) <= 1e-4 and float((h2.double() - h_r).abs().max()) <= 1e-4


def test_drop_ins_refuse_autograd():
    """Forward-only kernels: with autograd on, a parameter / input that requires grad raises instead of silently
    returning outputs without a grad_fn."""
    from jafpro_b200.convLSTM import ConvLSTMCell
    cell = ConvLSTMCell((8, 8), 4, 4, (3, 3), True).to(DEV)
    x, h, c = (torch.randn(1, 4, 8, 8, device=DEV) for _ in range(3))
    cell(x, (h, c))  # no_grad (fixture): fine
    with torch.enable_grad():
        with pytest.raises(RuntimeError, match="forward-only"):
            cell(x, (h, c))
        g = torch.zeros(1, 1, 8, 8, 2, device=DEV)
        img = torch.randn(1, 1, 3, 8, 8, device=DEV, requires_grad=True)
        with pytest.raises(RuntimeError, match="forward-only"):
            ops.warp_fuse(g, rgb=img)
        with pytest.raises(RuntimeError, match="forward-only"):
            ops.grid_sample_border(img[:, 0], g[:, 0])
        for p in cell.parameters():
            p.requires_grad_(False)
        cell(x, (h, c))  # frozen parameters: allowed with autograd on


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_convlstm_kernels_on_a_second_device():
    """Per-device state (dynamic shared-memory attribute, SM count) is set on every device a process uses
    (nn.DataParallel, test/conv_pro_test.py:114-141)."""
    from jafpro_b200.convLSTM import ConvLSTMCell, ConvLSTMCellTC
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        torch.manual_seed(0)
        cell = ConvLSTMCell((50, 50), 24, 24, (3, 3), True).to(dev)
        x, h, c = (torch.randn(2, 24, 50, 50, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(3))
        h2, _ = cell(x, (h, c))
        tc = ConvLSTMCellTC(64, 64, torch.randn(256, 128, 3, 3, generator=torch.Generator().manual_seed(9)).to(dev), None)
        xb = torch.randn(1, 64, 64, 64, generator=torch.Generator().manual_seed(4)).to(dev).to(torch.bfloat16)
        hb, cb = torch.zeros(1, 64, 64, 64, device=dev, dtype=torch.bfloat16), torch.zeros(1, 64, 64, 64, device=dev)
        h3, _ = tc(xb, (hb, cb))
        outs.append((h2.cpu(), h3.float().cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_forward_face_index_map_has_the_extensions_contract():
    """jafpro_b200.cuda_rasterize.forward_face_index_map: the pybind signature of NR/cuda/rasterize_cuda.cpp:70-95 with
    caller-prefilled outputs filled in place, rows not flipped — against the reference kernels compiled unmodified."""
    import jafpro_b200.cuda_rasterize as rasterize_cuda
    try:
        ref = oracle.RefRaster()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libjaf_ref_raster.so not built")
    _, faces_idx = load_smpl_template()
    cam, verts = synth.smpl_poses(2, seed=11, device=DEV)
    faces = ops.project_gather(cam, verts, _cu(faces_idx))
    S, far = 128, 100.0
    for return_depth in (0, 1):
        rfim, rwim, rdepth, rfinv_map, rfaces_inv = ref.raw(faces, S, 0.1, far, return_depth=bool(return_depth))
        B, Fn = faces.shape[:2]
        fim = torch.full((B, S, S), -1, dtype=torch.int32, device=DEV)          # NR/rasterize.py:50-52
        wim = torch.zeros((B, S, S, 3), device=DEV)
        depth = torch.full((B, S, S), far, device=DEV)
        finv_map = torch.zeros((B, S, S, 3, 3) if return_depth else (1,), device=DEV)
        faces_inv = torch.zeros_like(faces)                                     # NR/rasterize.py:164
        out = rasterize_cuda.forward_face_index_map(faces.clone(), fim, wim, depth, finv_map, faces_inv, S, 0.1, far,
                                                    0, 0, return_depth)
        assert out[0] is fim and out[1] is wim and out[2] is depth and out[3] is finv_map   # same handles (:650)
        assert torch.equal(fim, rfim)
        assert np.array_equal(_bits(wim), _bits(rwim)) and np.array_equal(_bits(depth), _bits(rdepth))
        assert np.array_equal(_bits(faces_inv), _bits(rfaces_inv))
        if return_depth:
            assert np.array_equal(_bits(finv_map), _bits(rfinv_map))
    # in-place contract: background pixels keep whatever the caller put there
    fim2 = torch.full((B, S, S), -7, dtype=torch.int32, device=DEV)
    wim2 = torch.full((B, S, S, 3), 0.25, device=DEV)
    depth2 = torch.full((B, S, S), 42.0, device=DEV)
    rasterize_cuda.forward_face_index_map(faces, fim2, wim2, depth2, torch.zeros(1, device=DEV), torch.zeros_like(faces),
                                          S, 0.1, far, 0, 0, 0)
    bg = rfim == -1
    assert bool((fim2[bg] == -7).all()) and bool((wim2[bg] == 0.25).all()) and bool((depth2[bg] == 42.0).all())
    assert torch.equal(fim2[~bg], rfim[~bg])
    # CHECK_INPUT behaviour (rasterize_cuda.cpp:66-68)
    with pytest.raises(RuntimeError):
        rasterize_cuda.forward_face_index_map(faces.cpu(), fim, wim, depth, finv_map, faces_inv, S, 0.1, far, 0, 0, 0)
    with pytest.raises(RuntimeError):
        rasterize_cuda.forward_face_index_map(faces.transpose(0, 1), fim, wim, depth, finv_map, faces_inv, S, 0.1, far,
                                              0, 0, 0)


# ------------------------------------------------------------------ row F: per-reference visibility (get_vis_f2pts rule)
def test_face_visibility_matches_oracle_and_reference_fixture(golden_dir):
    d = _load(golden_dir, "vis_f2pts.npz")
    out = SMPLRenderer.get_vis_f2pts(_cu(d["f2pts"]), _cu(d["fim"]))
    assert np.array_equal(_np(out), d["out"])  # the reference function's own output, quirk included
    out1 = SMPLRenderer.get_vis_f2pts(_cu(d["f2pts"][0]), _cu(d["fim"][0]))  # unbatched form (:541-542)
    assert np.array_equal(_np(out1), d["out"][0])
    rng = np.random.default_rng(3)
    B, K, S, F = 3, 4, 40, 500
    fim_src = rng.integers(-1, F, (B, K, S, S)).astype(np.int32)
    fim_src[:, :, : S // 2] = -1
    fim_tgt = rng.integers(-1, F, (B, S, S)).astype(np.int32)
    seen, vis = ops.face_visibility(_cu(fim_src), _cu(fim_tgt), F)
    o_seen, o_vis = oracle.face_visibility(fim_src, fim_tgt, F)
    assert np.array_equal(_np(seen), o_seen) and np.array_equal(_np(vis), o_vis)
    assert 0.05 < float(o_vis.mean()) < 0.95


def test_warp_fuse_from_poses_with_per_reference_visibility():
    from jafpro_b200.fusion import reference_visibility, warp_fuse_from_poses
    _, faces_idx = load_smpl_template()
    B, K, S = 2, 3, 96
    cam, verts = synth.smpl_poses(B * (K + 1), seed=33)
    tcam, tverts = cam[:B].contiguous().to(DEV), verts[:B].contiguous().to(DEV)
    scam, sverts = cam[B:].reshape(B, K, 3).contiguous().to(DEV), verts[B:].reshape(B, K, -1, 3).contiguous().to(DEV)
    rend = SMPLRenderer(image_size=S).to(DEV)
    rgb, feat = synth.reference_sets(B, K, 64, S, S, seed=8, device=DEV)
    out_rgb, out_feat, T, fim = warp_fuse_from_poses(rend, scam, sverts, tcam, tverts, rgb=rgb, feat=feat,
                                                     per_reference_visibility=True)
    vis = reference_visibility(rend, scam, sverts, fim)
    # oracle: source rasters by the CPU restatement, the rule in numpy, then the fused op
    o_fim_src = np.stack([oracle.render_fim_wim(_np(scam[:, k]), _np(sverts[:, k]), faces_idx, S)[1] for k in range(K)], 1)
    _, o_vis = oracle.face_visibility(o_fim_src, _np(fim), faces_idx.shape[0])
    assert np.array_equal(_np(vis), o_vis)
    frac = float(o_vis[:, :, _np(fim)[0] >= 0].mean()) if False else float(o_vis.sum() / max(1, K * (_np(fim) >= 0).sum()))
    assert 0.3 < frac < 1.0  # rotated references hide part of the surface
    fb = _bf16_bits(feat.permute(0, 1, 3, 4, 2).contiguous())
    o = oracle.warp_fuse(_np(T), rgb=_np(rgb), feat=fb, feat_layout="nhwc", feat_bf16=True, vis=o_vis)
    assert float(np.abs(_np(out_rgb) - o["out_rgb"]).max()) <= 2e-6
    ok, frac_bad = _bf16_close(_bf16_bits(out_feat.permute(0, 2, 3, 1).contiguous()), o["out_feat"])
    assert ok and frac_bad < 2e-3


# ------------------------------------------------------------------ §8f rank 3: SpatioTempoCRN multi-scale warps
@pytest.mark.parametrize("C,h,w,H,W,ac", [(64, 128, 128, 256, 256, False), (512, 16, 16, 256, 256, False),
                                          (5, 13, 7, 50, 30, True), (3, 4, 4, 256, 256, False)])
def test_flow_warp_pair_matches_oracle_and_torch(C, h, w, H, W, ac):
    from jafpro_b200.crn_model import warp_level
    torch.manual_seed(C + h)
    B = 2
    prev_pool, pool = torch.randn(B, C, h, w, device=DEV), torch.randn(B, C, h, w, device=DEV)
    flow = torch.randn(B, 2, H, W, device=DEV) * 0.2
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, h, device=DEV), torch.linspace(-1, 1, w, device=DEV), indexing="ij")
    grid = torch.stack([xs, ys])[None].expand(B, 2, h, w).contiguous()
    w_prev, w_cur = warp_level(prev_pool, pool, grid, flow, align_corners=ac)
    o_prev, o_cur = oracle.flow_warp_pair(_np(prev_pool), _np(pool), _np(grid), _np(flow), ac)
    assert np.array_equal(_bits(w_prev), o_prev.view(np.int32)) and np.array_equal(_bits(w_cur), o_cur.view(np.int32))
    fs = F.interpolate(flow, (h, w), mode="nearest")  # the reference's sequence on the GPU
    t_prev = F.grid_sample(prev_pool, (grid + fs).permute(0, 2, 3, 1), padding_mode="border", align_corners=ac)
    t_cur = F.grid_sample(pool, (grid - fs).permute(0, 2, 3, 1), padding_mode="border", align_corners=ac)
    assert float((w_prev - t_prev).abs().max()) <= 1e-5 and float((w_cur - t_cur).abs().max()) <= 1e-5
    only, none = warp_level(prev_pool, None, grid, flow, align_corners=ac)
    assert none is None and torch.equal(only, w_prev)


# ------------------------------------------------------------------ §8f rank 2: texture-space assembly
def test_texture_space_assembly_matches_the_reference_loops():
    """The slicing / OR / masking / re-assembly loops of test/conv_pro_test.py:209-236 and src/networks.py:1685-1691,
    written here exactly as the reference writes them (torch slicing), against the one-launch kernels."""
    from jafpro_b200.texture import assemble_atlas, gather_parts, mask_common_area_
    torch.manual_seed(2)
    B, Kmax, ph, pw = 2, 5, 20, 12  # the reference: B=1, Kmax=5, 200x200 parts; smaller parts keep the test quick
    src_texture_im = torch.randn(B, Kmax, 3, 4 * ph, 6 * pw, device=DEV)
    src_mask_im = (torch.rand(B, Kmax, 4 * ph, 6 * pw, device=DEV) > 0.7).float()
    random_index = np.array([0, 2, 3])
    # :209-217
    ref_in = []
    for i in range(4):
        for j in range(6):
            ref_in.append([src_texture_im[:, random_index[z], :, i * ph:(i + 1) * ph, j * pw:(j + 1) * pw].squeeze(1)
                           for z in range(random_index.shape[0])])
    got = gather_parts(src_texture_im, random_index)
    assert tuple(got.shape) == (24, 3, B, 3, ph, pw)
    for p in range(24):
        assert torch.equal(got[p].flatten(0, 1), torch.cat(ref_in[p], dim=0))  # src/networks.py:1316
    # :221-236 (unused frames zeroed, OR over all Kmax as bytes, float, repeat to 3 channels, multiply per part)
    mask = src_mask_im.clone()
    for i in range(Kmax):
        if i not in list(random_index):
            mask[:, i] = mask[:, i] * 0
    common = (mask[:, 0] * 0).byte()
    for i in range(Kmax):
        common = common | mask[:, i].byte()
    common = common.float().unsqueeze(1).repeat(1, 3, 1, 1)
    accu_out = [torch.randn(B, 3, ph, pw, device=DEV) for _ in range(24)]
    ref_masked = [accu_out[i * 6 + j] * common[:, :, i * ph:(i + 1) * ph, j * pw:(j + 1) * pw] for i in range(4) for j in range(6)]
    parts = mask_common_area_(torch.stack(accu_out, 0), src_mask_im, random_index)
    for p in range(24):
        assert torch.equal(parts[p], ref_masked[p])
    # src/networks.py:1685-1691
    tex = torch.empty(B, 3, 4 * ph, 6 * pw, device=DEV)
    for i in range(4):
        for j in range(6):
            tex[:, :, i * ph:(i + 1) * ph, j * pw:(j + 1) * pw] = ref_masked[i * 6 + j]
    assert torch.equal(assemble_atlas(parts), tex) and torch.equal(assemble_atlas(ref_masked), tex)
    # full-size shapes once
    big = torch.randn(1, 5, 3, 800, 1200, device=DEV)
    g = gather_parts(big, [1, 4])
    assert torch.equal(g[7, 1, 0], big[0, 4, :, 200:400, 200:400])


# ------------------------------------------------------------------ §8f rank 4: IUV preprocessing of the data loader
def test_transfer_texture_and_compute_angle_match_the_reference_functions(golden_dir):
    """Fixture = TransferTexture (src/utils.py:369-394) and compute_angle (src/computer_angle.py:4-39) executed
    unmodified by tools/make_golden.py on oracle.inputs.iuv_preprocessing_inputs()."""
    from jafpro_b200.computer_angle import compute_angle, compute_angles
    from jafpro_b200.utils import TransferTexture
    from oracle.inputs import iuv_preprocessing_inputs
    d = _load(golden_dir, "iuv_preprocessing.npz")
    iuv, tex, im = iuv_preprocessing_inputs()
    n = iuv.shape[0]
    out_bg = TransferTexture(_cu(tex), _cu(iuv), _cu(im))         # batched, on the GPU
    assert np.array_equal(_np(out_bg), d["out_bg"])
    assert np.array_equal(TransferTexture(tex, iuv[1]), d["out_nobg"][1])  # the reference's numpy call form
    ones = TransferTexture(_cu(np.ones((800, 1200, 3), np.uint8)), _cu(iuv))   # src/data.py:108
    packed = np.unpackbits(d["ones"])[: n * 256 * 256].reshape(n, 256, 256)
    assert np.array_equal(_np(ones)[..., 0], packed) and np.array_equal(_np(ones)[..., 2], packed)
    batched_tex = TransferTexture(_cu(np.stack([tex] * n)), _cu(iuv))
    assert np.array_equal(_np(batched_tex), d["out_nobg"])
    angles = compute_angles(_cu(iuv))
    assert [float(a) for a in angles] == [float(a) for a in d["angles"]]   # bit-identical float64
    assert float(compute_angle(iuv[2])) == float(d["angles"][2])


def test_get_texture_matches_oracle_and_reference_fixture(golden_dir):
    """DensePose texture extraction (src/utils.py:232-255): scatter kernel + fp64 bilinear resize, against the numpy
    restatement (same operation order: <= 1e-13) and the reference function's own output (cv2.resize: <= 1e-12)."""
    from jafpro_b200.utils import get_texture
    from oracle.inputs import iuv_preprocessing_inputs
    d = _load(golden_dir, "get_texture.npz")
    iuv, _, im = iuv_preprocessing_inputs()
    small = ops.get_texture(_cu(im[[0, 3]]), _cu(iuv[[0, 3]]), 8, 25)          # batched, stays on the GPU
    assert tuple(small.shape) == (2, 24, 25, 25, 3) and small.dtype == torch.float64
    assert float(np.abs(_np(small) - d["small"]).max()) <= 1e-12
    for j, i in enumerate((0, 3)):
        assert float(np.abs(_np(small[j]) - oracle.get_texture(im[i], iuv[i], 8, 25)).max()) <= 1e-13
    parts = get_texture(im[1], iuv[1])                                          # the reference's numpy call form
    assert isinstance(parts, list) and len(parts) == 24 and parts[0].shape == (200, 200, 3)
    full = np.stack(parts)
    assert float(np.abs(full[:, ::7, ::7] - d["full_sub"]).max()) <= 1e-12
    assert float(np.abs(full - oracle.get_texture(im[1], iuv[1])).max()) <= 1e-13
    # a frame without any body pixel: all-zero parts
    z = ops.get_texture(_cu(im[:1]), torch.zeros(1, 256, 256, 3, dtype=torch.uint8, device=DEV))
    assert float(z.abs().max()) == 0.0


@pytest.mark.parametrize("num_inputs", [4, 3, 1])
def test_shard_loader_equals_the_reference_dataset_item(golden_dir, tmp_path, num_inputs):
    """jafpro_b200.shards.load_test_item on a packed shard == Fusion_dataset_smpl_test.__getitem__ (src/data.py:471-602)
    on the same synthetic video stored in the reference's PNG + pickle layout: every returned array has the reference's
    dtype, shape and bytes (fixture: SHA-256 digests written by tools/make_golden.py running the reference loader)."""
    import hashlib
    import json
    from jafpro_b200 import shards
    from oracle.inputs import synthetic_video
    rec = json.load(open(os.path.join(golden_dir, "dataset_item.json")))
    v = synthetic_video()
    path = str(tmp_path / "v.jafshard")
    shards.pack_video(path, v, "Synth_video_0_1", [f"frame_{t}.png" for t in range(v["img"].shape[0])])
    src_data, tgt_data, data_255, smpl_data, vid_name, names, pro_frames = shards.load_test_item(
        shards.VideoShard(path), num_inputs=num_inputs, output_mask=True, device=DEV)
    pre = f"n{num_inputs}/"
    assert [int(x) for x in pro_frames] == rec[pre + "pro_frames"]
    assert vid_name == rec[pre + "vid_name"] and names == rec[pre + "img_names"]
    flat = {"src_%d" % i: a for i, a in enumerate(src_data)}
    flat.update({"tgt_%d" % i: a for i, a in enumerate(tgt_data)})
    flat.update({"u255_%d" % i: a for i, a in enumerate(data_255)})
    flat.update({"smpl_%d" % i: a for i, a in enumerate(smpl_data)})
    assert sorted(pre + k for k in flat) == sorted(k for k in rec if k.startswith(pre) and k.split("/")[1][:3] in ("src", "tgt", "u25", "smp"))
    for k, t in flat.items():
        assert t.is_cuda
        a = np.ascontiguousarray(t.cpu().numpy())
        dtype, shape, digest = rec[pre + k]
        assert str(a.dtype) == dtype and list(a.shape) == shape, (k, a.dtype, a.shape)
        assert hashlib.sha256(a.tobytes()).hexdigest() == digest, k


def test_example_video_pipeline_runs():
    """examples/video_pipeline.py chains every drop-in on one synthetic video; it must keep running end to end."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "examples", "video_pipeline.py")], capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout.strip().endswith("ok")


def test_raster_box_margin_stress_against_reference_kernels():
    """The scatter pass tests only the pixels of [floor(min), ceil(max)] (+ the analytic sliver margin).  Guard that
    choice where it could bite: many poses and soups of sub-pixel triangles at random sub-pixel positions, bit-exact
    against the reference's own kernels (a slimmed tools/probes/raster_stress.py)."""
    try:
        ref = oracle.RefRaster()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libjaf_ref_raster.so not built (needs /root/reference at build time)")
    rend = SMPLRenderer(image_size=256).to(DEV)
    cam, verts = synth.smpl_poses(16, seed=404, device=DEV)
    for i in range(0, 16, 8):
        faces, fim, wim = rend.render_fim_wim(cam[i:i + 8].contiguous(), verts[i:i + 8].contiguous())
        rfim, rwim, _ = ref(faces, 256)
        assert torch.equal(fim, rfim) and np.array_equal(_bits(wim), _bits(rwim))
    g = torch.Generator().manual_seed(9)
    for size, nf, scale in ((128, 30000, 0.01), (64, 20000, 0.004), (200, 20000, 0.02)):
        c = torch.rand((2, nf, 1, 3), generator=g) * 2.2 - 1.1
        tri = c + (torch.rand((2, nf, 3, 3), generator=g) - 0.5) * 2 * scale
        tri[..., 2] = torch.rand((2, nf, 3), generator=g) * 3 + 0.5
        tri = tri.to(DEV).contiguous()
        a = ops.raster_fim_wim(tri, size, return_depth=True)
        b = ref(tri, size)
        assert torch.equal(a[0], b[0]), size
        assert np.array_equal(_bits(a[1]), _bits(b[1])) and np.array_equal(_bits(a[2]), _bits(b[2])), size
