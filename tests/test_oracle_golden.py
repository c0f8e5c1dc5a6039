"""CPU tests: the oracle (oracle/jaf_oracle.c) against every golden vector the
reference's own tests hold for this path and against fixtures produced by the
reference's own Python modules (tools/make_golden.py).  No GPU, no /root/reference."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_teapot_silhouette_matches_blender_golden(golden_dir):
    """NR tests/test_rasterize_silhouettes.py:16-35 / test_rasterize_depth.py:15-35:
    coverage at 256x256, no anti-aliasing, must equal teapot_blender.png exactly."""
    d = _load(golden_dir, "teapot_faces.npz")
    faces = d["faces"][None]
    shape = tuple(d["shape"])
    ref = np.unpackbits(d["silhouette_bits"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
    fim, wim, depth = oracle.raster_fim_wim(faces, 256, near=0.1, far=100.0, flip_rows=True, return_depth=True)
    sil = fim[0] != -1
    assert ref.sum() == 7580
    assert np.array_equal(sil, ref)
    # test_rasterize_depth.py:15-35: depth != depth.max() is the same silhouette
    assert np.array_equal(depth[0] != depth[0].max(), ref)
    # weights are normalised barycentrics on covered pixels, zero elsewhere
    assert np.allclose(wim[0][sil].sum(-1), 1.0, atol=1e-5)
    assert np.all(wim[0][~sil] == 0)


def test_teapot_depth_matches_golden_png(golden_dir):
    """NR tests/test_rasterize_depth.py:37-54 (atol 1e-2 on the normalised depth)."""
    d = _load(golden_dir, "teapot_faces.npz")
    _, _, depth = oracle.raster_fim_wim(d["faces"][None], 256, flip_rows=True, return_depth=True)
    image = depth[0].copy()
    image[image == image.max()] = image.min()
    image = (image - image.min()) / (image.max() - image.min())
    ref = d["depth_png"].astype(np.float32) / 255.0
    if ref.ndim == 3:
        ref = ref[..., 0]
    assert np.allclose(image, ref, atol=1e-2)


def test_minibatch_of_zero_faces_is_empty(golden_dir):
    """NR tests/utils.py:11-27 puts the sample at index 2 of a batch of all-zero meshes:
    degenerate (0/0) faces must never win a pixel."""
    d = _load(golden_dir, "teapot_faces.npz")
    faces = np.zeros((2, 64, 3, 3), np.float32)
    faces[1] = d["faces"][:64]
    fim, wim = oracle.raster_fim_wim(faces, 32)
    assert np.all(fim[0] == -1) and np.all(wim[0] == 0)


def test_look_at_identity_for_smpl_eye(golden_dir):
    """NR/look_at.py:6-62 with the SMPLRenderer eye (src/nmr.py:177) is exactly v - eye."""
    d = _load(golden_dir, "look_at.npz")
    v = d["verts"]
    exp = v.copy()
    exp[..., 2] = v[..., 2] - np.float32(oracle.EYE_Z)
    assert np.array_equal(exp, d["smpl_eye_out"])
    # the reference's own known answers (tests/test_look_at.py:9-25) travel with the fixture
    assert np.allclose(d["known_out"][1], [1, 0, 10])


def test_project_gather_matches_reference_render_faces(golden_dir):
    """src/nmr.py:263-276 executed by the reference itself (rasteriser stubbed)."""
    d = _load(golden_dir, "render_faces.npz")
    tmpl = np.load(os.path.join(os.path.dirname(golden_dir), "..", "jafpro_b200", "data", "smpl_template.npz"))
    faces_idx = tmpl["faces"].astype(np.int32)
    out = oracle.project_gather(d["cam"], d["verts"], faces_idx)
    assert np.array_equal(out[:, d["face_subset"]], d["faces_xyz_subset"])


def test_flow_compose_matches_reference_cal_bc_transform(golden_dir):
    """src/nmr.py:617-659 executed by the reference itself."""
    d = _load(golden_dir, "bc_transform.npz")
    T = oracle.flow_compose(d["src"], d["fim"], d["wim"])
    assert np.array_equal(T, d["T"])
    assert np.all(T[d["fim"] == -1] == -2.0)


def test_convlstm_cell_matches_reference(golden_dir):
    """src/convLSTM.py:41-56 executed by the reference itself (fp32 CPU)."""
    d = _load(golden_dir, "convlstm.npz")
    h2, c2 = oracle.convlstm_step(d["x"], d["h"], d["c"], d["weight"], d["bias"])
    assert np.abs(h2 - d["h_out"]).max() <= 2e-6
    assert np.abs(c2 - d["c_out"]).max() <= 2e-6
    # 3-step ConvLSTM.forward (:102-147) from zero state (:58-63)
    B, K = d["seq_x"].shape[:2]
    h = np.zeros_like(d["seq_h"])
    c = np.zeros_like(d["seq_c"])
    for t in range(K):
        h, c = oracle.convlstm_step(d["seq_x"][:, t], h, c, d["seq_weight"], d["seq_bias"])
        assert np.abs(h - d["seq_out"][:, t]).max() <= 5e-6
    assert np.abs(c - d["seq_c"]).max() <= 5e-6


def test_softmax_fuse_matches_reference_downsampler_mask(golden_dir):
    """src/networks.py:1259-1286 executed by the reference itself."""
    d = _load(golden_dir, "softmax_fuse.npz")
    out = oracle.softmax_fuse(d["feat"], d["logits"])
    assert np.abs(out - d["out"]).max() <= 1e-6


def test_mask_blend_matches_reference_propagation_net(golden_dir):
    """src/flow_net.py:87-99 executed by the reference itself."""
    d = _load(golden_dir, "mask_blend.npz")
    masked, pred = oracle.mask_blend(d["tsf"], d["mask"], d["fake"], d["weight"])
    assert np.array_equal(masked, d["tsf"] * d["mask"])
    assert np.abs(pred - d["pred"]).max() <= 1e-6


@pytest.mark.parametrize("align_corners", [False, True])
def test_grid_sample_matches_installed_torch(align_corners):
    """src/cal_flow.py:37-39; third-party arithmetic pinned against the installed torch CPU kernel.
    Includes the -2 background sentinel, exact borders, far out-of-range values and NaN-free ragged sizes."""
    rng = np.random.default_rng(0)
    N, C, Hs, Ws, H, W = 2, 3, 13, 17, 11, 9
    src = rng.normal(size=(N, C, Hs, Ws)).astype(np.float32)
    grid = rng.uniform(-1.3, 1.3, size=(N, H, W, 2)).astype(np.float32)
    grid[0, 0, :3] = -2.0
    grid[0, 1, 0] = (-1.0, 1.0)
    grid[0, 1, 1] = (1.0, -1.0)
    grid[1, 2, 2] = (37.0, -55.0)
    ref = F.grid_sample(torch.from_numpy(src), torch.from_numpy(grid), mode="bilinear",
                        padding_mode="border", align_corners=align_corners).numpy()
    out = oracle.grid_sample_border(src, grid, align_corners)
    assert np.abs(out - ref).max() <= 1e-5  # torch CPU unnormalises in a different operation order than the CUDA header formula
    # background sentinel clamps to source pixel (0, 0)
    assert np.allclose(out[0, :, 0, 0], src[0, :, 0, 0], atol=1e-6)


def _torch_warp_fuse(grid, rgb, logits, vis, mask, align_corners):
    """Row F written with the reference's torch primitives only."""
    B, K = grid.shape[:2]
    g = torch.from_numpy(grid)
    r = torch.from_numpy(rgb)
    warped = torch.stack([F.grid_sample(r[:, k], g[:, k], padding_mode="border", align_corners=align_corners)
                          for k in range(K)], 1)                      # cal_flow.py:37-39
    a = torch.softmax(torch.from_numpy(logits), dim=1)                 # networks.py:1230-1244
    a = a * torch.from_numpy(vis)
    fused = (warped * a[:, :, None]).sum(1)                            # networks.py:1276-1286
    return (fused * torch.from_numpy(mask)).numpy(), warped.numpy()    # flow_net.py:91


@pytest.mark.parametrize("K", [1, 3, 4])
def test_warp_fuse_matches_torch_composition(K):
    rng = np.random.default_rng(K)
    B, H, W = 2, 12, 10
    rgb = rng.normal(size=(B, K, 3, H, W)).astype(np.float32)
    grid = rng.uniform(-1.1, 1.1, size=(B, K, H, W, 2)).astype(np.float32)
    logits = rng.normal(size=(B, K, H, W)).astype(np.float32)
    vis = (rng.random((B, K, H, W)) > 0.3).astype(np.float32)
    mask = (rng.random((B, 1, H, W)) > 0.2).astype(np.float32)
    ref, warped = _torch_warp_fuse(grid, rgb, logits, vis, mask, False)
    out = oracle.warp_fuse(grid, rgb=rgb, logits=logits, vis=vis, tgt_mask=mask, return_warped=True)
    assert np.abs(out["warped_rgb"] - warped).max() <= 1e-5
    assert np.abs(out["out_rgb"] - ref).max() <= 1e-5


def test_warp_fuse_k1_reduces_to_warp_image_times_mask():
    """K=1, no logits => exactly warp_image(src, T) * mask."""
    rng = np.random.default_rng(7)
    B, H, W = 2, 16, 16
    rgb = rng.normal(size=(B, 1, 3, H, W)).astype(np.float32)
    grid = rng.uniform(-1.2, 1.2, size=(B, 1, H, W, 2)).astype(np.float32)
    mask = (rng.random((B, 3, H, W)) > 0.5).astype(np.float32)
    out = oracle.warp_fuse(grid, rgb=rgb, tgt_mask=mask)
    ws = oracle.grid_sample_border(rgb[:, 0], grid[:, 0])
    assert np.array_equal(out["out_rgb"], ws * mask)


def test_warp_fuse_feature_layouts_and_bf16_agree():
    rng = np.random.default_rng(3)
    B, K, C, H, W = 2, 3, 8, 9, 11
    feat = rng.normal(size=(B, K, C, H, W)).astype(np.float32)
    grid = rng.uniform(-1.1, 1.1, size=(B, K, H, W, 2)).astype(np.float32)
    logits = rng.normal(size=(B, K, H, W)).astype(np.float32)
    planar = oracle.warp_fuse(grid, feat=feat, logits=logits)["out_feat"]
    nhwc = oracle.warp_fuse(grid, feat=feat.transpose(0, 1, 3, 4, 2), feat_layout="nhwc", logits=logits)["out_feat"]
    assert np.array_equal(planar, nhwc.transpose(0, 3, 1, 2))
    hb = oracle.f32_to_bf16_bits(feat)
    fb = oracle.bf16_bits_to_f32(hb)
    exact = oracle.warp_fuse(grid, feat=fb, logits=logits)["out_feat"]
    got = oracle.warp_fuse(grid, feat=hb, feat_bf16=True, logits=logits)["out_feat"]
    assert np.array_equal(got, oracle.f32_to_bf16_bits(exact))
    # bf16 conversion agrees with torch's
    t = torch.from_numpy(feat).to(torch.bfloat16)
    assert np.array_equal(hb, t.view(torch.int16).numpy().view(np.uint16))


def test_flow_self_transfer_reproduces_pixel_centres():
    """SURVEY 'facts verified': warping a pose onto itself gives pixel-centre NDC
    (align_corners=False convention) on covered pixels, -2 elsewhere."""
    tmpl = np.load(os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "..", "jafpro_b200", "data",
                                "smpl_template.npz"))
    v = tmpl["verts"][None].astype(np.float32)
    faces_idx = tmpl["faces"].astype(np.int32)
    cam = np.array([[0.85, 0.0, 0.25]], np.float32)
    S = 64
    T, fim, wim = oracle.cal_flow(cam, v, cam, v, faces_idx, S)
    fg = fim[0] != -1
    assert 0.05 < fg.mean() < 0.4
    ys, xs = np.nonzero(fg)
    cx = (2 * xs + 1 - S) / S
    cy = (2 * ys + 1 - S) / S
    assert np.abs(T[0][fg][:, 0] - cx).max() < 2e-3
    assert np.abs(T[0][fg][:, 1] - cy).max() < 2e-3
    assert np.all(T[0][~fg] == -2.0)


def test_texture_warp_matches_reference_function(golden_dir):
    """test/conv_pro_test.py:41-74 executed from the reference script itself (tools/make_golden.py)."""
    d = _load(golden_dir, "texture_warp.npz")
    out = oracle.texture_warp(d["tex"], d["iuv"], align_corners=False)
    assert np.abs(out - d["out"]).max() <= 1e-6
    assert np.all(out[:, d["iuv"][..., 0] == 0] == 0)
    # batched call == per-frame calls
    both = oracle.texture_warp(d["tex"], np.stack([d["iuv"], d["iuv"][::-1].copy()]))
    assert np.array_equal(both[0], out)


def test_get_vis_f2pts_restatement_matches_the_reference_function(golden_dir):
    """Fixture = SMPLRenderer.get_vis_f2pts (src/nmr.py:507-546) executed by tools/make_golden.py, one item with and
    one without background pixels; the per-reference visibility rule of row F is derived from it."""
    d = np.load(os.path.join(golden_dir, "vis_f2pts.npz"))
    out = oracle.get_vis_f2pts(d["f2pts"], d["fim"])
    assert np.array_equal(out, d["out"])
    F = d["f2pts"].shape[1]
    seen, vis = oracle.face_visibility(d["fim"][:, None], d["fim"], F)
    # item 0 (has background): invisible faces <=> -2 rows of the reference output
    assert np.array_equal(seen[0, 0] == 0, d["out"][0, :, 0, 0] == -2)
    # every non-background target pixel of a pose is visible from the same pose
    assert np.array_equal(vis[:, 0] == 1, d["fim"] >= 0)


@pytest.mark.parametrize("h,w,H,W", [(8, 8, 64, 64), (13, 7, 50, 30), (5, 9, 17, 40)])
def test_flow_warp_pair_oracle_matches_torch_composition(h, w, H, W):
    """The oracle of §8f rank 3 against the reference's own op sequence (src/crn_model.py:457-466) in torch on CPU:
    the nearest-neighbour index must agree exactly (odd ratios included), the warp within grid_sample's 1e-5."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(h * w)
    B, C = 2, 3
    a, b = rng.normal(size=(2, B, C, h, w)).astype(np.float32)
    flow = (rng.normal(size=(B, 2, H, W)) * 0.3).astype(np.float32)
    ys, xs = np.meshgrid(np.linspace(-1, 1, h, dtype=np.float32), np.linspace(-1, 1, w, dtype=np.float32), indexing="ij")
    grid = np.broadcast_to(np.stack([xs, ys])[None], (B, 2, h, w)).copy()
    o_f, o_b = oracle.flow_warp_pair(a, b, grid, flow)
    fs = F.interpolate(torch.from_numpy(flow), (h, w), mode="nearest")
    t_f = F.grid_sample(torch.from_numpy(a), (torch.from_numpy(grid) + fs).permute(0, 2, 3, 1), padding_mode="border",
                        align_corners=False)
    t_b = F.grid_sample(torch.from_numpy(b), (torch.from_numpy(grid) - fs).permute(0, 2, 3, 1), padding_mode="border",
                        align_corners=False)
    assert float(np.abs(o_f - t_f.numpy()).max()) <= 1e-5 and float(np.abs(o_b - t_b.numpy()).max()) <= 1e-5


def test_get_texture_oracle_matches_the_reference_function(golden_dir):
    """Fixture = get_texture (src/utils.py:232-255) executed by tools/make_golden.py with the container's OpenCV.  The
    restatement rounds the resize coefficients exactly; OpenCV builds differ in theirs, hence 1e-12 instead of bits."""
    from oracle.inputs import iuv_preprocessing_inputs
    d = np.load(os.path.join(golden_dir, "get_texture.npz"))
    iuv, _, im = iuv_preprocessing_inputs()
    for j, i in enumerate((0, 3)):
        assert float(np.abs(oracle.get_texture(im[i], iuv[i], 8, 25) - d["small"][j]).max()) <= 1e-12
    full = oracle.get_texture(im[1], iuv[1])
    assert float(np.abs(full[:, ::7, ::7] - d["full_sub"]).max()) <= 1e-12
    assert float(np.abs(full.sum(axis=(1, 2, 3)) - d["full_sum"]).max()) <= 1e-8
    assert int((d["full_sum"] > 0).sum()) == 12 and float(full.min()) >= 0.0 and float(full.max()) <= 1.0
