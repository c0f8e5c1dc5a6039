"""CPU tests of the host-side logic: synthetic inputs, byte accounting, video sharding, and the
world-size-2 process-group path on gloo."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import torch

import oracle
from jafpro_b200 import dist as jdist
from jafpro_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_survey_figures():
    assert synth.warp_fuse_bytes(4, 256, 256, 64) == 49_283_072
    assert synth.warp_fuse_bytes(4, 256, 256, 0) == 7_340_032
    assert synth.warp_fuse_bytes(8, 512, 512, 64) == 356_515_840


def test_identity_grid_is_a_fixed_point_of_the_warp():
    H = W = 16
    g = synth.identity_grid(H, W)[None].numpy()
    src = np.random.default_rng(0).normal(size=(1, 2, H, W)).astype(np.float32)
    out = oracle.grid_sample_border(src, g, align_corners=False)
    assert np.abs(out - src).max() < 1e-6


def test_dense_flows_stay_within_displacement_bound():
    g = synth.dense_flows(2, 3, 32, 32, seed=1, max_disp_px=4.0)
    d = (g - synth.identity_grid(32, 32)[None, None]).abs()
    assert float(d[..., 0].max()) <= 2 * 4.0 / 32 + 1e-6
    assert tuple(g.shape) == (2, 3, 32, 32, 2)


def test_synthetic_pose_foreground_fraction():
    cam, verts = synth.smpl_poses(2, seed=3)
    from jafpro_b200.nmr import load_smpl_template
    _, faces = load_smpl_template()
    _, fim, _ = oracle.render_fim_wim(cam.numpy(), verts.numpy(), faces, 64)
    frac = (fim != -1).mean()
    assert 0.05 < frac < 0.3


def test_shard_videos_partitions_exactly():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in jdist.shard_videos(64, r, world))
        assert seen == list(range(64))
    assert jdist.shard_videos(8, 1, 2) == [1, 3, 5, 7]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo_sharding_and_reduction(tmp_path):
    """Two ranks shard 6 videos, 'process' them, and reduce (max time, sum units) over gloo."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import torch
        from jafpro_b200 import dist as jd
        rank, world, _ = jd.init("gloo")
        vids = jd.shard_videos(6, rank, world)
        jd.barrier()
        ms, units = jd.reduce_max_sum(10.0 * (rank + 1), 30.0 * len(vids), device="cpu")
        got = jd.gather_results(torch.tensor(vids))
        if rank == 0:
            print("RESULT", ms, units, [g.tolist() for g in got])
    """))
    port = _free_port()
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT")][0]
    assert "20.0 180.0 [[0, 2, 4], [1, 3, 5]]" in line


def test_compute_angle_host_formula_matches_reference_fixture(golden_dir):
    """The float64 tail of compute_angle (src/computer_angle.py:20-39) restated on top of per-part (count, sum x)
    statistics — the GPU kernel only supplies those integers.  Fixture: the reference function's own output."""
    import numpy as np
    from jafpro_b200.computer_angle import _angle_from_stats
    from oracle.inputs import iuv_preprocessing_inputs
    iuv, _, _ = iuv_preprocessing_inputs()
    d = np.load(os.path.join(golden_dir, "iuv_preprocessing.npz"))
    for i in range(iuv.shape[0]):
        part = iuv[i, :, :, 0]
        counts = np.bincount(part.ravel(), minlength=32)[:32]
        xs = np.broadcast_to(np.arange(part.shape[1])[None], part.shape)
        sumx = np.array([xs[part == p].sum() for p in range(32)])
        assert float(_angle_from_stats(counts, sumx)) == float(d["angles"][i])


def test_bind_to_gpu_cpus_is_a_no_op_without_nvml():
    """The per-rank CPU binding used by the host-buffer path must never raise: without a GPU (or NVML) it reports
    False and leaves the affinity alone."""
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert jdist.bind_to_gpu_cpus(0) is False
    assert os.sched_getaffinity(0) == before or torch.cuda.is_available()


def test_hard_and_perm_flows_cover_the_frame():
    g = synth.hard_flows(2, 3, 64, 64, seed=2, block=16, max_disp_px=16.0)
    assert tuple(g.shape) == (2, 3, 64, 64, 2) and bool(torch.isfinite(g).all())
    d = (g - synth.identity_grid(64, 64)[None, None]).abs() * 32          # pixels
    assert 2.0 < float(d.mean()) < 40.0                                    # far from the identity, bounded
    assert float((g.abs() > 1).any(-1).float().mean()) < 0.05              # (almost) every sample lands in the frame
    # piecewise-affine: inside a cell the flow is exactly affine along x (second difference ~ 0)
    row = g[0, 0, 5, :16, 0]
    assert float((row[2:] - 2 * row[1:-1] + row[:-2]).abs().max()) < 1e-5
    p = synth.perm_flows(1, 2, 16, 16, seed=3)
    px = ((p[0, 0, ..., 0] * 16 + 16 - 1) / 2).round().long() + 16 * ((p[0, 0, ..., 1] * 16 + 16 - 1) / 2).round().long()
    assert sorted(px.flatten().tolist()) == list(range(256))               # a permutation of the source pixels


def test_bench_visible_pixel_bytes_never_exceed_dense_bytes():
    import bench
    K, S, C, B = 4, 256, 64, 10
    dense = bench.wf_alg_bytes_dense(K, S, C) * B
    full = bench.wf_alg_bytes_visible(K, S, C, B, B * S * S)
    some = bench.wf_alg_bytes_visible(K, S, C, B, B * S * S // 8)
    assert full == dense + B * S * S * 4        # every pixel visible: the dense bytes + the face-index map
    assert some < dense / 3


def test_drop_ins_refuse_autograd_before_touching_the_gpu():
    """Forward-only contract (CPU part): a tensor / parameter that requires grad raises while autograd is enabled."""
    import pytest
    from jafpro_b200 import ops
    from jafpro_b200.convLSTM import ConvLSTMCell
    cell = ConvLSTMCell((4, 4), 2, 4, (3, 3), True)
    x, h, c = torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4, 4)
    with pytest.raises(RuntimeError, match="forward-only"):
        cell(x, (h, c))
    with pytest.raises(RuntimeError, match="forward-only"):
        ops.warp_fuse(torch.zeros(1, 1, 4, 4, 2, requires_grad=True), rgb=torch.zeros(1, 1, 3, 4, 4))
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA tensor"):   # without grad: the usual CPU-tensor error
        ops.warp_fuse(torch.zeros(1, 1, 4, 4, 2, requires_grad=True), rgb=torch.zeros(1, 1, 3, 4, 4))


def test_patched_reference_renderer_has_all_it_needs(monkeypatch):
    """INTEGRATION.md §3 binds jafpro_b200.nmr.SMPLRenderer's methods onto the REFERENCE class, whose instances only
    carry what src/nmr.py:104-177 sets (eye, faces, image_size, proj_func ...; no _eye_z).  The bound methods must work
    with exactly that attribute set."""
    import jafpro_b200.nmr as jnmr
    from jafpro_b200 import ops

    class RefRendererStub:                       # attribute set of the reference SMPLRenderer (src/nmr.py:104-177)
        def __init__(self):
            self.image_size = 256
            self.faces = torch.zeros(5, 3, dtype=torch.int32)
            self.eye = [0, 0, -(1. / np.tan(np.radians(30)) + 1)]
            self.proj_func = jnmr.orthographic_proj_withz_idrot

    RefRendererStub.render_fim_wim = lambda self, cam, v, faces=None: jnmr.SMPLRenderer.render_fim_wim(self, cam, v, faces)
    RefRendererStub.cal_bc_transform = jnmr.SMPLRenderer.cal_bc_transform
    seen = {}

    def fake_render(cam, verts, faces_idx, image_size, eye_z=None, **kw):
        seen.update(eye_z=eye_z, image_size=image_size, faces_dtype=faces_idx.dtype)
        return None, None, None
    monkeypatch.setattr(ops, "render_fim_wim", fake_render)
    r = RefRendererStub()
    r.render_fim_wim(torch.zeros(1, 3), torch.zeros(1, 7, 3))
    assert seen["eye_z"] == float(np.float32(-(1. / np.tan(np.radians(30)) + 1)))
    assert seen["image_size"] == 256 and seen["faces_dtype"] == torch.int32


def test_video_shard_round_trip_and_directory_converter(tmp_path):
    """The packed shard format (jafpro_b200/shards.py): raw arrays come back bit for bit, payloads are 256-byte aligned,
    and converting the reference's on-disk layout (PNG files + pose_shape.pkl, src/utils.py:26-58) gives the same shard
    content as packing the arrays directly."""
    import pickle
    import pytest
    from jafpro_b200 import shards
    from oracle.inputs import synthetic_video
    v = synthetic_video(T=3, seed=2)
    names = [f"frame_{t}.png" for t in range(3)]
    path = str(tmp_path / "v.jafshard")
    shards.pack_video(path, v, "Synth_video_0_1", names)
    sh = shards.VideoShard(path)
    assert sh.num_frames == 3 and sh.meta["vid_name"] == "Synth_video_0_1" and sh.meta["img_names"] == names
    for k in list(shards.VIDEO_ARRAYS) + list(shards.SMPL_ARRAYS):
        assert np.array_equal(sh[k], v[k]) and sh[k].dtype == v[k].dtype
        assert (sh._base + sh._table[k]["offset"]) % shards.ALIGN == 0
    with pytest.raises(ValueError):
        bad = tmp_path / "bad"
        bad.write_bytes(b"not a shard at all")
        shards.VideoShard(str(bad))
    cv2 = pytest.importorskip("cv2")
    vid, msk = tmp_path / "data" / "Synth_video_0_1", tmp_path / "mask" / "Synth_video_0_1"
    vid.mkdir(parents=True)
    msk.mkdir(parents=True)
    for t in range(3):
        cv2.imwrite(str(vid / f"frame_{t}.png"), v["img"][t])
        cv2.imwrite(str(vid / f"frame_{t}_IUV.png"), v["iuv"][t])
        cv2.imwrite(str(vid / f"frame_{t}_text.png"), v["text"][t])
        cv2.imwrite(str(vid / f"frame_{t}_mask.png"), v["text_mask"][t])
        cv2.imwrite(str(msk / f"frame_{t}_mask.png"), v["real_mask"][t])
    with open(tmp_path / "pose_shape.pkl", "wb") as fh:
        pickle.dump({k: v[k] for k in shards.SMPL_ARRAYS}, fh)
    path2 = str(tmp_path / "v2.jafshard")
    shards.pack_video_dir(path2, str(vid), str(tmp_path / "pose_shape.pkl"), str(msk))
    sh2 = shards.VideoShard(path2)
    for k in list(shards.VIDEO_ARRAYS) + list(shards.SMPL_ARRAYS):
        assert np.array_equal(sh2[k], v[k]), k
    assert sh2.meta["img_names"] == names


def test_reference_frame_selection_rule():
    """select_reference_frames restates src/data.py:505-527 (argmax / argsort thirds / argmin, clipped to 0..30)."""
    from jafpro_b200.shards import select_reference_frames
    angle = np.array([10.0, -40.0, 65.0, 3.0, -5.0, 30.0, -60.0])
    pro, fr = select_reference_frames(angle, 4)
    order = np.argsort(angle)
    assert list(pro) == [2, order[7 // 3], order[14 // 3], 6] and list(fr) == list(pro)
    assert list(select_reference_frames(angle, 1)[0]) == [3]
    assert list(select_reference_frames(angle, 3)[0]) == [2, order[3], 6]
    assert len(select_reference_frames(angle, 5)[0]) == 5
    big = np.arange(40, dtype=np.float64)
    assert list(select_reference_frames(big, 1)[1]) == [0] and list(select_reference_frames(-big, 1)[1]) == [0]
    assert list(select_reference_frames(big, 4)[1])[0] == 30     # np.clip(frames, 0, 30), src/data.py:527


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU arm): one JSON line with the contract keys and the SAME `config` as our arm
    would print for that workload (the driver compares them).  Run on the small latency workload."""
    import json
    import types
    import bench
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1_latency",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    args = types.SimpleNamespace(workload="c1_latency", flow="dense")
    assert line["config"] == bench.config_for(args, 1)
    # and the headline workload's config is built by the same function for both arms
    a2 = types.SimpleNamespace(workload="dancevideo_256_k4_c64", flow="dense")
    cfg = bench.config_for(a2, 1)
    assert cfg["frames_per_step_per_gpu"] == 240 and cfg["K"] == 4 and cfg["C"] == 64 and "knobs" in cfg
