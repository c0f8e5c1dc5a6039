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
