#!/usr/bin/env python
"""Secondary measurements (BASELINE configs 1, 3, 4 and the stand-alone elementwise kernels), one JSON
line each: CUDA-event time, algorithmic bytes / flops per launch and the fraction of the
measured peak.  GPU only.  The headline (config 2) lives in bench.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from jafpro_b200 import _lib, ops, synth  # noqa: E402
from jafpro_b200.nmr import load_smpl_template  # noqa: E402

DEV = "cuda"


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ev[0].elapsed_time(ev[-1]) / n, ts[len(ts) // 2], ts[0]


def main():
    hbm, tf_burst, tf_sus, src = peaks()
    out = []
    # ---- config 1: one 256x256 RGB reference, batch 1 (latency-bound)
    rgb, _ = synth.reference_sets(1, 1, 0, 256, 256, seed=1, device=DEV)
    grid = synth.dense_flows(1, 1, 256, 256, seed=1, device=DEV)
    src_img, g1 = rgb[:, 0].contiguous(), grid[:, 0].contiguous()
    avg, med, best = timeit(lambda: ops.grid_sample_border(src_img, g1), n=200)
    tavg, tmed, _ = timeit(lambda: F.grid_sample(src_img, g1, padding_mode="border", align_corners=False), n=200)
    out.append({"config": "C1 warp 1x3x256x256 (jaf_warp_image)", "us_avg": round(avg * 1e3, 2), "us_median": round(med * 1e3, 2),
                "torch_cuda_grid_sample_us_median": round(tmed * 1e3, 2), "bytes": 2359296 - 262144,
                "note": "launch-latency bound: 0.36 us of HBM time at peak"})
    # ---- config 3: transfer-flow construction, 30 frames, 6890 v / 13776 f
    _, fidx = load_smpl_template()
    f_idx = torch.from_numpy(fidx).to(DEV)
    cam, verts = synth.smpl_poses(60, seed=3, device=DEV)
    sc, sv, tc, tv = cam[:30].contiguous(), verts[:30].contiguous(), cam[30:].contiguous(), verts[30:].contiguous()
    n0 = _lib.launch_count()
    avg, med, best = timeit(lambda: ops.cal_flow(sc, sv, tc, tv, f_idx, 256), n=50)
    launches = (_lib.launch_count() - n0) // 55
    T, fim, wim = ops.cal_flow(sc, sv, tc, tv, f_idx, 256, return_maps=True)
    by = 30 * (2 * 82680 + 165312 + 524288)  # both poses' vertices + face indices + flow out
    out.append({"config": "C3 cal_flow fused (project + raster + compose), 30 frames 256^2", "ms_avg": round(avg, 4),
                "ms_median": round(med, 4), "frames_per_s": round(30 / avg * 1e3, 1), "kernels_per_call": launches,
                "compulsory_GBps": round(by / avg / 1e6, 1), "foreground_frac": round(float((fim != -1).float().mean()), 4),
                "brute_force_tests_per_s_equiv": round(30 * 902823936 / avg * 1e3, 0)})
    avg2, med2, _ = timeit(lambda: ops.render_fim_wim(tc, tv, f_idx, 256), n=50)
    by2 = 30 * (82680 + 262144 + 786432 + 495936)
    out.append({"config": "C3 render_fim_wim (faces + fim + wim materialised), 30 frames", "ms_avg": round(avg2, 4),
                "frames_per_s": round(30 / avg2 * 1e3, 1), "GBps": round(by2 / avg2 / 1e6, 1)})
    src_pts = torch.randn(30, 13776, 3, 2, device=DEV)
    avg3, med3, _ = timeit(lambda: ops.flow_compose(src_pts, fim, wim), n=50)
    out.append({"config": "a9 flow_compose stand-alone, 30 frames", "ms_avg": round(avg3, 4),
                "GBps": round(30 * 1903488 / avg3 / 1e6, 1), "frac_of_hbm": round(30 * 1903488 / avg3 / 1e6 / hbm, 3)})
    # ---- config 4: ConvLSTM step, B=16, 256+256 -> 1024, 64x64, K=4 sequential steps
    B, Cin, Ch, H, W = 16, 256, 256, 64, 64
    torch.manual_seed(0)
    wgt = torch.randn(4 * Ch, Cin + Ch, 3, 3, device=DEV) * 0.01
    bias = torch.zeros(4 * Ch, device=DEV)
    wpack = ops.convlstm_pack_weight(wgt, Cin, Ch)
    xs = [torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16) for _ in range(4)]
    h = torch.zeros(B, H, W, Ch, device=DEV, dtype=torch.bfloat16)
    c = torch.zeros(B, H, W, Ch, device=DEV)
    flop = 2 * (B * H * W) * (4 * Ch) * (9 * (Cin + Ch))

    def lstm4():
        hh, cc = h, c
        for t in range(4):
            hh, cc = ops.convlstm_step_tc(xs[t], hh, cc, wpack, bias, Cin, Ch)
        return hh

    avg, med, best = timeit(lambda: ops.convlstm_step_tc(xs[0], h, c, wpack, bias, Cin, Ch), n=30)
    avg4, med4, _ = timeit(lstm4, n=10)
    out.append({"config": "C4 ConvLSTM step tcgen05, B=16 256+256->1024 3x3 64x64", "ms_avg": round(avg, 4), "ms_best": round(best, 4),
                "TFLOPs_avg": round(flop / avg / 1e9, 1), "TFLOPs_best": round(flop / best / 1e9, 1),
                "frac_of_bf16_burst": round(flop / best / 1e9 / tf_burst, 3), "frac_of_bf16_sustained": round(flop / avg / 1e9 / tf_sus, 3),
                "K4_steps_ms": round(avg4, 4), "peak_source": src, "flop_per_step": flop})
    # cuDNN bf16 conv alone (no gates) for context: the library baseline of the same contraction
    xin = torch.cat((xs[0], h), -1).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    wb = wgt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    cavg, cmed, cbest = timeit(lambda: F.conv2d(xin, wb, None, padding=1), n=20)
    out.append({"config": "C4 context: cuDNN bf16 conv2d only (no cat, no gates)", "ms_avg": round(cavg, 4), "TFLOPs_avg": round(flop / cavg / 1e9, 1)})
    # reference-sized fp32 cell (CUDA-core fused kernel)
    x32, h32, c32 = (torch.randn(24, n, 100, 100, device=DEV) for n in (24, 24, 24))
    w32 = torch.randn(96, 48, 3, 3, device=DEV) * 0.05
    b32 = torch.zeros(96, device=DEV)
    favg, _, _ = timeit(lambda: ops.convlstm_step(x32, h32, c32, w32, b32), n=30)
    tavg, _, _ = timeit(lambda: F.conv2d(torch.cat((x32, h32), 1), w32, b32, padding=1), n=30)
    out.append({"config": "a13 fp32 fused cell, reference size Ch=24 @100^2 x24 parts", "ms_avg": round(favg, 4),
                "torch_conv_only_ms": round(tavg, 4)})
    # ---- elementwise
    t = torch.randn(240, 3, 256, 256, device=DEV)
    m = torch.ones(240, 1, 256, 256, device=DEV)
    fk = torch.randn_like(t)
    wc = torch.rand(240, 1, 256, 256, device=DEV)
    eavg, _, _ = timeit(lambda: ops.mask_blend(t, m, fk, wc), n=30)
    eb = 240 * 65536 * 4 * (3 + 1 + 3 + 1 + 3 + 3)
    out.append({"config": "a11 mask+blend 240x3x256^2", "ms_avg": round(eavg, 4), "GBps": round(eb / eavg / 1e6, 1), "frac_of_hbm": round(eb / eavg / 1e6 / hbm, 3)})
    fc = torch.randn(64, 3 * 24, 100, 100, device=DEV)
    lg = torch.randn(64, 3, 100, 100, device=DEV)
    savg, _, _ = timeit(lambda: ops.softmax_fuse(fc, lg), n=30)
    sb = 64 * 10000 * 4 * (72 + 3 + 24)
    out.append({"config": "a12 softmax fuse K=3 C=24 @100^2 x64", "ms_avg": round(savg, 4), "GBps": round(sb / savg / 1e6, 1), "frac_of_hbm": round(sb / savg / 1e6 / hbm, 3)})
    # ---- K-source transfer flows: one target raster, K composes
    Bm, Km = 30, 4
    cam5, verts5 = synth.smpl_poses(Bm * (Km + 1), seed=5, device=DEV)
    tcm, tvm = cam5[:Bm].contiguous(), verts5[:Bm].contiguous()
    scm, svm = cam5[Bm:].reshape(Bm, Km, 3).contiguous(), verts5[Bm:].reshape(Bm, Km, -1, 3).contiguous()
    mavg, _, _ = timeit(lambda: ops.cal_flow_multi(scm, svm, tcm, tvm, f_idx, 256), n=50)
    pavg, _, _ = timeit(lambda: [ops.cal_flow(scm[:, k].contiguous(), svm[:, k].contiguous(), tcm, tvm, f_idx, 256) for k in range(Km)], n=20)
    out.append({"config": "a8 x K: cal_flow_multi, 30 target frames x 4 source poses", "ms_avg": round(mavg, 4),
                "flows_per_s": round(Bm * Km / mavg * 1e3, 1), "four_cal_flow_calls_ms": round(pavg, 4)})
    # ---- row F per-reference visibility
    fim30 = ops.render_fim_wim(tc, tv, f_idx, 256, return_faces=False)[1]  # real face-index maps (~10 % foreground)
    ftgt = fim30.repeat(8, 1, 1).contiguous()
    fsrc = torch.stack([torch.roll(ftgt, k + 1, 0) for k in range(4)], 1).contiguous()
    vavg, _, _ = timeit(lambda: ops.face_visibility(fsrc, ftgt, 13776), n=30)
    frnd = torch.randint(-1, 13776, (240, 4, 256, 256), device=DEV, dtype=torch.int32)
    ravg, _, _ = timeit(lambda: ops.face_visibility(frnd, ftgt, 13776), n=10)
    vb = 240 * 65536 * (4 * 4 + 4 + 4 * 4)
    out.append({"config": "F per-reference visibility 240 frames x K=4 @256^2 (SMPL face-index maps)", "ms_avg": round(vavg, 4),
                "GBps": round(vb / vavg / 1e6, 1), "worst_case_random_fim_ms": round(ravg, 4)})
    # ---- §8f rank 3: bidirectional feature warp per SpatioTempoCRN level (B=8) vs the reference's op sequence in torch
    flow8 = torch.randn(8, 2, 256, 256, device=DEV) * 0.1
    tot_j = tot_t = 0.0
    for C_, s_ in ((64, 128), (128, 64), (256, 32), (512, 16), (512, 8), (512, 4)):
        pp, pl = torch.randn(8, C_, s_, s_, device=DEV), torch.randn(8, C_, s_, s_, device=DEV)
        ys, xs_ = torch.meshgrid(torch.linspace(-1, 1, s_, device=DEV), torch.linspace(-1, 1, s_, device=DEV), indexing="ij")
        g_ = torch.stack([xs_, ys])[None].expand(8, 2, s_, s_).contiguous()
        ja, _, _ = timeit(lambda: ops.flow_warp_pair(pp, pl, g_, flow8), n=50)

        def ref_level():
            fs = F.interpolate(flow8, (s_, s_), mode="nearest")
            a_ = F.grid_sample(pp, (g_ + fs).permute(0, 2, 3, 1), padding_mode="border", align_corners=False)
            b_ = F.grid_sample(pl, (g_ - fs).permute(0, 2, 3, 1), padding_mode="border", align_corners=False)
            return a_, b_
        ta, _, _ = timeit(ref_level, n=50)
        tot_j += ja
        tot_t += ta
    out.append({"config": "rank 3: 6-level bidirectional feature warp pyramid, B=8 (64ch@128^2 ... 512ch@4^2)",
                "ms_total": round(tot_j, 4), "torch_ref_ops_ms_total": round(tot_t, 4), "launches": 6, "torch_launches": 42})
    # ---- §8f rank 2 / 4: texture-space assembly and IUV preprocessing
    atlas = torch.randn(1, 5, 3, 800, 1200, device=DEV)
    idx = torch.tensor([0, 1, 2, 3], dtype=torch.int32, device=DEV)
    gavg, _, _ = timeit(lambda: ops.texture_parts_gather(atlas, idx), n=50)
    out.append({"config": "rank 2: gather 24 parts x 4 refs from the 800x1200 atlas", "ms_avg": round(gavg, 4),
                "GBps": round(2 * 4 * 3 * 800 * 1200 * 4 / gavg / 1e6, 1)})
    iuv = torch.randint(0, 25, (240, 256, 256, 3), device=DEV, dtype=torch.uint8)
    texu = torch.randint(0, 256, (800, 1200, 3), device=DEV, dtype=torch.uint8)
    imu = torch.randint(0, 256, (240, 256, 256, 3), device=DEV, dtype=torch.uint8)
    tavg2, _, _ = timeit(lambda: ops.transfer_texture(texu, iuv, imu), n=30)
    savg2, _, _ = timeit(lambda: ops.iuv_part_stats(iuv), n=30)
    out.append({"config": "rank 4: TransferTexture, 240 frames @256^2", "ms_avg": round(tavg2, 4), "frames_per_s": round(240 / tavg2 * 1e3, 1)})
    out.append({"config": "rank 4: compute_angle statistics, 240 frames @256^2", "ms_avg": round(savg2, 4), "frames_per_s": round(240 / savg2 * 1e3, 1)})
    tex24 = torch.randn(24, 3, 200, 200, device=DEV)
    xavg, _, _ = timeit(lambda: ops.texture_warp(tex24, iuv), n=30)
    out.append({"config": "rank 1: texture_warp (IUV lookup), 240 frames @256^2", "ms_avg": round(xavg, 4), "frames_per_s": round(240 / xavg * 1e3, 1)})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
