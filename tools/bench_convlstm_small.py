#!/usr/bin/env python
"""Row a13 at the reference's real cell sizes: the five pyramid levels of Accumulate_LSTM_no_loss
(src/networks.py:1304-1313), 24 part-specific cells each, one recurrent step.  One JSON line per level:
the grouped tcgen05 split-bf16 kernel, the exact-fp32 CUDA-core kernel (24 launches), and the reference's
own op sequence in torch on the same GPU (24 x [cat, cuDNN conv, split, gates]; TF32 off and on)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from jafpro_b200 import ops  # noqa: E402
from tools.bench_aux import timeit  # noqa: E402

DEV = "cuda"
LEVELS = [(12, 200), (24, 100), (24, 50), (48, 25), (96, 13)]


def ref_step(x, h, c, w, b):
    cc = F.conv2d(torch.cat((x, h), 1), w, b, padding=1)
    i, f, o, g = torch.split(cc, h.shape[1], dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def main():
    G, B = 24, 1
    tot = {"grouped": 0.0, "f32": 0.0, "torch_fp32": 0.0, "torch_tf32": 0.0}
    for Ch, S in LEVELS:
        torch.manual_seed(Ch + S)
        x, h, c = (torch.randn(G, B, Ch, S, S, device=DEV) for _ in range(3))
        w = torch.randn(G, 4 * Ch, 2 * Ch, 3, 3, device=DEV) * (1.5 / (18 * Ch) ** 0.5)
        b = torch.randn(G, 4 * Ch, device=DEV)
        wp = ops.convlstm_gpack_weight(w, Ch, Ch)
        t_g, _, best_g = timeit(lambda: ops.convlstm_step_grouped(x, h, c, wp, b, Ch, Ch), n=30)
        t_f, _, _ = timeit(lambda: [ops.convlstm_step(x[g], h[g], c[g], w[g], b[g]) for g in range(G)], n=5, warm=2)
        torch.backends.cudnn.allow_tf32 = False
        t_t, _, _ = timeit(lambda: [ref_step(x[g], h[g], c[g], w[g], b[g]) for g in range(G)], n=10, warm=3)
        torch.backends.cudnn.allow_tf32 = True
        t_t32, _, _ = timeit(lambda: [ref_step(x[g], h[g], c[g], w[g], b[g]) for g in range(G)], n=10, warm=3)
        torch.backends.cudnn.allow_tf32 = False
        h2, c2 = ops.convlstm_step_grouped(x, h, c, wp, b, Ch, Ch)
        err = 0.0
        for g in range(0, G, 8):
            hr, cr = ref_step(x[g].double(), h[g].double(), c[g].double(), w[g].double(), b[g].double())
            err = max(err, float((h2[g].double() - hr).abs().max()), float((c2[g].double() - cr).abs().max()))
        flop = 2.0 * G * B * S * S * 4 * Ch * 9 * 2 * Ch
        byts = G * B * S * S * Ch * 4 * 6 + w.numel() * 4  # x, h, c in; h', c' out (+1 c): fp32
        tot["grouped"] += t_g; tot["f32"] += t_f; tot["torch_fp32"] += t_t; tot["torch_tf32"] += t_t32
        print(json.dumps({"config": f"a13 24 cells Ch={Ch} @{S}^2 B=1, one step", "grouped_tc_ms": round(t_g, 4),
                          "grouped_tc_best_ms": round(best_g, 4), "cuda_core_fp32_24_launches_ms": round(t_f, 4),
                          "torch_ref_ops_fp32_ms": round(t_t, 4), "torch_ref_ops_tf32_ms": round(t_t32, 4),
                          "max_abs_err_vs_fp64": err, "TFLOPs_useful": round(flop / t_g / 1e9, 1),
                          "GBps_compulsory": round(byts / t_g / 1e6, 1)}))
    print(json.dumps({"config": "a13 whole pyramid (5 levels x 24 parts), one step", **{k + "_ms": round(v, 4) for k, v in tot.items()}}))
    # the five levels of a step are independent (src/networks.py:1346-1355 runs five separate ConvLSTMs): the whole pyramid
    # step (a) back to back on one stream, (b) one stream per level (fork / join by events), (c) = (b) as a CUDA graph
    lv = []
    for Ch, S in LEVELS:
        torch.manual_seed(Ch + S)
        x, h, c = (torch.randn(G, B, Ch, S, S, device=DEV) for _ in range(3))
        w = torch.randn(G, 4 * Ch, 2 * Ch, 3, 3, device=DEV) * (1.5 / (18 * Ch) ** 0.5)
        b = torch.randn(G, 4 * Ch, device=DEV)
        ho, co = torch.empty_like(h), torch.empty_like(c)
        lv.append((x, h, c, ops.convlstm_gpack_weight(w, Ch, Ch), b, Ch, ho, co))

    def serial():
        for x, h, c, wp, b, Ch, ho, co in lv:
            ops.convlstm_step_grouped(x, h, c, wp, b, Ch, Ch)

    def forked(order=(0, 1, 2, 3, 4)):
        return ops.run_concurrently([(lambda t=lv[i]: ops.convlstm_step_grouped(t[0], t[1], t[2], t[3], t[4], t[5], t[5]))
                                     for i in order])

    res = {"config": "a13 whole pyramid step as ONE unit of work"}
    res["serial_one_stream_ms"] = round(timeit(serial, n=30)[0], 4)
    # (eager: the results are handed to the current stream with record_stream, so the caching allocator needs a long
    # warm-up before its per-stream pools stop growing; the graph flavour below is the intended use)
    res["one_stream_per_level_ms"] = round(timeit(forked, n=30, warm=60)[0], 4)
    g = ops.FrameGraph(serial)
    res["serial_cuda_graph_ms"] = round(timeit(g.replay, n=30)[0], 4)
    g2 = ops.FrameGraph(forked)
    res["one_stream_per_level_cuda_graph_ms"] = round(timeit(g2.replay, n=30)[0], 4)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
