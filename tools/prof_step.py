#!/usr/bin/env python
"""One workload step repeated a few times, for ncu (launch lists / --set full captures) without the rest of bench.py.

    python tools/prof_step.py --what from_poses|warp_fuse|cal_flow|convlstm_grouped [--flow dense|hard|smpl] [--reps 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from jafpro_b200 import _lib, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="from_poses")
    ap.add_argument("--workload", default="dancevideo_256_k4_c64")
    ap.add_argument("--flow", default="dense")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--videos-per-gpu", type=int, default=0)
    args = ap.parse_args()
    if args.videos_per_gpu > 0:
        w = bench.WF_WORKLOADS[args.workload]
        bench.WF_WORKLOADS[args.workload] = (args.videos_per_gpu,) + tuple(w[1:])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    with torch.no_grad():
        if args.what == "from_poses":
            inp = bench.make_inputs(args.workload, 0, dev, "dense")
            poses = bench.pose_inputs(args, 0, dev)
            V, Fv, S, K, C = bench.WF_WORKLOADS[args.workload]
            fn = lambda: ops.warp_fuse_from_poses(poses["scam"], poses["sverts"], poses["tcam"], poses["tverts"], poses["f_idx"], S,
                                                  rgb=inp["rgb"], feat=inp["feat"], logits=inp["logits"], tgt_mask=inp["mask"])
        elif args.what == "warp_fuse":
            inp = bench.make_inputs(args.workload, 0, dev, args.flow)
            fn = lambda: ops.warp_fuse(inp["grid"], rgb=inp["rgb"], feat=inp["feat"], logits=inp["logits"], fim=inp["fim"],
                                       tgt_mask=inp["mask"])
        elif args.what == "cal_flow":
            from jafpro_b200.nmr import load_smpl_template
            f_idx = torch.from_numpy(load_smpl_template()[1]).to(dev)
            cam, verts = synth.smpl_poses(60, seed=3, device=dev)
            fn = lambda: ops.cal_flow(cam[:30].contiguous(), verts[:30].contiguous(), cam[30:].contiguous(), verts[30:].contiguous(),
                                      f_idx, 256)
        else:
            raise SystemExit("unknown --what")
        for _ in range(args.reps):
            fn()
        torch.cuda.synchronize()
        print(_lib.last_kernel())


if __name__ == "__main__":
    main()
