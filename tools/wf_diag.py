#!/usr/bin/env python
"""Diagnostics: time the fused warp+fuse kernel under feature toggles (GPU only)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jafpro_b200 import ops, synth

def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    dev = "cuda"
    B, K, C, S = 120, 4, 64, 256
    rgb, feat = synth.reference_sets(B, K, C, S, S, seed=1, device=dev)
    dense = synth.dense_flows(B, K, S, S, seed=1, device=dev)
    ident = synth.identity_grid(S, S, dev)[None, None].expand(B, K, S, S, 2).contiguous()
    logits = torch.randn(B, K, S, S, device=dev)
    mask = torch.ones(B, 1, S, S, device=dev)
    cases = {
        "full(dense)": dict(grid=dense, rgb=rgb, feat=feat, logits=logits, tgt_mask=mask),
        "full(identity)": dict(grid=ident, rgb=rgb, feat=feat, logits=logits, tgt_mask=mask),
        "no-rgb": dict(grid=dense, feat=feat, logits=logits, tgt_mask=mask),
        "no-rgb no-logits no-mask": dict(grid=dense, feat=feat),
        "rgb-only(generic)": dict(grid=dense, rgb=rgb, logits=logits, tgt_mask=mask),
    }
    for name, kw in cases.items():
        g = kw.pop("grid")
        ms = timeit(lambda: ops.warp_fuse(g, **kw))
        by = synth.warp_fuse_bytes(K, S, S, C if "feat" in kw else 0)
        if "rgb" not in kw: by -= K * S * S * 12 + S * S * 12
        print(f"{name:28s} {ms:8.3f} ms  {B/ms*1000:9.0f} fps  {by*B/ms/1e6:8.1f} GB/s")
    for k in (1, 2, 8):
        g = synth.dense_flows(B, k, S, S, seed=2, device=dev)
        r, f = synth.reference_sets(B, k, C, S, S, seed=2, device=dev)
        l = torch.randn(B, k, S, S, device=dev)
        ms = timeit(lambda: ops.warp_fuse(g, rgb=r, feat=f, logits=l, tgt_mask=mask))
        print(f"K={k:<26d} {ms:8.3f} ms  {B/ms*1000:9.0f} fps  {synth.warp_fuse_bytes(k,S,S,C)*B/ms/1e6:8.1f} GB/s")
        del g, r, f, l
    # plain copy of the same volume for reference
    x = torch.empty(B * synth.warp_fuse_bytes(K, S, S, C) // 8, dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    ms = timeit(lambda: y.copy_(x))
    print(f"{'torch copy (r+w)':28s} {ms:8.3f} ms  {2*x.numel()*4/ms/1e6:8.1f} GB/s")

if __name__ == "__main__":
    main()
