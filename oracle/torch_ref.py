"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Row F written with exactly the torch primitives the reference calls, on the CPU:
the checker for the third-party grid_sample arithmetic and the CPU baseline that bench.py times
beside the GPU kernels (``cpu_baseline.kind == "port"``: our composition of the reference's call
sites, since the reference never fuses K warps itself).

    warped_k = F.grid_sample(ref_k, T_k, padding_mode='border')      src/cal_flow.py:37-39
    alpha    = nn.Softmax(dim=1)(logits)                              src/networks.py:1230-1244
    fused    = sum_k (alpha_k * vis_k) * warped_k                     src/networks.py:1276-1286
    fused   *= tgt_mask                                               src/flow_net.py:91
"""
import torch
import torch.nn.functional as F


@torch.no_grad()
def warp_fuse_torch(grid, rgb=None, feat=None, logits=None, vis=None, tgt_mask=None, align_corners=False):
    """grid [B,K,H,W,2]; rgb [B,K,3,Hs,Ws]; feat [B,K,C,Hs,Ws] (fp32, NCHW: the reference's layout)."""
    B, K = grid.shape[:2]
    a = torch.softmax(logits, dim=1) if logits is not None else torch.full(grid.shape[:4], 1.0 / K)
    if vis is not None:
        a = a * vis
    outs = []
    for t in (rgb, feat):
        if t is None:
            outs.append(None)
            continue
        acc = None
        for k in range(K):
            w = F.grid_sample(t[:, k], grid[:, k], mode="bilinear", padding_mode="border",
                              align_corners=align_corners)
            w = w * a[:, k:k + 1]
            acc = w if acc is None else acc + w
        if tgt_mask is not None:
            acc = acc * tgt_mask[:, :1] if t is feat or tgt_mask.shape[1] == 1 else acc * tgt_mask
        outs.append(acc)
    return outs[0], outs[1]
