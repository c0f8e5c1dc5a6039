"""Synthetic DanceVideo-shaped inputs: random SMPL poses from the packaged T-pose
template, reference images / feature maps, dense smooth transfer flows.  Seeded and device-agnostic
(generation uses torch ops on whatever device is asked for; it is input plumbing, not the path)."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .nmr import load_smpl_template


def smpl_poses(n: int, seed: int = 0, device="cpu", jitter: float = 2e-3, motion: bool = True):
    """-> cam [n,3] (s,tx,ty), verts [n,6890,3] float32.  Template centred, rotated about y by
    theta ~ U(-0.6, 0.6) (plus a slow sinusoid over the n frames when `motion`), per-vertex N(0, 2 mm) jitter."""
    g = torch.Generator().manual_seed(seed)
    v, _ = load_smpl_template()
    v = torch.from_numpy(v)
    v = v - v.mean(0, keepdim=True)
    theta = (torch.rand(n, generator=g) * 1.2 - 0.6)
    if motion:
        theta = theta * 0.3 + 0.4 * torch.sin(torch.linspace(0, 2 * math.pi, n) + float(torch.rand(1, generator=g)) * 6.28)
    cs, sn = torch.cos(theta), torch.sin(theta)
    x = cs[:, None] * v[None, :, 0] + sn[:, None] * v[None, :, 2]
    z = -sn[:, None] * v[None, :, 0] + cs[:, None] * v[None, :, 2]
    verts = torch.stack([x, v[None, :, 1].expand(n, -1), z], -1)
    verts = verts + torch.randn(verts.shape, generator=g) * jitter
    cam = torch.stack([0.75 + 0.2 * torch.rand(n, generator=g), -0.05 + 0.1 * torch.rand(n, generator=g),
                       0.25 - 0.05 + 0.1 * torch.rand(n, generator=g)], -1)
    return cam.float().contiguous().to(device), verts.float().contiguous().to(device)


def identity_grid(H: int, W: int, device="cpu"):
    """Pixel-centre NDC grid [H,W,2] (x, y): what the rasteriser emits for an untouched frame
    (align_corners=False convention, rasterize_cuda_kernel.cu:96-97)."""
    xs = (2 * torch.arange(W, device=device, dtype=torch.float32) + 1 - W) / W
    ys = (2 * torch.arange(H, device=device, dtype=torch.float32) + 1 - H) / H
    return torch.stack([xs[None, :].expand(H, W), ys[:, None].expand(H, W)], -1)


def dense_flows(B: int, K: int, H: int, W: int, seed: int = 0, device="cpu", max_disp_px: float = 8.0):
    """[B,K,H,W,2] identity + low-frequency displacement of at most `max_disp_px` pixels: every target
    pixel is visible and samples a nearby source pixel (worst case for the fused kernel: no background
    pixels to skip)."""
    g = torch.Generator(device=device).manual_seed(seed)
    coarse = torch.rand((B * K, 2, 5, 5), generator=g, device=device) * 2 - 1
    disp = F.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=True).clamp_(-1, 1)
    disp = disp.permute(0, 2, 3, 1).reshape(B, K, H, W, 2)
    scale = torch.tensor([2.0 * max_disp_px / W, 2.0 * max_disp_px / H], device=device)
    return (identity_grid(H, W, device)[None, None] + disp * scale).contiguous()


def hard_flows(B: int, K: int, H: int, W: int, seed: int = 0, device="cpu", block: int = 32,
               max_disp_px: float = 64.0, max_rot: float = 0.6, scale_range=(0.7, 1.4)):
    """[B,K,H,W,2] full-frame coverage, piecewise-affine like real SMPL transfer flows (one affine map per body
    part / triangle, src/nmr.py:617-659) but without the background: the frame is cut into `block`-pixel cells and
    every (frame, reference, cell) gets its own rotation (|theta| <= max_rot rad), isotropic scale and translation
    (up to +-max_disp_px pixels).  Gather locality is what real flows have INSIDE a part and is broken at every
    cell boundary; samples that leave the reference are clamped by the border rule like real ones."""
    g = torch.Generator(device=device).manual_seed(seed)
    by, bx = (H + block - 1) // block, (W + block - 1) // block
    n = B * K
    th = (torch.rand((n, by, bx), generator=g, device=device) * 2 - 1) * max_rot
    sc = scale_range[0] + (scale_range[1] - scale_range[0]) * torch.rand((n, by, bx), generator=g, device=device)
    tx = (torch.rand((n, by, bx), generator=g, device=device) * 2 - 1) * max_disp_px
    ty = (torch.rand((n, by, bx), generator=g, device=device) * 2 - 1) * max_disp_px
    ys = torch.arange(H, device=device, dtype=torch.float32)
    xs = torch.arange(W, device=device, dtype=torch.float32)
    cy = (torch.div(ys, block, rounding_mode="floor") + 0.5) * block  # cell centres
    cx = (torch.div(xs, block, rounding_mode="floor") + 0.5) * block
    iy = torch.div(ys, block, rounding_mode="floor").long()
    ix = torch.div(xs, block, rounding_mode="floor").long()
    up = lambda t: t[:, iy][:, :, ix]                                  # [n,H,W]
    dx = (xs - cx)[None, None, :].expand(n, H, W)
    dy = (ys - cy)[None, :, None].expand(n, H, W)
    c, s_ = torch.cos(up(th)) * up(sc), torch.sin(up(th)) * up(sc)
    m = 0.75 * block                                                   # keep the source cell (mostly) inside the frame
    scx = (cx[None, None, :] + up(tx)).clamp(m, W - m)
    scy = (cy[None, :, None] + up(ty)).clamp(m, H - m)
    sx = scx + c * dx - s_ * dy                                        # source position in pixels
    sy = scy + s_ * dx + c * dy
    gx = (2 * sx + 1 - W) / W                                          # pixel-centre NDC (align_corners=False)
    gy = (2 * sy + 1 - H) / H
    return torch.stack([gx, gy], -1).reshape(B, K, H, W, 2).contiguous()


def perm_flows(B: int, K: int, H: int, W: int, seed: int = 0, device="cpu"):
    """[B,K,H,W,2] worst case for any gather: every target pixel samples the centre of an independently drawn random
    source pixel (no two neighbouring pixels share a cache line of the reference)."""
    g = torch.Generator(device=device).manual_seed(seed)
    flat = torch.empty((B * K, H * W), dtype=torch.int64, device=device)
    for i in range(B * K):
        flat[i] = torch.randperm(H * W, generator=g, device=device)
    sy, sx = torch.div(flat, W, rounding_mode="floor").float(), (flat % W).float()
    gx = (2 * sx + 1 - W) / W
    gy = (2 * sy + 1 - H) / H
    return torch.stack([gx, gy], -1).reshape(B, K, H, W, 2).contiguous()


def reference_sets(R: int, K: int, C: int, H: int, W: int, seed: int = 0, device="cpu", channels_last: bool = True):
    """-> rgb [R,K,3,H,W] f32 in [-1,1] (data range of src/data.py:591), feat [R,K,C,H,W] bf16 ~ N(0,1)
    (channels-last strides when asked)."""
    g = torch.Generator(device=device).manual_seed(seed)
    rgb = torch.randn((R, K, 3, H, W), generator=g, device=device).clamp_(-1, 1)
    feat = None
    if C > 0:
        if channels_last:
            feat = torch.empty((R, K, H, W, C), dtype=torch.bfloat16, device=device)
            for r in range(R):  # chunked: the fp32 temporary of one reference set at a time
                feat[r] = torch.randn((K, H, W, C), generator=g, device=device).to(torch.bfloat16)
            feat = feat.permute(0, 1, 4, 2, 3)
        else:
            feat = torch.empty((R, K, C, H, W), dtype=torch.bfloat16, device=device)
            for r in range(R):
                feat[r] = torch.randn((K, C, H, W), generator=g, device=device).to(torch.bfloat16)
    return rgb, feat


def warp_fuse_bytes(K: int, H: int, W: int, C: int) -> int:
    """Algorithmic bytes per target frame of the fused warp+fuse: every input read once,
    outputs written once, no credit for cache reuse.  256^2, K=4, C=64 -> 49,283,072."""
    return K * H * W * (8 + 4 + 3 * 4 + C * 2) + H * W * 4 + H * W * (3 * 4 + C * 2)
