#!/usr/bin/env python
"""Bit-exactness stress of the scatter rasteriser against the reference's own CUDA kernels (oracle/_ref): many SMPL
poses at 256^2 / 512^2 and soups of tiny, randomly placed triangles (the regime where the pixel box margin matters).
Prints the number of mismatching pixels; used to validate JAF_RASTER_MARGIN settings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from jafpro_b200 import ops, synth  # noqa: E402
from jafpro_b200.nmr import SMPLRenderer  # noqa: E402

DEV = "cuda"
ref = oracle.RefRaster()
bad = 0
tot = 0


def cmp(faces, size, tag):
    global bad, tot
    a = ops.raster_fim_wim(faces, size, return_depth=True)
    b = ref(faces, size)
    nf = int((a[0] != b[0]).sum())
    nw = int((a[1].view(torch.int32) != b[1].view(torch.int32)).sum())
    nd = int((a[2].view(torch.int32) != b[2].view(torch.int32)).sum())
    bad += nf + nw + nd
    tot += a[0].numel()
    print(f"{tag}: fim {nf} wim {nw} depth {nd} mismatches over {a[0].numel()} px, fg {float((b[0] != -1).float().mean()):.3f}", flush=True)


for size, n, seed in ((256, 48, 101), (512, 8, 202), (128, 16, 303)):
    rend = SMPLRenderer(image_size=size).to(DEV)
    cam, verts = synth.smpl_poses(n, seed=seed, device=DEV)
    for i in range(0, n, 8):
        faces, _, _ = rend.render_fim_wim(cam[i:i + 8].contiguous(), verts[i:i + 8].contiguous())
        cmp(faces, size, f"smpl {size} [{i}:{i + 8}]")
g = torch.Generator().manual_seed(7)
for size, nf, scale in ((128, 40000, 0.01), (128, 40000, 0.03), (64, 20000, 0.004), (200, 30000, 0.02), (256, 60000, 0.006)):
    c = torch.rand((2, nf, 1, 3), generator=g) * 2.2 - 1.1
    tri = c + (torch.rand((2, nf, 3, 3), generator=g) - 0.5) * 2 * scale
    tri[..., 2] = torch.rand((2, nf, 3), generator=g) * 3 + 0.5
    cmp(tri.to(DEV).contiguous(), size, f"soup {size} x{nf} scale {scale}")
print("TOTAL mismatches", bad, "of", tot, "pixels")
