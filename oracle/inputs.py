"""Deterministic synthetic inputs shared by tools/make_golden.py (which runs the reference on them) and the tests.
Test infrastructure only."""
import numpy as np


def iuv_preprocessing_inputs():
    """Deterministic inputs shared by the fixture generator and the tests (256x256: TransferTexture hard-codes it)."""
    rng = np.random.default_rng(11)
    n, S = 4, 256
    iuv = np.zeros((n, S, S, 3), np.uint8)
    yy, xx = np.mgrid[0:S, 0:S]
    for i in range(n):
        part = np.zeros((S, S), np.uint8)
        for pid in rng.permutation(np.arange(1, 25))[: 8 + 4 * i]:
            cx, cy, r = rng.integers(30, S - 30), rng.integers(30, S - 30), rng.integers(10, 30)
            part[(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = pid
        iuv[i, :, :, 0] = part
        iuv[i, :, :, 1] = (xx * 3 + yy + 17 * i) % 256
        iuv[i, :, :, 2] = (yy * 5 + xx // 2 + 29 * i) % 256
        iuv[i][part == 0] = 0
    iuv[2, :, :, 0][iuv[2, :, :, 0] == 2] = 0               # no frontal torso
    iuv[3, :, :, 0][np.isin(iuv[3, :, :, 0], (1, 2))] = 0   # neither torso half (the nan path of compute_angle)
    ty, tx = np.mgrid[0:800, 0:1200]
    tex = np.stack([(ty * 7 + tx * 13 + c * 101) % 256 for c in range(3)], -1).astype(np.uint8)
    tex[(ty + tx) % 11 == 0] = 0                            # zero texels: TransferTexture treats them as background
    im = np.stack([np.stack([(xx + 2 * yy + 50 * c + 9 * i) % 256 for c in range(3)], -1) for i in range(n)]).astype(np.uint8)
    return iuv, tex, im
