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


def synthetic_video(T: int = 6, seed: int = 5):
    """One synthetic video in the arrays of jafpro_b200.shards.VIDEO_ARRAYS / SMPL_ARRAYS (tools/make_golden.py writes
    it to disk in the reference's layout and runs the reference loader on it)."""
    rng = np.random.default_rng(seed)
    iuv4, _, im4 = iuv_preprocessing_inputs()
    S = 256
    iuv = np.stack([np.roll(iuv4[t % 4], 7 * t, axis=1) for t in range(T)])
    # make the torso halves differ per frame so that the view angles are distinct
    for t in range(T):
        iuv[t, 40:40 + 12 * (t + 1), 100:140, 0] = 2
        iuv[t, 150:150 + 9 * (T - t), 60:90, 0] = 1
    img = np.stack([np.roll(im4[t % 4], 3 * t, axis=0) for t in range(T)])
    ty, tx = np.mgrid[0:800, 0:1200]
    text = np.stack([np.stack([(ty * (3 + t) + tx * (5 + c) + 31 * t) % 256 for c in range(3)], -1) for t in range(T)]).astype(np.uint8)
    text_mask = np.stack([(((ty // 40 + tx // 60 + t) % 3) == 0).astype(np.uint8) * 255 for t in range(T)])
    real_mask = np.stack([np.repeat(((iuv[t, :, :, 0] > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2) for t in range(T)])
    return dict(img=img, iuv=iuv, text=text, text_mask=text_mask, real_mask=real_mask,
                cams=rng.normal(size=(T, 3)).astype(np.float32), pose=rng.normal(size=(T, 72)).astype(np.float32),
                shape=rng.normal(size=(T, 10)).astype(np.float32), vertices=rng.normal(size=(T, 6890, 3)).astype(np.float32))
