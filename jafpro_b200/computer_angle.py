"""Drop-in for ``src/computer_angle.py``: the view-angle heuristic the data loader uses to pick reference frames
(src/data.py:504).  The per-part pixel counts and x sums come from one histogram kernel over the whole batch of IUV
maps; the scalar formula (:27-39) is evaluated on the host in float64 exactly as the reference writes it."""
from __future__ import annotations

import numpy as np
import torch

from . import ops

FRONT_INDEX = [2, 9, 10, 13, 14]  # :5
BACK_INDEX = [1, 7, 8, 11, 12]    # :6


def _angle_from_stats(counts, sumx):
    front_area = float(sum(float(counts[p]) for p in FRONT_INDEX))
    back_area = float(sum(float(counts[p]) for p in BACK_INDEX))
    n_front, n_back = int(counts[2]), int(counts[1])
    frontal_avg_x = np.float64(sumx[2]) / np.float64(n_front) if n_front > 0 else np.float64("nan")  # np.average([]) = nan
    back_avg_x = np.float64(sumx[1]) / np.float64(n_back) if n_back > 0 else frontal_avg_x           # :20-23
    if n_front == 0:
        frontal_avg_x = back_avg_x                                                                   # :24-25
    if frontal_avg_x < back_avg_x:
        ratio = (front_area + 10e-5) / (back_area + 10e-5)
        angle = np.arctan(ratio) / np.pi * 180 - 90
    else:
        ratio = -(front_area + 10e-5) / (back_area + 10e-5)
        angle = np.arctan(ratio) / np.pi * 180 + 90
    if angle < -65:
        return 65
    return angle


def compute_angles(iuv_batch):
    """IUV maps [B,H,W,3] (uint8 tensor on the GPU, or numpy) -> list of B angles, one histogram launch."""
    if isinstance(iuv_batch, np.ndarray):
        iuv_batch = torch.from_numpy(np.ascontiguousarray(iuv_batch)).cuda()
    counts, sumx = ops.iuv_part_stats(iuv_batch.to(torch.uint8).contiguous())
    counts, sumx = counts.cpu().numpy(), sumx.cpu().numpy()
    return [_angle_from_stats(counts[b], sumx[b]) for b in range(counts.shape[0])]


def compute_angle(IUV):
    """Same call as the reference: one IUV map [H,W,3] -> angle."""
    if isinstance(IUV, np.ndarray):
        return compute_angles(IUV[None])[0]
    return compute_angles(IUV[None])[0]
