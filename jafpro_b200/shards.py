"""Packed binary video shards: the data-format side of the hot path.

The reference keeps a video as a directory of image files plus a pickle and re-decodes everything per item
(``Fusion_dataset_smpl_test.__getitem__``, src/data.py:471-602: ~5 ``cv2.imread`` per frame, ``pose_shape.pkl`` with
``cams / pose / shape / vertices``, src/data.py:583-596).  Once the GPU path runs at 10^5 frames/s that decode is the
whole cost, so a video is packed ONCE into one flat file of raw arrays:

    "JAFSHRD1" | u64 header length | JSON header (padded) | array payloads, each 256-byte aligned

``VideoShard`` maps the file (zero-copy numpy views, one H2D copy per array) and ``load_test_item`` rebuilds exactly
what the reference loader returns — frame selection by view angle included (``compute_angle`` on the GPU), same
values, same dtypes — with the tensors on the device.  ``pack_video_dir`` converts the reference's on-disk layout
(needs OpenCV; it is the only place that does).
"""
from __future__ import annotations

import json
import os
import pickle
import struct

import numpy as np
import torch

MAGIC = b"JAFSHRD1"
ALIGN = 256
# array name -> (dtype, trailing shape); the leading dimension is the frame count T
VIDEO_ARRAYS = {
    "img": (np.uint8, (256, 256, 3)),          # frame_N.png            (BGR, as cv2.imread returns it)
    "iuv": (np.uint8, (256, 256, 3)),          # frame_N_IUV.png
    "text": (np.uint8, (800, 1200, 3)),        # frame_N_text.png       (24-part texture atlas)
    "text_mask": (np.uint8, (800, 1200)),      # frame_N_mask.png[..., 0]
    "real_mask": (np.uint8, (256, 256, 3)),    # <mask_root>/frame_N_mask.png
}
SMPL_ARRAYS = ("cams", "pose", "shape", "vertices")  # pose_shape.pkl fields (src/data.py:583-596)


def _pad(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def write_shard(path: str, arrays: dict, meta: dict) -> None:
    """arrays: name -> numpy array (any dtype / shape); meta: JSON-serialisable dict."""
    table, off = {}, 0
    blobs = []
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        table[name] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": off, "nbytes": int(a.nbytes)}
        blobs.append(a)
        off += _pad(a.nbytes)
    header = json.dumps({"version": 1, "meta": meta, "arrays": table}).encode("utf-8")
    hlen = _pad(len(MAGIC) + 8 + len(header)) - len(MAGIC) - 8
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", hlen))
        f.write(header.ljust(hlen, b" "))
        for a in blobs:
            f.write(a.tobytes())
            f.write(b"\0" * (_pad(a.nbytes) - a.nbytes))


class VideoShard:
    """Read side: ``shard["img"]`` is a zero-copy numpy view of the mapped file; ``shard.tensor("img", device)`` the same
    array on a device (uploaded once and cached)."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path}: not a jafpro_b200 video shard")
            (hlen,) = struct.unpack("<Q", f.read(8))
            hdr = json.loads(f.read(hlen).decode("utf-8"))
        if hdr.get("version") != 1:
            raise ValueError(f"{path}: unsupported shard version {hdr.get('version')}")
        self.meta, self._table = hdr["meta"], hdr["arrays"]
        self._base = len(MAGIC) + 8 + hlen
        self._map = np.memmap(path, dtype=np.uint8, mode="r")
        self._dev = {}

    def keys(self):
        return list(self._table)

    def __getitem__(self, name: str) -> np.ndarray:
        e = self._table[name]
        lo = self._base + e["offset"]
        return self._map[lo:lo + e["nbytes"]].view(np.dtype(e["dtype"])).reshape(e["shape"])

    def tensor(self, name: str, device="cuda") -> torch.Tensor:
        key = (name, str(device))
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(np.array(self[name])).to(device, non_blocking=True)
        return self._dev[key]

    @property
    def num_frames(self) -> int:
        return int(self._table["img"]["shape"][0])


def pack_video(path: str, video: dict, vid_name: str, img_names) -> None:
    """video: the arrays of VIDEO_ARRAYS + SMPL_ARRAYS, leading dimension T."""
    T = video["img"].shape[0]
    for name, (dt, tail) in VIDEO_ARRAYS.items():
        a = video[name]
        if a.dtype != dt or tuple(a.shape) != (T,) + tail:
            raise ValueError(f"{name}: expected {np.dtype(dt)} {(T,) + tail}, got {a.dtype} {a.shape}")
    arrays = {k: video[k] for k in list(VIDEO_ARRAYS) + list(SMPL_ARRAYS)}
    write_shard(path, arrays, {"vid_name": vid_name, "img_names": list(img_names), "num_frames": int(T)})


def pack_video_dir(path: str, vid_path: str, smpl_pkl: str, mask_dir: str) -> None:
    """Convert one video of the reference's on-disk layout (src/utils.py:26-58: frame_N.png, frame_N_IUV.png,
    frame_N_text.png, frame_N_mask.png, <mask_dir>/frame_N_mask.png, pose_shape.pkl) into a shard.  Needs OpenCV."""
    import cv2

    def num(fn, tail):
        return int(fn[6:-tail])

    files = os.listdir(vid_path)
    plain = sorted((f for f in files if all(t not in f for t in ("IUV", "mask", "text", "bbox", "pkl"))), key=lambda f: num(f, 4))
    iuv = sorted((f for f in files if f.find("IUV") > 0), key=lambda f: num(f, 8))
    msk = sorted((f for f in files if f.find("mask") > 0), key=lambda f: num(f, 9))
    txt = sorted((f for f in files if f.find("text") > 0), key=lambda f: num(f, 9))
    real = sorted(os.listdir(mask_dir), key=lambda f: num(f, 9))
    rd = lambda d, f: cv2.imread(os.path.join(d, f))
    with open(smpl_pkl, "rb") as fh:
        smpl = pickle.load(fh)
    video = {"img": np.stack([rd(vid_path, f) for f in plain]), "iuv": np.stack([rd(vid_path, f) for f in iuv]),
             "text": np.stack([rd(vid_path, f) for f in txt]), "text_mask": np.stack([rd(vid_path, f)[:, :, 0] for f in msk]),
             "real_mask": np.stack([rd(mask_dir, f) for f in real])}
    video.update({k: np.asarray(smpl[k]) for k in SMPL_ARRAYS})
    pack_video(path, video, os.path.basename(vid_path.rstrip("/")), plain)


def select_reference_frames(angle: np.ndarray, num_inputs: int):
    """The view-angle heuristic of src/data.py:505-527 -> (pro_frames, frames)."""
    T = angle.shape[0]
    max_index, min_index = np.argmax(angle), np.argmin(angle)
    if num_inputs == 4:
        order = np.argsort(angle)
        frames = np.array([max_index, order[T // 3], order[T * 2 // 3], min_index], int)
    elif num_inputs == 1:
        frames = np.array([np.argmin(np.abs(angle))], int)
    elif num_inputs < 4:
        frames = np.array([max_index, np.argsort(angle)[T // 2], min_index], int)
    elif num_inputs == 5:
        order = np.argsort(angle)
        frames = np.array([max_index, order[T // 4], order[T * 2 // 4], order[T * 3 // 4], min_index], int)
    else:
        raise ValueError("num_inputs must be 1..5")
    return frames, np.clip(frames, 0, 30)


def load_test_item(shard: VideoShard, num_inputs: int = 4, output_mask: bool = True, device="cuda"):
    """``Fusion_dataset_smpl_test.__getitem__`` (src/data.py:471-602) from a shard, tensors on `device`:
    -> (src_data, tgt_data, data_255, smpl_data, vid_name, img_name_list, pro_frames) with the reference's values and
    dtypes (float64 where numpy's ``/ 255.0`` produces it).  The only host work is the scalar tail of the frame
    selection; the per-part statistics, TransferTexture and every normalisation run on the device."""
    from .computer_angle import compute_angles
    from .utils import TransferTexture
    T = shard.num_frames
    iuv = shard.tensor("iuv", device)
    angle = np.array([float(a) for a in compute_angles(iuv)], np.float64)
    pro_frames, frames = select_reference_frames(angle, num_inputs)
    if len(frames) != num_inputs:  # the reference fills `num_inputs` slots from `frames` (src/data.py:537-551)
        raise ValueError("num_inputs = 2 is not served by the reference's frame selection either")
    idx = torch.from_numpy(frames.astype(np.int64)).to(device)
    # numpy's `/ 255.0` is an IEEE division; torch's CUDA division by a Python scalar multiplies by the rounded
    # reciprocal, so the divisor is a device tensor (tensor / tensor is a true division)
    d255 = torch.full((), 255.0, dtype=torch.float64, device=device)
    norm = lambda t: (t.double() / d255 - 0.5) * 2
    img, text, text_mask = shard.tensor("img", device), shard.tensor("text", device), shard.tensor("text_mask", device)
    src_iuv255, src_mask_u8 = iuv[idx], text_mask[idx]
    src_data = [norm(img[idx]), norm(src_iuv255), norm(text[idx]), src_mask_u8.double() / d255]
    tgt_data = [norm(img), norm(iuv)]
    data_255 = [src_iuv255, iuv]
    if output_mask:
        src_common_area = (src_mask_u8 != 0).any(dim=0)                       # OR of (mask / 255) != 0, src/data.py:566-568
        ones = torch.ones((800, 1200, 3), dtype=torch.uint8, device=device)
        src_data.append(src_common_area)
        src_data.append(TransferTexture(ones, src_iuv255.contiguous()))        # src/data.py:570-572
    smpl_seq = torch.from_numpy(np.concatenate([shard["cams"], shard["pose"], shard["shape"]], axis=1)).to(device)
    smpl_data = [smpl_seq, shard.tensor("real_mask", device).double() / d255, shard.tensor("vertices", device)]
    return src_data, tgt_data, data_255, smpl_data, shard.meta["vid_name"], list(shard.meta["img_names"]), pro_frames
