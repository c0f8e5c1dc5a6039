"""Multi-GPU plumbing: one process per GPU, videos sharded round-robin by rank, NO collective on the
hot path (every (video, frame) is independent); torch.distributed (NCCL on GPUs, gloo in
CPU tests) is used only for the barrier around the timed region and the final gather of counters."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str | None = None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            # NCCL writes its banner ("NCCL version ...") and debug lines to stdout by default; rank 0's stdout is
            # reserved for the one JSON line of bench.py
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_cpus(local: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to GPU `local` (same NUMA node / PCIe root), so that
    pinned host buffers allocated afterwards are first-touched on that node.  Matters for the host-buffer path when
    several ranks share a two-socket host; harmless otherwise.  Returns False when NVML is not usable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            props = torch.cuda.get_device_properties(local)
            bus_id = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode() if hasattr(bus_id, "encode") else bus_id)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


def shard_videos(n_videos: int, rank: int, world: int):
    """Rank r takes videos r, r+G, r+2G, ... (keeps a video's K references and 30 frames on one GPU)."""
    return list(range(rank, n_videos, world))


def barrier():
    if dist.is_initialized():
        dist.barrier()


def reduce_max_sum(local_ms: float, local_units: float, device=None):
    """-> (max over ranks of local_ms, sum over ranks of local_units).  The only collective of a run."""
    if not dist.is_initialized():
        return float(local_ms), float(local_units)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    u = torch.tensor([local_units], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def gather_results(local: torch.Tensor, dst: int = 0):
    """Final result gather (NCCL all_gather over NVLink on GPUs).  Returns the list on every rank."""
    if not dist.is_initialized():
        return [local]
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local)
    return out
