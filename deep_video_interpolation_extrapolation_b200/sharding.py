"""Batch sharding of the warp over the GPUs of one box.

The op has no cross-sample term (utils/net_utils.py:93-114 warps every sample with its own flow), so the
multi-GPU path is the reference's own: contiguous batch slices, one process per GPU, exactly what its
DistributedSampler + `batch_size // gpus` does (runners/InterTrainer.py:84-87, main.py:154).  No collective is
issued by the op; the helpers below are the host logic bench.py and the tests share.
"""
from __future__ import annotations

from typing import Sequence, Tuple


def batch_slice(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the samples rank `rank` of `world` owns: contiguous, sizes differ by at most one,
    earlier ranks take the remainder (empty slices are legal: the op accepts N == 0)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n < 0:
        raise ValueError("negative batch")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def chunk_slices(n: int, chunk: int):
    """Consecutive `slice`s of at most `chunk` samples covering [0, n) (the host pipeline's batch chunks)."""
    if chunk < 1:
        raise ValueError("chunk must be >= 1")
    if n < 0:
        raise ValueError("negative batch")
    return [slice(a, min(a + chunk, n)) for a in range(0, n, chunk)]


def shard(tensors: Sequence, rank: int, world: int):
    """Slice every tensor of a clip batch along dim 0 with `batch_slice` (views, no copies)."""
    if not tensors:
        return []
    a, b = batch_slice(tensors[0].shape[0], rank, world)
    return [t[a:b] for t in tensors]


def max_over_ranks(seconds: float, device=None) -> float:
    """Step time of the job = the slowest rank's device time (bench.py's timing rule).  No-op without a process
    group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def job_throughput(units_per_rank: float, steps: int, seconds_max: float, world: int) -> float:
    """Whole-job units per second under weak scaling: every rank processes `units_per_rank` per step."""
    return world * units_per_rank * steps / seconds_max


def bind_to_gpu_numa(device_index: int) -> int:
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index`, so that the pinned host
    buffers it allocates afterwards (first touch) sit on the GPU's own NUMA node: with one process per GPU, eight
    ranks copying 1.3 GB per step each otherwise fight over one socket's memory controllers.  Returns the number
    of cores bound to (0 = left as is: NVML unavailable, no affinity support, or an empty mask)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001 - best effort: the binding is an optimisation, never a requirement
        return 0


def sync(loss_dict, mean: bool = True, world: int | None = None):
    """Drop-in for the trainers' `sync(loss_dict, mean)` (runners/InterTrainer.py:859-864, ExtraTrainer.py:760-765): every
    tensor of `loss_dict` is all-reduced in place and, with `mean`, divided by the number of ranks.

    The reference issues ONE `dist.all_reduce` (+ one `div_`) per scalar loss — ~40 latency-bound NCCL launches per step
    (SURVEY 8f row 4).  Here all values are flattened into one buffer per (device, dtype), reduced with ONE collective and
    scattered back, so the cost is one launch latency regardless of how many losses the dictionary holds.  Results are
    identical to the per-tensor loop (the same sum per element; the division is the same elementwise op).  `world`
    defaults to the process group's size (the reference divides by `args.gpus`)."""
    import torch
    import torch.distributed as dist
    tensors = [t for t in loss_dict.values() if isinstance(t, torch.Tensor)]
    if not tensors:
        return loss_dict
    on = dist.is_available() and dist.is_initialized()
    if world is None:
        world = dist.get_world_size() if on else 1
    groups = {}
    for t in tensors:
        groups.setdefault((t.device, t.dtype), []).append(t)
    for (_, _), ts in groups.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        if on and dist.get_world_size() > 1:
            dist.all_reduce(flat)
        if mean:
            flat.div_(world)
        off = 0
        with torch.no_grad():
            for t in ts:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
    return loss_dict
