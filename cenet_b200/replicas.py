"""Batch-sharded replicas: one process per GPU, no data-path collective.

Inference needs no exchange between images, so the only distributed pieces are (a) which slices of a global batch a rank
owns, (b) the timing protocol of bench.py: barrier, per-rank device time, MAX over ranks, whole-job throughput.
The functions take a `torch.distributed` process group of any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous, balanced [lo, hi) range of `n_items` for `rank` (first `n_items % world` ranks get one extra)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def barrier(device=None):
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    r, w = world()
    if w > 1:
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of python floats over all ranks (the slowest rank defines the job time)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    r, w = world()
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def job_throughput(units_per_rank_step: int, steps: int, max_ms: float):
    """Whole-job units/s: every rank processed `units_per_rank_step * steps` units in at most `max_ms` (weak scaling)."""
    r, w = world()
    return w * units_per_rank_step * steps / (max_ms / 1e3)


def gather_labels(local: torch.Tensor):
    """Concatenate per-rank label maps on every rank (only needed when a caller wants the global batch back)."""
    r, w = world()
    if w == 1:
        return local
    out = [torch.empty_like(local) for _ in range(w)]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out, 0)
