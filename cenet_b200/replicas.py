"""Batch-sharded replicas: one process per GPU.

Inference needs no exchange between images, so its only distributed pieces are (a) which slices of a global batch a rank
owns, (b) the timing protocol of bench.py: barrier, per-rank device time, MAX over ranks, whole-job throughput.
Training adds ONE collective per step: the gradient all-reduce (`GradSync`), bucketed per backward-completion group and
overlapped with the rest of the backward pass.  BatchNorm statistics stay per GPU, as under the reference's
`nn.DataParallel` (main_acdc.py:179).
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


class GradSync:
    """Gradient averaging for `cenet_b200.train.TrainEngine` replicas.

    The engine keeps all gradients in one flat fp32 buffer laid out in backward-completion order (encoder stage 1..4,
    decoder, head).  Its backward pass calls `on_bucket(g)` the moment group g is final -- head first, encoder stage 1
    last -- and this class answers with an asynchronous all-reduce of that contiguous range (6 buckets, 1.5-60 MB).
    NCCL runs it on its own stream, ordered after the kernels already enqueued and concurrent with the backward kernels
    that follow; `finish` (the engine's grad_hook, right before AdamW) waits for all of them.  With CUDA graphs the
    engine cuts its capture at the bucket markers and issues these calls between graph replays."""

    def __init__(self, engine, group=None, exposed=False):
        """exposed=True is a measurement mode (bench.py `synapse_dp`): every bucket's all-reduce is waited for inline, so
        the backward kernels that follow cannot overlap it -- the step-time difference to the default is the overlap gain."""
        self.eng, self.group, self.exposed = engine, group, exposed
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self.works, self.pending = [], []
        if self.world > 1:
            engine.on_bucket = self.launch
            engine.grad_hook = self.finish
            self.broadcast_parameters()

    def broadcast_parameters(self):
        """rank 0's parameters / buffers / optimizer state become everyone's (replicas start identical)"""
        for t in (self.eng.pflat, self.eng.adam_m, self.eng.adam_v):
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)

    def launch(self, g):
        lo, hi = self.eng.bucket_ranges[g]
        t = self.eng.gflat[lo:hi]
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        self.works.append(dist.all_reduce(t, op=op, group=self.group, async_op=True))
        if not self.avg:
            self.pending.append(t)
        if self.exposed:
            self.works[-1].wait()                     # the compute stream now waits for this all-reduce before going on

    def finish(self, _gflat=None):
        for w in self.works:
            w.wait()
        for t in self.pending:
            t.div_(self.world)
        self.works, self.pending = [], []
