"""N>1 path on CPU: world_size-2 gloo process group exercising the batch sharding + timing protocol of bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cenet_b200 import replicas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = replicas.shard_range(129, rank, world)
    replicas.barrier()
    t_dev, t_e2e = replicas.max_over_ranks([10.0 + rank, 20.0 - rank])
    thr = replicas.job_throughput(64, 10, t_dev)
    labels = torch.full((hi - lo if False else 2, 4, 4), rank, dtype=torch.int64)
    allb = replicas.gather_labels(labels)
    q.put((rank, lo, hi, t_dev, t_e2e, thr, allb[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, td0, te0, thr0, l0), (r1, lo1, hi1, td1, te1, thr1, l1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 65, 65, 129)                # balanced contiguous shards, nothing dropped
    assert td0 == td1 == 11.0 and te0 == te1 == 20.0              # MAX over ranks, identical on every rank
    assert abs(thr0 - 2 * 64 * 10 / 0.011) < 1e-6 and thr0 == thr1  # whole-job aggregate (weak scaling)
    assert l0 == l1 == [0, 0, 1, 1]


def test_single_process_defaults():
    assert replicas.world() == (0, 1)
    assert replicas.shard_range(10, 0, 1) == (0, 10)
    assert replicas.max_over_ranks([3.5]) == [3.5]
    assert [replicas.shard_range(7, r, 3) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
