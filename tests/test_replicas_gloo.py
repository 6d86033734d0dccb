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


# ---------------------------------------------------------------------------------------------- training: GradSync
def _train_worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_ops
    import fake_train_ops
    import cenet_b200.train as T
    from cenet_b200.networks import CENet
    from oracle import fixtures
    T.ops, T.tops = fake_ops, fake_train_ops                        # CPU emulation of the kernels (host logic only)
    kw = fixtures.CONFIGS["acdc"]
    torch.manual_seed(1234 + rank)                                  # replicas start DIFFERENT: the broadcast must fix it
    m = CENet(**kw).train()
    eng = T.TrainEngine(m, "cpu", "fp32")
    eng.use_graph, eng.drop_path = False, False
    sync = replicas.GradSync(eng)
    p_start = eng.pflat.clone()
    x = fixtures.synth_input("acdc", 2, size=64, seed=rank)         # each rank owns its shard of the global batch
    y = torch.randint(0, 4, (2, 64, 64), generator=torch.Generator().manual_seed(rank))
    order = []
    launch = eng.on_bucket
    eng.on_bucket = lambda g: (order.append(g), launch(g))
    eng.train_step(x, y, optimize=False)
    g_sync = eng.gflat.clone()
    # local (un-synchronised) gradient of the same replica for the reference average
    eng.on_bucket, eng.grad_hook = None, None
    eng.train_step(x, y, optimize=False)
    ref = eng.gflat.clone()
    dist.all_reduce(ref)                                            # reference average with one plain collective
    ref /= world
    p_all = [torch.empty_like(p_start) for _ in range(world)]
    dist.all_gather(p_all, p_start)
    g_all = [torch.empty_like(g_sync) for _ in range(world)]
    dist.all_gather(g_all, g_sync)
    q.put((rank, all(torch.equal(p_all[0], t) for t in p_all), all(torch.equal(g_all[0], t) for t in g_all),
           (g_sync - ref).abs().max().item(), ref.abs().max().item(), order, dict(eng.bucket_ranges), eng.n_flat))
    dist.destroy_process_group()


def test_gradient_sync_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for (_, same_p, same_g, err, gmax, order, ranges, n) in res:
        assert same_p                                               # parameters broadcast from rank 0
        assert same_g                                               # every rank holds the same averaged gradient
        assert err <= 1e-6 * max(gmax, 1.0), (err, gmax)            # == mean of the per-rank gradients
        assert order == [5, 4, 3, 2, 1, 0]                          # head first, encoder stage 1 last
        covered = sorted(ranges.values())
        assert covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
