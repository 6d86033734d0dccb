#!/usr/bin/env python
"""bench.py -- CENet hot-path throughput on B200 (BASELINE.json metric: 224^2 slices/s, batched inference).

    python bench.py --gpus 1 --steps 20 --warmup 5             # product arm (hand-written sm_100a kernels)
    python bench.py --impl reference --steps 3 --warmup 1      # reference arm: the REAL reference (baseline/_ref) on host cores
    python bench.py --impl reference-gpu                       # the reference module in PyTorch eager on the same B200 (the bar)
    torchrun --nproc-per-node N bench.py --gpus N ...          # one rank per GPU, weak scaling (batch 64 per GPU)

Workload (BASELINE.json configs[1]): Synapse 9-class batched slice inference, 224x224, batch 64 per GPU, bf16
storage / fp32 accumulation, random-init weights (deterministic recipe oracle.fixtures), synthetic slices, including
the fused softmax/argmax -> int64 label map.  A step = one forward pass over one batch.

  value : slices/s with the batch resident in HBM (device time, CUDA events per step, L2 flushed between steps,
          max over ranks).
  e2e   : same metric through the public call `CENet.predict` with pinned HOST buffers: H2D of the slices and D2H
          of the label maps inside the timed region.
  roofline : dominant kernel (differential flash attention of the 56x56 DSE block) timed live with CUDA events in
          an eager pass; achieved = algorithmic FLOPs / duration against the measured bf16 peak.
  cpu_baseline : the reference itself (baseline/_ref, byte copy of /root/reference/src; the oracle port if that copy is
          missing) on this box's host cores, bounded sample, rank 0 at N=1 only.
  eager    : the reference module in PyTorch eager (cuDNN / cuBLAS) on the SAME B200, fp32 and autocast(bf16) -- the bar the
          hand-written kernels have to beat; `vs_eager` = value / best eager value.  N=1 only.
  skin512  : BASELINE.json configs[3] (512x512 skin, infer + train), N=1 only.
  synapse_dp : BASELINE.json configs[4] (Synapse training, batch 32 per GPU, gradient all-reduce overlapped vs exposed), N>1 only.
  train    : second half of BASELINE.json's metric -- training images/s on configs[2] (ACDC, batch 24 per GPU, Dice+CE, AdamW),
          same timing protocol; under torchrun the gradients are all-reduced over NCCL in 6 buckets overlapped with backward.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG_NAME = "synapse"
TRAIN_CONFIG = "acdc"
TRAIN_BATCH = 24
BATCH = 64
SIZE = 224
METRIC = "slices_per_s_infer_224"
UNIT = "slices/s"


# dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu captures of this workload (profiles/, per round)
NCU_TRAFFIC = {"source": "profiles/r2_ncu_top_kernels.md, profiles/r2_gemm_traffic.md", "diffattn_flash_kernel": 191.4e6,
               "gemm_tc_kernel": 43.6e6}      # mean over the 146 launches of one forward


def workload_config(world=1):
    """the `config` object: identical in the product line and the reference line (the driver compares them)"""
    return {"workload": f"CENet (PVTv2-b2) Synapse 9-class batched slice inference {SIZE}x{SIZE}, batch {BATCH} per GPU, "
                        "logits -> fused softmax/argmax int64 labels (BASELINE.json configs[1])",
            "batch_per_gpu": BATCH, "size": SIZE, "num_classes": 9,
            "l2": "flushed between timed steps (256 MB memset; GPU arms)",
            "parallelism": "replicas (batch-sharded, no data-path collective); gradient all-reduce only in the train objects"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (profiling recipe's clocks line)."""

    def __init__(self, index):
        self.rows = []
        self._stop = threading.Event()
        self.index = index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


def build_model():
    import torch
    from cenet_b200.networks import CENet
    from oracle import fixtures
    import contextlib
    kw = fixtures.CONFIGS[CONFIG_NAME]
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):          # the constructor prints like the reference does
        m = CENet(**kw)
    sd = fixtures.perturb_state(m.state_dict(), 1234)
    m.load_state_dict(sd)
    return m, sd, kw


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_oracle_throughput(sd, kw, sample_batch, steps, warmup):
    """Times the oracle port (the reference's algorithm restated in plain torch fp32) on the host cores."""
    import torch
    from oracle import cenet_oracle as O
    from oracle import fixtures
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = fixtures.synth_input(CONFIG_NAME, sample_batch, SIZE)
    cfg = O.Cfg(**kw)
    with torch.no_grad():
        for _ in range(warmup):
            O.predict_labels(O.cenet_forward(sd, cfg, x))
        t0 = time.perf_counter()
        for _ in range(steps):
            O.predict_labels(O.cenet_forward(sd, cfg, x))
        dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps * 1e3, cores


def cpu_reference_throughput(sd, kw, sample_batch, steps, warmup):
    """-> (slices/s, ms/step, cores, kind): the reference itself when baseline/_ref is present, else the oracle port"""
    from baseline import reference_arms as R
    if R.available():
        v, ms, cores = R.cpu_inference(CONFIG_NAME, sample_batch, steps, warmup)
        return v, ms, cores, "reference"
    v, ms, cores = cpu_oracle_throughput(sd, kw, sample_batch, steps, warmup)
    return v, ms, cores, "port"


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores (all threads), on a bounded
    sample of the product arm's workload: 8 slices per step (at batch 64 the reference needs ~80 GB of RAM)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, sd, kw = build_model()
    sample = 8
    v, ms, cores, kind = cpu_reference_throughput(sd, kw, sample, args.steps, args.warmup)
    what = ("the unmodified reference (baseline/_ref/src/networks, PyTorch eager fp32)" if kind == "reference"
            else "the oracle port (baseline/_ref not present)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{what}: {sample} slices/step x {args.steps} steps of the same synthetic Synapse workload "
                                   "(forward + argmax(softmax))"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_reference_gpu(args):
    """`--impl reference-gpu`: the reference module in PyTorch eager on cuda:0 (the real bar); prints ONE JSON object"""
    from baseline import reference_arms as R
    if not R.available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "baseline/_ref not vendored (tools/vendor_reference.py)"}))
        return
    res = R.gpu_eager(args.steps, args.warmup, infer_name=CONFIG_NAME, infer_batch=BATCH, train_name=TRAIN_CONFIG,
                      train_batch=TRAIN_BATCH, modes=args.eager_modes.split(",") if args.eager_modes else None)
    res["impl"] = "reference-gpu"
    res["steps"], res["warmup"] = args.steps, max(args.warmup, 3)
    print(json.dumps(res))


def eager_subprocess(steps, warmup, timeout=900):
    """runs `bench.py --impl reference-gpu` in its own process (own CUDA context, no cenet_b200 state) and parses its line"""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-gpu", "--steps", str(steps), "--warmup", str(warmup)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": f"rc {r.returncode}: {r.stderr[-400:]}"}
    except Exception as e:                                       # a failed baseline must not take the product line down
        return {"error": repr(e)}


def cpu_oracle_train_throughput(sample_batch, steps):
    """One training iteration (forward, Dice+CE, autograd backward, AdamW) of the oracle port on the host cores."""
    import contextlib
    import torch
    from cenet_b200.networks import CENet
    from oracle import cenet_oracle as O
    from oracle import fixtures
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = fixtures.CONFIGS[TRAIN_CONFIG]
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):
        sd = fixtures.perturb_state(CENet(**kw).state_dict(), 1234)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    opt = torch.optim.AdamW([leaf[k] for k in names], lr=1e-4, weight_decay=1e-4)
    x = fixtures.synth_input(TRAIN_CONFIG, sample_batch, SIZE)
    y = torch.randint(0, kw["num_classes"], (sample_batch, SIZE, SIZE), generator=torch.Generator().manual_seed(5))
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = O.criterion_dice_ce(O.cenet_forward(leaf, O.Cfg(**kw), x, training=True), y, kw["num_classes"])
        opt.zero_grad()
        loss.backward()
        opt.step()
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps * 1e3, cores


# ---------------------------------------------------------------------------------------------------- product arm
def run_train_leg(args, dev, world, rank, flush, name=TRAIN_CONFIG, batch=TRAIN_BATCH, size=SIZE, steps=None,
                  sync_mode="overlapped", with_e2e=True):
    """A training leg.  Default = BASELINE.json configs[2]: ACDC 4-class training step (Dice+CE 0.5/0.5, AdamW lr 1e-4 wd
    1e-4), 224x224, batch 24 per GPU, train-mode BatchNorm + DropPath, bf16 activations / fp32 master weights.  A step =
    forward, fused loss, backward, gradient all-reduce (N > 1), AdamW -- all hand-written kernels, replayed from CUDA graph
    segments.  sync_mode (N > 1): "overlapped" = bucketed NCCL all-reduce launched as each bucket completes, "exposed" = each
    all-reduce waited for inline, "none" = no gradient exchange (compute only)."""
    import contextlib
    import torch
    from cenet_b200 import ops, replicas
    from cenet_b200.networks import CENet
    from oracle import fixtures

    class _A:                                                     # local view of args with this leg's step count
        pass
    a_ = _A()
    a_.steps, a_.warmup = (steps or args.steps), args.warmup
    args = a_
    TRAIN_BATCH, SIZE = batch, size                               # (shadow the module constants for this leg)
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):
        m = CENet(**kw)
    m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
    m = m.to(dev).train()
    eng = m.train_engine(dev)
    sync = replicas.GradSync(eng, exposed=(sync_mode == "exposed")) if (world > 1 and sync_mode != "none") else None
    x_host = fixtures.synth_input(name, TRAIN_BATCH, SIZE, seed=100 + rank).pin_memory()
    y_host = torch.randint(0, kw["num_classes"], (TRAIN_BATCH, SIZE, SIZE), generator=torch.Generator().manual_seed(200 + rank)).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    loss_host = torch.zeros(1 + kw["num_classes"]).pin_memory()
    n0 = ops.launch_count()
    for _ in range(max(args.warmup, 3)):
        eng.train_step(x_dev, y_dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    replicas.barrier(dev)
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        loss = eng.train_step(x_dev, y_dev)
        e1.record()
    replicas.barrier(dev)
    t_dev = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    if not with_e2e:
        (t_dev,) = replicas.max_over_ranks([t_dev], device=dev)
        torch.cuda.synchronize(dev)
        res = {"value": replicas.job_throughput(TRAIN_BATCH, args.steps, t_dev), "unit": "img/s", "ms_per_step": t_dev / args.steps,
               "steps": args.steps, "batch_per_gpu": TRAIN_BATCH, "sync_mode": sync_mode if world > 1 else "single GPU",
               "launches_per_step": int(eng.launches_per_step or 0), "gpu_launches": int(ops.launch_count() - n0)}
        m._engines.clear()
        del eng, m, sync
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        return res
    for _ in range(max(args.warmup, 3)):                     # warm-up of the end-to-end path (first-touch allocations)
        loss_host.copy_(eng.train_step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True)), non_blocking=True)
    # end to end: every step copies its images + label maps from pinned host memory (on a copy stream, double-buffered, so the
    # H2D of step i+1 overlaps the kernels of step i), runs the public fused step and reads the loss back; ONE timed bracket
    # around all K steps (it includes the L2-flush memsets)
    main = torch.cuda.current_stream()
    s_in = torch.cuda.Stream(dev)
    xs, ys = [torch.empty_like(x_dev) for _ in range(2)], [torch.empty_like(y_dev) for _ in range(2)]
    ev_in, ev_free = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        j = i % 2
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[j])
            xs[j].copy_(x_host, non_blocking=True)
            ys[j].copy_(y_host, non_blocking=True)
            ev_in[j].record(s_in)
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    replicas.barrier(dev)
    s_in.wait_stream(main)
    e_start.record(main)
    stage(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            stage(i + 1)
        main.wait_event(ev_in[i % 2])
        flush.zero_()
        loss = eng.train_step(xs[i % 2], ys[i % 2])         # public call
        ev_free[i % 2].record(main)
        loss_host.copy_(loss, non_blocking=True)            # D2H of the loss + per-class dice
    e_end.record(main)
    replicas.barrier(dev)
    t_e2e = e_start.elapsed_time(e_end)
    t_dev, t_e2e = replicas.max_over_ranks([t_dev, t_e2e], device=dev)
    torch.cuda.synchronize(dev)
    return {
        "metric": "train_imgs_per_s_224", "unit": "img/s",
        "value": replicas.job_throughput(TRAIN_BATCH, args.steps, t_dev), "ms_per_step": t_dev / args.steps,
        "e2e": {"value": replicas.job_throughput(TRAIN_BATCH, args.steps, t_e2e), "unit": "img/s",
                "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": loss_host.numel() * 4,
                "ms_per_step": t_e2e / args.steps},
        "config": {"workload": f"CENet (PVTv2-b2) {name} {kw['num_classes']}-class training step {SIZE}x{SIZE}, batch {TRAIN_BATCH} per GPU, "
                               "Dice+CE, AdamW, train-mode BatchNorm, DropPath on", "batch_per_gpu": TRAIN_BATCH,
                   "l2": "flushed between steps", "cuda_graph": bool(eng.use_graph), "flops_per_image": 76.0e9 * (SIZE / 224.0) ** 2,
                   "parallelism": f"dp{world} (gradient all-reduce in 6 buckets overlapped with backward)" if world > 1 else "dp1"},
        "dtype": "bf16", "launches_per_step": int(eng.launches_per_step or 0),
        "gpu_launches": int(ops.launch_count() - n0), "final_loss": float(loss_host[0]),
        "model_tflops": 76.0e9 * TRAIN_BATCH * world / (t_dev / args.steps / 1e3) / 1e12 if SIZE == 224 else None,
        "steps": args.steps,
    }


def run_infer_leg(args, dev, flush, name, batch, size, steps):
    """an extra inference leg (another BASELINE config) with inputs resident in HBM: CUDA events per step, L2 flushed"""
    import contextlib
    import torch
    from cenet_b200.networks import CENet
    from oracle import fixtures
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):
        m = CENet(**kw)
    m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
    m = m.to(dev).eval()
    x = fixtures.synth_input(name, batch, size).to(dev)
    eng = m._engine(x)
    out = torch.empty((batch, size, size), device=dev, dtype=torch.int64)
    for _ in range(3):
        eng.forward(x, labels=True, out=out)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize(dev)
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        eng.forward(x, labels=True, out=out)
        e1.record()
    torch.cuda.synchronize(dev)
    t = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    prof = eng.profile_ops(x, labels=True, steps=1) if steps > 1 else {}
    top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]
    res = {"value": batch * steps / (t / 1e3), "unit": "img/s", "ms_per_step": t / steps, "steps": steps, "batch": batch,
           "launches_per_step": int(eng.launches_per_forward or 0),
           "op_breakdown_ms": {f"{op}@{tag}": round(ms, 4) for (op, tag), (ms, n) in top}, "_prof": prof}
    m._engines.clear()
    del eng, m
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def diffattn_flops(N, E, B):
    """Algorithmic FLOPs of the differential attention core: QK^T + PV over 2h maps = 4*N^2*E per image."""
    return 4.0 * N * N * E * B


def run_product(args):
    import torch
    import torch.distributed as dist
    from cenet_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the product arm needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from oracle import fixtures
    m, sd, kw = build_model()
    m = m.to(dev).eval()
    x_host = fixtures.synth_input(CONFIG_NAME, BATCH, SIZE, seed=rank).pin_memory()
    x_dev = x_host.to(dev)
    eng = m._engine(x_dev)
    labels_dev = torch.empty((BATCH, SIZE, SIZE), device=dev, dtype=torch.int64)
    labels_host = torch.empty((BATCH, SIZE, SIZE), dtype=torch.int64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)        # > 126 MB L2

    from cenet_b200 import replicas

    def barrier():
        replicas.barrier(dev)

    # ---- warm-up (also captures the CUDA graph) ----
    for _ in range(max(args.warmup, 3)):
        eng.forward(x_dev, labels=True, out=labels_dev)
    launches_per_step = eng.launches_per_forward
    # ---- timed region 1: inputs resident in HBM ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    clocks = ClockSampler(local)                            # samples clocks / throttle reasons over ALL timed regions below
    clocks.__enter__()
    for e0, e1 in ev:
        flush.zero_()                                       # L2 flush between timed iterations (not timed)
        e0.record()
        eng.forward(x_dev, labels=True, out=labels_dev)
        e1.record()
    barrier()
    t_dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    # ---- timed region 2: end to end through the public API with host buffers ----
    for _ in range(max(args.warmup, 3)):
        labels_host.copy_(m.predict(x_host.to(dev, non_blocking=True)), non_blocking=True)
    # Every step: H2D of its slices from pinned host memory, the public call `CENet.predict`, D2H of its int64 label maps.
    # The copies run on their own streams (double-buffered device input), so the transfer of step i+1 / i-1 overlaps the
    # kernels of step i, as a serving loop would do; the timed region is ONE bracket around all K steps incl. the last D2H
    # (and incl. the L2 flush memsets, which a per-step bracket would have excluded).
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    xd = [torch.empty_like(x_dev) for _ in range(2)]
    ld = [torch.empty_like(labels_dev) for _ in range(2)]   # rotating device label buffers (no allocator traffic in the loop)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        j = i % 2
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[j])                 # predict(i-2) has consumed this staging buffer
            xd[j].copy_(x_host, non_blocking=True)
            ev_in[j].record(s_in)
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s_in.wait_stream(main); s_out.wait_stream(main)
    e_start.record(main)
    stage(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            stage(i + 1)
        main.wait_event(ev_in[i % 2])
        if i >= 2:
            main.wait_event(ev_out[i % 2])                  # the D2H of step i-2 has drained this label buffer
        flush.zero_()
        lab = m.predict(xd[i % 2], out=ld[i % 2])           # public call: logits -> argmax(softmax) fused
        ev_free[i % 2].record(main)
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            labels_host.copy_(lab, non_blocking=True)       # D2H of the label maps
            ev_out[i % 2].record(s_out)
    main.wait_stream(s_out)
    e_end.record(main)
    barrier()
    t_e2e_ms = e_start.elapsed_time(e_end)
    t_dev_ms, t_e2e_ms = replicas.max_over_ranks([t_dev_ms, t_e2e_ms], device=dev)
    # ---- sustained run (SURVEY 8d asks >= 100 timed iterations; the driver's K is usually 20) and single-slice latency ----
    n_sus = 100
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_sus)]
    barrier()
    for e0, e1 in ev2:
        flush.zero_()
        e0.record()
        eng.forward(x_dev, labels=True, out=labels_dev)
        e1.record()
    barrier()
    (t_sus_ms,) = replicas.max_over_ranks([sum(e0.elapsed_time(e1) for e0, e1 in ev2)], device=dev)
    x1 = x_dev[:1].contiguous()
    lab1 = torch.empty((1, SIZE, SIZE), device=dev, dtype=torch.int64)
    for _ in range(3):
        eng.forward(x1, labels=True, out=lab1)
    ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    for e0, e1 in ev3:
        e0.record()
        eng.forward(x1, labels=True, out=lab1)
        e1.record()
    torch.cuda.synchronize(dev)
    lat = sorted(e0.elapsed_time(e1) for e0, e1 in ev3)
    latency_b1 = {"median_ms": lat[len(lat) // 2], "min_ms": lat[0], "launches": int(eng.launches_per_forward or 0),
                  "note": "B=1 forward + fused argmax, CUDA-graph replay (the reference's per-slice eval mode, metrics_eval.py:48-52)"}
    eng.forward(x_dev, labels=True, out=labels_dev)                 # restore the batch-64 plan as the engine's current one

    train = None
    if not args.no_train:
        train = run_train_leg(args, dev, world, rank, flush)
    # ---- BASELINE.json configs[4]: Synapse data-parallel training, batch 32 per GPU, all-reduce overlapped vs exposed ----
    synapse_dp = None
    if world > 1 and not args.no_train:
        synapse_dp = {"workload": f"CENet Synapse 9-class training step 224x224, batch 32 per GPU x {world} GPUs = global batch "
                                  f"{32 * world} (BASELINE.json configs[4]), Dice+CE, AdamW"}
        for mode in ("overlapped", "exposed", "none"):
            synapse_dp[mode] = run_train_leg(args, dev, world, rank, flush, name="synapse", batch=32, size=SIZE,
                                             steps=max(10, args.steps // 2), sync_mode=mode, with_e2e=False)
        synapse_dp["allreduce_exposed_cost_ms"] = synapse_dp["exposed"]["ms_per_step"] - synapse_dp["none"]["ms_per_step"]
        synapse_dp["allreduce_overlapped_cost_ms"] = synapse_dp["overlapped"]["ms_per_step"] - synapse_dp["none"]["ms_per_step"]
    # ---- BASELINE.json configs[3]: 512x512 skin (3-channel, binary), inference + training, one GPU ----
    skin512 = None
    if world == 1 and not args.no_skin512:
        peaks = _peaks()
        inf = run_infer_leg(args, dev, flush, "skin", 16, 512, max(5, args.steps // 4))
        prof5 = inf.pop("_prof")
        N5 = (512 // 4) ** 2
        k1 = prof5.get(("diffattn_flash", "se1"))
        k2 = prof5.get(("nonlocal_flash", "dec1"))
        if k1:
            fl = diffattn_flops(N5, 128, 16)
            inf["roofline_K1_diffattn_N16384"] = {"bound": "tensor", "achieved": fl / (k1[0] / 1e3) / 1e12, "peak": peaks["tf_sustained"],
                                                  "unit": "TFLOP/s", "frac": fl / (k1[0] / 1e3) / 1e12 / peaks["tf_sustained"],
                                                  "ms_per_launch": k1[0], "flops_per_launch": fl}
        if k2:
            fl = 4.0 * N5 * N5 * 64 * 16
            inf["roofline_K2_nonlocal_N16384"] = {"bound": "tensor", "achieved": fl / (k2[0] / 1e3) / 1e12, "peak": peaks["tf_sustained"],
                                                  "unit": "TFLOP/s", "frac": fl / (k2[0] / 1e3) / 1e12 / peaks["tf_sustained"],
                                                  "ms_per_launch": k2[0], "flops_per_launch": fl}
        inf["model_tflops"] = 330.4e9 * inf["value"] / 1e12
        skin512 = {"workload": "CENet skin (HAM10000-shaped) binary segmentation 512x512 (BASELINE.json configs[3]): inference batch 16 "
                               "(fused argmax), training batch 8 (Dice+CE 0.5/0.5, AdamW)", "infer": inf}
        if not args.no_train:
            skin512["train"] = run_train_leg(args, dev, world, rank, flush, name="skin", batch=8, size=512,
                                             steps=max(5, args.steps // 4), with_e2e=False)
            skin512["train"]["model_tflops"] = 3 * 330.4e9 * skin512["train"]["value"] / 1e12
    clocks.__exit__()

    if rank == 0:
        peaks = _peaks()
        value = replicas.job_throughput(BATCH, args.steps, t_dev_ms)
        e2e = replicas.job_throughput(BATCH, args.steps, t_e2e_ms)
        # ---- roofline of the dominant kernel, timed live (eager launches bracketed by CUDA events) ----
        prof = eng.profile_ops(x_dev, labels=True, steps=4)
        total_ms = sum(v[0] for v in prof.values())
        by_op = {}
        for (op, tag), (ms, n) in prof.items():
            a = by_op.setdefault(op, [0.0, 0]); a[0] += ms; a[1] += n
        top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]
        # (a) the dominant kernel of the step: gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv; every nn.Linear, 1x1,
        #     3x3 and 5x5 conv) -- all its launches of one forward, algorithmic bytes (each operand once) over the
        #     summed launch durations.  Mostly K <= 128 shapes -> HBM-bound; the tensor-pipe view is stated next to it.
        gw = dict(eng.last_gemm_work)
        mf_work = gw.pop("mixffn_tail", None)
        g_bytes = sum(v[0] for v in gw.values()); g_flops = sum(v[1] for v in gw.values()); g_n = sum(v[2] for v in gw.values())
        g_ms = sum(by_op.get(op, [0.0, 0])[0] for op in ("linear", "gemm", "conv_nhwc"))
        roof = None
        if g_n and g_ms > 0:
            ach = g_bytes / (g_ms / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": f"gemm_tc_kernel (tcgen05 GEMM / conv), {g_n} launches per forward",
                    "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": NCU_TRAFFIC.get("gemm_tc_kernel"), "traffic_source": NCU_TRAFFIC.get("source"),
                    "bytes_per_launch": g_bytes / g_n, "ms_per_launch": g_ms / g_n, "share_of_step": g_ms / total_ms,
                    "tensor_view": {"achieved_tflops": g_flops / (g_ms / 1e3) / 1e12, "peak_tflops": peaks["tf_sustained"],
                                    "frac": g_flops / (g_ms / 1e3) / 1e12 / peaks["tf_sustained"]},
                    "peak_source": peaks["source"] + ", HBM copy",
                    "note": "op-level times of linear/gemm/conv_nhwc calls; a few fp32/batched calls on the CUDA-core GEMM "
                            "are in the time but not in the bytes (conservative)"}
            # per-launch roofline: each launch against ITS bound, max(bytes / HBM peak, FLOPs / tensor peak)
            gl = getattr(eng, "last_gemm_launches", [])
            if gl:
                t_roof = sum(max(by / (peaks["hbm_gbs"] * 1e9), fl / (peaks["tf_sustained"] * 1e12)) * 1e3 for _, _, by, fl, _ in gl)
                t_act = sum(ms for *_, ms in gl)
                roof["per_launch_view"] = {"t_roofline_ms": t_roof, "t_measured_ms": t_act, "frac": t_roof / t_act,
                                           "note": "sum over launches of max(bytes/HBM peak, FLOPs/bf16 peak) / measured time"}
                hb, tf = peaks["hbm_gbs"] * 1e9, peaks["tf_sustained"] * 1e12
                exc = sorted(gl, key=lambda r: -(r[4] - max(r[2] / hb, r[3] / tf) * 1e3))[:12]
                roof["top_excess"] = [{"op": op, "tag": tag, "ms": round(ms, 4), "roofline_ms": round(max(by / hb, fl / tf) * 1e3, 4)}
                                      for op, tag, by, fl, ms in exc]
                worst = sorted(gl, key=lambda r: -r[4])[:16]
                roof["top_launches"] = [{"op": op, "tag": tag, "ms": round(ms, 4), "MB": round(by / 1e6, 1), "GFLOP": round(fl / 1e9, 2),
                                         "GB/s": round(by / ms / 1e6), "TFLOP/s": round(fl / ms / 1e9, 1)} for op, tag, by, fl, ms in worst]
        # (a2) the fused Mix-FFN tail (depthwise 3x3 + GELU -> fc2 MMAs + residual, mixffn_tc.cu): HBM view on its algorithmic bytes
        roof_mf = None
        if mf_work and by_op.get("mixffn_tail", [0.0, 0])[0] > 0:
            mf_ms = by_op["mixffn_tail"][0]
            ach = mf_work[0] / (mf_ms / 1e3) / 1e9
            roof_mf = {"bound": "hbm", "kernel": f"mixffn_tail_kernel (tcgen05 + TMA), {mf_work[2]} launches per forward", "achieved": ach,
                       "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "ms_per_launch": mf_ms / mf_work[2],
                       "share_of_step": mf_ms / total_ms, "bytes_per_launch": mf_work[0] / mf_work[2],
                       "note": "replaces dwconv3x3 + the fc2 GEMM of the stage-1/2 encoder blocks; CUDA-core bound "
                               "(~27 instructions per depthwise output), see DESIGN.md 4a"}
        # (b) the largest single launch: differential flash attention of the 56x56 DSE block (exp-bound)
        E1, N1 = 128, (SIZE // 4) ** 2
        ms_da, n_da = prof.get(("diffattn_flash", "se1"), (None, 0))
        roof_attn = None
        if ms_da:
            fl = diffattn_flops(N1, E1, BATCH)
            ach = fl / (ms_da / 1e3) / 1e12
            roof_attn = {"bound": "tensor", "kernel": "diffattn_tc_kernel<8,1> (tcgen05; DSEB 56x56, 16 softmax maps, head_dim 8)",
                         "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                         "traffic": NCU_TRAFFIC.get("diffattn_flash_kernel"), "ms_per_launch": ms_da,
                         "share_of_step": ms_da / total_ms, "peak_source": peaks["source"] + ", sustained bf16",
                         "mufu_bound": {"exp_per_launch": 16.0 * N1 * N1 * BATCH, "peak_exp_per_s": 148 * 16 * 1.965e9,
                                        "frac": 16.0 * N1 * N1 * BATCH / (ms_da / 1e3) / (148 * 16 * 1.965e9)},
                         "note": "softmax-bound: one MUFU.EX2 per score (16/clk/SM); see DESIGN.md 4a"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": labels_host.numel() * 8, "ms_per_step": t_e2e_ms / args.steps},
            "gpu_launches": int(launches_per_step * args.steps * 2 + launches_per_step * 2) + (train["gpu_launches"] if train else 0),
            "launches_per_step": int(launches_per_step),
            "cuda_graph": bool(eng.use_graph),
            "clocks": clocks.summary(),
            "roofline": roof,
            "roofline_attention": roof_attn,
            "roofline_mixffn": roof_mf,
            "sustained": {"steps": n_sus, "value": replicas.job_throughput(BATCH, n_sus, t_sus_ms), "unit": UNIT,
                          "ms_per_step": t_sus_ms / n_sus},
            "latency_b1": latency_b1,
            "model_tflops": 25.35e9 * value / 1e12,
            "op_breakdown_ms": {f"{op}@{tag}": round(ms, 4) for (op, tag), (ms, n) in top},
            "op_family_ms": {op: round(v[0], 4) for op, v in sorted(by_op.items(), key=lambda kv: -kv[1][0])},
            "eager_step_ms": total_ms,
            "train": train,
        }
        if skin512 is not None:
            line["skin512"] = skin512
        if synapse_dp is not None:
            line["synapse_dp"] = synapse_dp
        if world == 1:
            # host-core baseline and the same-GPU eager bar: rank 0 at N=1 only (under torchrun the other ranks spin and
            # OMP_NUM_THREADS is 1, which made the N>1 numbers of round 1 meaningless)
            cpu_v, cpu_ms, cores, kind = cpu_reference_throughput(sd, kw, 8, 2, 1)
            line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": ("the unmodified reference (baseline/_ref)" if kind == "reference" else "oracle port")
                                              + ", 8 slices/step x 2 steps (1 warm-up) of the same synthetic workload"}
            if train is not None and not args.no_cpu_train:
                from baseline import reference_arms as R
                if R.available():
                    tv, tms, tc = R.cpu_train(TRAIN_CONFIG, 4, 1)
                    kind_t = "reference"
                else:
                    tv, tms, tc = cpu_oracle_train_throughput(2, 1)
                    kind_t = "port"
                line["train"]["cpu_baseline"] = {"value": tv, "unit": "img/s", "cores": tc, "kind": kind_t,
                                                 "sample": "reference module + its Criterion('dice,ce') + torch AdamW, 4 images x 1 step"}
            if not args.no_eager:
                torch.cuda.empty_cache()
                eg = eager_subprocess(args.steps, args.warmup)
                line["eager"] = eg
                try:
                    bi, bt = eg["infer"]["best"], eg["train"]["best"]
                    line["vs_eager"] = {"value": value / bi["value"], "e2e": e2e / bi["e2e"], "infer_mode": bi["mode"],
                                        "infer_batch": bi["batch"],
                                        "train": (train["value"] / bt["value"]) if train else None, "train_mode": bt["mode"],
                                        "vs_stock_fp32": {"value": value / eg["infer"]["fp32"]["value"],
                                                          "train": (train["value"] / eg["train"]["fp32"]["value"]) if train else None},
                                        "note": "product / the reference module in PyTorch eager on the same B200 (best of the eager modes)"}
                except Exception:
                    pass
        print(json.dumps(line))
        if train is not None:                                     # second line (stderr): the collective path's own metric
            print(json.dumps({"metric": train["metric"], "value": train["value"], "unit": train["unit"], "n_gpus": world,
                              "ms_per_step": train["ms_per_step"], "scaling": "weak", "higher_is_better": True}), file=sys.stderr)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference", "reference-gpu"])
    ap.add_argument("--no-eager", action="store_true", help="skip the reference-in-eager-on-this-GPU bar (N=1)")
    ap.add_argument("--no-skin512", action="store_true", help="skip the 512x512 skin legs (BASELINE configs[3], N=1)")
    ap.add_argument("--eager-modes", default="", help="comma list of eager modes for --impl reference-gpu")
    ap.add_argument("--no-train", action="store_true", help="skip the training leg (BASELINE configs[2])")
    ap.add_argument("--no-cpu-train", action="store_true", help="skip the CPU training baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
