#!/usr/bin/env python
"""bench.py -- CENet hot-path throughput on B200 (BASELINE.json metric: 224^2 slices/s, batched inference).

    python bench.py --gpus 1 --steps 20 --warmup 5             # product arm (hand-written sm_100a kernels)
    python bench.py --impl reference --steps 3 --warmup 1      # reference arm: the CPU oracle port on host cores
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
  cpu_baseline : the oracle (CPU port of the reference) on this box's host cores, bounded sample, rank 0 only.
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
NCU_TRAFFIC = {"source": "profiles/r1_ncu_top_kernels.md, profiles/r1_gemm_traffic.md", "diffattn_flash_kernel": 193.4e6,
               "gemm_tc_kernel": 39.67e6}     # mean over the 154 launches of one forward


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, sd, kw = build_model()
    sample = 4
    v, ms, cores = cpu_oracle_throughput(sd, kw, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CENet Synapse 9-class slice inference {SIZE}x{SIZE}, CPU oracle port, {sample} slices/step",
                   "batch_per_step": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} slices/step x {args.steps} steps of the same synthetic Synapse workload"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
def run_train_leg(args, dev, world, rank, flush):
    """BASELINE.json configs[2]: ACDC 4-class training step (Dice+CE 0.5/0.5, AdamW lr 1e-4 wd 1e-4), 224x224, batch 24 per
    GPU, train-mode BatchNorm + DropPath, bf16 activations / fp32 master weights.  A step = forward, fused loss, backward,
    gradient all-reduce (N > 1), AdamW -- all hand-written kernels, replayed from CUDA graph segments."""
    import contextlib
    import torch
    from cenet_b200 import ops, replicas
    from cenet_b200.networks import CENet
    from oracle import fixtures
    kw = fixtures.CONFIGS[TRAIN_CONFIG]
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):
        m = CENet(**kw)
    m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
    m = m.to(dev).train()
    eng = m.train_engine(dev)
    sync = replicas.GradSync(eng) if world > 1 else None
    x_host = fixtures.synth_input(TRAIN_CONFIG, TRAIN_BATCH, SIZE, seed=100 + rank).pin_memory()
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
        "config": {"workload": f"CENet (PVTv2-b2) ACDC 4-class training step {SIZE}x{SIZE}, batch {TRAIN_BATCH} per GPU, Dice+CE, AdamW, "
                               "train-mode BatchNorm, DropPath on", "batch_per_gpu": TRAIN_BATCH, "l2": "flushed between steps",
                   "cuda_graph": bool(eng.use_graph), "flops_per_image": 76.0e9,
                   "parallelism": f"dp{world} (gradient all-reduce in 6 buckets overlapped with backward)" if world > 1 else "dp1"},
        "dtype": "bf16", "launches_per_step": int(eng.launches_per_step or 0),
        "gpu_launches": int(ops.launch_count() - n0), "final_loss": float(loss_host[0]),
        "model_tflops": 76.0e9 * TRAIN_BATCH * world / (t_dev / args.steps / 1e3) / 1e12,
    }


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
    train = None
    if not args.no_train:
        train = run_train_leg(args, dev, world, rank, flush)
    clocks.__exit__()

    if rank == 0:
        peaks = _peaks()
        value = replicas.job_throughput(BATCH, args.steps, t_dev_ms)
        e2e = replicas.job_throughput(BATCH, args.steps, t_e2e_ms)
        # ---- roofline of the dominant kernel, timed live (eager launches bracketed by CUDA events) ----
        prof = eng.profile_ops(x_dev, labels=True, steps=2)
        total_ms = sum(v[0] for v in prof.values())
        by_op = {}
        for (op, tag), (ms, n) in prof.items():
            a = by_op.setdefault(op, [0.0, 0]); a[0] += ms; a[1] += n
        top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]
        # (a) the dominant kernel of the step: gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv; every nn.Linear, 1x1,
        #     3x3 and 5x5 conv) -- all its launches of one forward, algorithmic bytes (each operand once) over the
        #     summed launch durations.  Mostly K <= 128 shapes -> HBM-bound; the tensor-pipe view is stated next to it.
        gw = eng.last_gemm_work
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
        # (b) the largest single launch: differential flash attention of the 56x56 DSE block (exp-bound)
        E1, N1 = 128, (SIZE // 4) ** 2
        ms_da, n_da = prof.get(("diffattn_flash", "se1"), (None, 0))
        roof_attn = None
        if ms_da:
            fl = diffattn_flops(N1, E1, BATCH)
            ach = fl / (ms_da / 1e3) / 1e12
            roof_attn = {"bound": "tensor", "kernel": "diffattn_flash_kernel<8,16> (DSEB 56x56, 16 softmax maps, head_dim 8)",
                         "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                         "traffic": NCU_TRAFFIC.get("diffattn_flash_kernel"), "ms_per_launch": ms_da,
                         "share_of_step": ms_da / total_ms, "peak_source": peaks["source"] + ", sustained bf16",
                         "mufu_bound": {"exp_per_launch": 16.0 * N1 * N1 * BATCH, "peak_exp_per_s": 148 * 16 * 1.965e9,
                                        "frac": 16.0 * N1 * N1 * BATCH / (ms_da / 1e3) / (148 * 16 * 1.965e9)},
                         "note": "softmax-bound: one MUFU.EX2 per score (16/clk/SM); see DESIGN.md 4a"}
        _, _, cores = 0, 0, os.cpu_count()
        cpu_v, cpu_ms, cores = cpu_oracle_throughput(sd, kw, 4, 2, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"CENet (PVTv2-b2) Synapse 9-class batched slice inference {SIZE}x{SIZE}, batch {BATCH} per GPU, "
                                   "logits -> fused softmax/argmax int64 labels", "batch_per_gpu": BATCH, "l2": "flushed between steps",
                       "cuda_graph": bool(eng.use_graph), "parallelism": f"replicas x{world} (batch-sharded, no collective)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": labels_host.numel() * 8, "ms_per_step": t_e2e_ms / args.steps},
            "gpu_launches": int(launches_per_step * args.steps * 2 + launches_per_step * 2) + (train["gpu_launches"] if train else 0),
            "launches_per_step": int(launches_per_step),
            "clocks": clocks.summary(),
            "roofline": roof,
            "roofline_attention": roof_attn,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "oracle port, 4 slices/step x 2 steps (1 warm-up) of the same synthetic workload"},
            "op_breakdown_ms": {f"{op}@{tag}": round(ms, 4) for (op, tag), (ms, n) in top},
            "op_family_ms": {op: round(v[0], 4) for op, v in sorted(by_op.items(), key=lambda kv: -kv[1][0])},
            "eager_step_ms": total_ms,
            "train": train,
        }
        if train is not None and not args.no_cpu_train:
            tv, tms, tc = cpu_oracle_train_throughput(2, 1)
            line["train"]["cpu_baseline"] = {"value": tv, "unit": "img/s", "cores": tc, "kind": "port",
                                             "sample": "oracle port + autograd + torch AdamW, 2 images/step x 1 step"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-train", action="store_true", help="skip the training leg (BASELINE configs[2])")
    ap.add_argument("--no-cpu-train", action="store_true", help="skip the CPU training baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
