"""Launch plan of one CENet TRAINING step on one B200 (train-mode BatchNorm, DropPath, hand-written backward).

Host-side mirror of the reference training iteration (main_acdc.py:234-265):
    outputs = net(x); loss = Criterion('dice,ce')(outputs, y); loss.backward(); AdamW.step()
for `CENet.forward` (net.py:53-64) in `train()` mode.  Nothing here is autograd: the forward issues the kernels of
cenet_b200/csrc through the C ABI and records, per launch group, the closure that launches its hand-written backward
kernels; `backward()` replays the closures in reverse.  All launches use static buffers, so a whole step
(forward, fused Dice+CE, backward, AdamW) is replayed from one CUDA graph.

Differences from the inference plan (engine.py):
  * BatchNorm uses batch statistics (53 BN2d + 4 BN1d) and updates the module's running stats, so it cannot be folded
    into GEMM weights: every BN is `bn_stats` (two-stage column reduction -> per-channel scale/shift) followed by an
    `affine_act` pass (or a consumer that takes per-channel scale/shift).
  * Every activation needed by a backward kernel is kept in its own buffer (no in-place residual stream).
  * Parameters live in ONE flat fp32 buffer (module parameters are views into it), gradients in a second flat buffer:
    the optimizer is one kernel and the gradient all-reduce is a few contiguous buckets.

Gradient bookkeeping: `G(t)` is the gradient buffer of activation `t` (same shape/dtype); `wr(t)` says whether it
already holds a contribution in this backward pass (-> the kernel accumulates instead of overwriting).  Residual
streams alias the gradient buffers of consecutive versions (`alias_grad`), which makes `d(x + f(x)) = dx + f'(x)`
implicit.

Precision: bf16 activations/gradients with fp32 statistics, fp32 parameter gradients, fp32 master weights.
precision="fp32" keeps everything fp32 with CUDA-core GEMMs and materialised attention (validation).
"""
from __future__ import annotations

import math
import os

import torch

from . import ops as _ops_mod
from . import train_ops as _tops_mod
from .engine import _MCA_RATES, _rup, lambda_init
from .ops import ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, GEMM_AUTO, GEMM_MMA, GEMM_SIMT
from .train_ops import ACT_GELU_GRAD


def _storage_ranges(objs, out):
    """[lo, hi) device address ranges of the storages behind every tensor found in `objs` (recursing into containers)"""
    for o in objs:
        if isinstance(o, torch.Tensor):
            if o.is_cuda:
                st = o.untyped_storage()
                out.append((st.data_ptr(), st.data_ptr() + st.nbytes()))
        elif isinstance(o, dict):
            _storage_ranges(o.values(), out)
        elif isinstance(o, (list, tuple)):
            _storage_ranges(o, out)
    return out


class _WgradStream:
    """Second CUDA stream for the weight-gradient kernels of the backward pass.

    At batch 24 most backward GEMMs occupy a fraction of the 148 SMs; d(weight) of a layer is needed by nobody until the
    gradient bucket closes, while d(input) is on the critical path.  `run(reads, fn)` launches `fn` (the wgrad kernels)
    on the side stream, ordered after everything already enqueued on the launching stream, and remembers the storages it
    READS.  Every later launch on the main stream goes through `guard`: if one of its tensors lives in a storage a
    pending side launch still reads (e.g. a residual-stream gradient buffer that is about to be accumulated into), the
    main stream first waits for that side launch.  The check is conservative (a main-stream READ of the same storage
    also waits).  Outputs of the side launches (parameter gradients, the side workspace) are disjoint from everything
    the main stream touches before `join`, which the engine calls before a gradient bucket is handed to the all-reduce
    and before the optimizer.  Inside CUDA graph capture the event record/wait pairs become fork/join edges."""

    def __init__(self, dev):
        self.stream = torch.cuda.Stream(dev)
        self.pending = []                       # [(done event, [(lo, hi), ...])] in launch order
        self.waits = self.launches = 0

    def run(self, reads, fn):
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self.stream.wait_event(ev)
        _hazards.cur = None                     # the guard must not see the side launches themselves
        try:
            with torch.cuda.stream(self.stream):
                fn()
                done = torch.cuda.Event()
                done.record(self.stream)
        finally:
            _hazards.cur = self
        self.pending.append((done, _storage_ranges(reads, [])))
        self.launches += 1

    def guard(self, args, kwargs):
        mine = _storage_ranges(kwargs.values(), _storage_ranges(args, []))
        last = -1
        for i, (_, rs) in enumerate(self.pending):
            if any(lo < b and a < hi for (lo, hi) in rs for (a, b) in mine):
                last = i
        if last >= 0:
            torch.cuda.current_stream().wait_event(self.pending[last][0])       # in-order stream: covers 0..last
            del self.pending[:last + 1]
            self.waits += 1

    def join(self):
        if self.pending:
            torch.cuda.current_stream().wait_event(self.pending[-1][0])
            self.pending.clear()


class _Hazards:
    cur = None                                  # the _WgradStream of the backward pass in flight (None: nothing to check)


_hazards = _Hazards()


class _Guarded:
    """Proxy over an ops module: every kernel launch first passes `_WgradStream.guard` (see there)."""
    _PLAIN = ("ccu_nchunk", "loss_nblocks", "launch_count", "dt", "make_tables")

    def __init__(self, inner):
        object.__setattr__(self, "_inner", inner)

    def __setattr__(self, name, value):
        setattr(self._inner, name, value)

    def __getattr__(self, name):
        fn = getattr(self._inner, name)
        if not callable(fn) or name in _Guarded._PLAIN:
            return fn

        def launch(*a, **k):
            side = _hazards.cur
            if side is not None and side.pending:
                side.guard(a, k)
            return fn(*a, **k)
        return launch


ops = _Guarded(_ops_mod)
tops = _Guarded(_tops_mod)

FLASH_DIMS = {(8, 16), (16, 32), (32, 64), (64, 64), (128, 128), (80, 160)}       # (dqk, dv) instantiated in train_attn.cu


class TrainEngine:
    def __init__(self, module, device, precision="bf16"):
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        from . import _lib
        _lib.load()
        self.mod = module
        self.dev = torch.device(device)
        self.precision = precision
        self.T = torch.bfloat16 if precision == "bf16" else torch.float32
        self.gemm_impl = GEMM_AUTO if precision == "bf16" else GEMM_SIMT
        # materialised attention (batched / transposed operands): mma.sync tensor-core GEMM in bf16, CUDA cores in fp32 validation
        self.attn_gemm_impl = GEMM_MMA if precision == "bf16" else GEMM_SIMT
        self.use_flash = precision == "bf16" and os.environ.get("CENET_B200_ATTN", "flash") == "flash"
        self.use_diffattn_tc = os.environ.get("CENET_B200_DIFFATTN_TC", "1") != "0"     # 0: per-map mma.sync flash forward
        self.use_graph = os.environ.get("CENET_B200_GRAPH", "1") == "1"
        self.cfg = module.cfg
        self.pvt = module.backbone.pvt_cfg                    # widths / depths / ratios of the PVTv2 variant (pvtv2.py:385-431)
        self.drop_path = True                 # stochastic depth of the encoder (pvtv2.py:146-147); tests switch it off
        self.momentum = 0.1
        self.w = {}                           # packed (compute-dtype / re-laid-out) weights
        self._bufs = {}
        self._g = {}                          # activation data_ptr -> gradient buffer
        self._galias = {}
        self._written = set()
        self.tape = TrainEngine._Tape()
        self._graphs = {}
        # eval()-mode semantics inside a gradient pass (networks.CENet._EvalForward): BatchNorm uses its running statistics and
        # updates nothing, DropPath is the identity; the plan runs eagerly (no graphs are cached for this rare path)
        self.frozen_stats = False
        self._dp_counter = self._dp_seed = None                 # DropPath generator state (device counter, seed)
        self._wg_pending, self._wg_off, self._wg_flush_idx, self._wg_tables = [], 0, 0, {}   # deferred wgrad reductions
        self.taps = None
        self.launches_per_step = None
        self._ab_graphs = {}                  # (B,H,W) -> forward / backward CUDA graphs of the autograd-boundary path
        self._pack_maps = None                # [(int32 index map, flat packed buffer)] once built (see pack)
        self._trace = None
        # weight-gradient kernels on a second stream (see _WgradStream); CENET_B200_WGRAD_STREAM=0 keeps one stream
        self.side = (_WgradStream(self.dev) if self.dev.type == "cuda" and
                     os.environ.get("CENET_B200_WGRAD_STREAM", "1") == "1" else None)
        self._flatten()

    # ------------------------------------------------------------------------------------------------ parameters
    def _flatten(self):
        """Move every parameter into one flat fp32 buffer (the module's Parameters become views) + a flat grad buffer."""
        # flat order = groups in the order their gradients complete LAST-to-FIRST reversed: encoder stage 1..4, decoder,
        # head.  The backward pass finishes them head -> decoder -> stage 4 .. 1, so each group is one contiguous
        # all-reduce bucket that can be launched as soon as its backward closures have run (see _bucket_ready).
        def group(n):
            if n.startswith("backbone."):
                for s_ in range(1, 5):
                    if n.startswith((f"backbone.patch_embed{s_}.", f"backbone.block{s_}.", f"backbone.norm{s_}.")):
                        return s_ - 1
                raise KeyError(n)
            return 4 if n.startswith("decoder.") else 5
        named = sorted(((n, p) for n, p in self.mod.named_parameters()), key=lambda np_: group(np_[0]))
        offs, tot = {}, 0
        self.bucket_ranges = {}
        for n, p in named:
            offs[n] = tot
            g_ = group(n)
            tot += _rup(p.numel(), 4)                                   # 16-byte aligned starts
            lo, _hi = self.bucket_ranges.get(g_, (offs[n], 0))
            self.bucket_ranges[g_] = (lo, tot)
        self.pflat = torch.zeros(tot, device=self.dev, dtype=torch.float32)
        self.gflat = torch.zeros(tot, device=self.dev, dtype=torch.float32)
        self.P, self.GP = {}, {}
        for n, p in named:
            v = self.pflat[offs[n]:offs[n] + p.numel()].view(p.shape)
            v.copy_(p.detach().to(self.dev, torch.float32))
            p.data = v
            self.P[n] = v
            self.GP[n] = self.gflat[offs[n]:offs[n] + p.numel()].view(p.shape)
        self.param_offsets = offs
        self.n_flat = tot
        self.BUF = {n: b for n, b in self.mod.named_buffers()}
        for n, b in self.BUF.items():
            if b.device != self.dev:
                raise RuntimeError(f"buffer {n} is on {b.device}; move the module to {self.dev} first")
        self.adam_m = torch.zeros_like(self.pflat)
        self.adam_v = torch.zeros_like(self.pflat)
        self.step_count = 0
        self.zero32 = torch.zeros(32, device=self.dev, dtype=torch.float32)
        probs = self.mod.backbone.drop_path_probs
        self.dp_keep = (1.0 - torch.tensor(probs, dtype=torch.float32).repeat_interleave(2).view(2 * len(probs), 1)).to(self.dev)
        # optimizer hyper-parameters live on the device so that a captured step can be replayed with a new lr:
        # [lr, beta1, beta2, eps, weight_decay, step]
        self.hyper = torch.zeros(8, device=self.dev, dtype=torch.float32)

    def attach_grads(self):
        """Expose the flat gradient buffer as `param.grad` views (drop-in for torch optimizers / DDP-style tools)."""
        for n, p in self.mod.named_parameters():
            p.grad = self.GP[n]

    # ------------------------------------------------------------------------------------------------ packing
    def _put(self, name, t, dtype=None):
        if self._trace is not None:                       # index-tracing pass of _build_pack_maps
            self._trace[name] = t.to(torch.float64).contiguous()
            return None
        t = t.to(self.dev, dtype or torch.float32).contiguous()
        old = self.w.get(name)
        if old is not None and old.shape == t.shape and old.dtype == t.dtype:
            old.copy_(t)
        else:
            self.w[name] = t
        return self.w[name]

    def _pad_cols(self, m, mult=8):
        N, K = m.shape
        Kp = _rup(K, mult)
        if Kp == K:
            return m
        o = torch.zeros(N, Kp, device=m.device, dtype=m.dtype)
        o[:, :K] = m
        return o

    def _pack_mat(self, name, w2d):
        """GEMM weight [N,K]: `.w` [N,Kp] for y = x W^T, `.wT` [K,Np] for dx = dy W."""
        self._put(name + ".w", self._pad_cols(w2d), self.T)
        self._put(name + ".wT", self._pad_cols(w2d.t()), self.T)

    def _pack_conv(self, name, w4d, im2col=False):
        """dense conv weight [N,Cin,KH,KW]: `.w` [N,(kh,kw,ci)]; `.wT` [Cin,(kh',kw',n)] with flipped taps (dgrad as a conv);
        im2col convs also get `.wTp` [(kh,kw,ci) (padded), N] for dcol = dy W."""
        N, Cin, KH, KW = w4d.shape
        m = w4d.permute(0, 2, 3, 1).reshape(N, -1)
        self._put(name + ".w", self._pad_cols(m), self.T)
        if im2col:
            mp = self._pad_cols(m)                                               # [N, Kp]
            self._put(name + ".wTp", self._pad_cols(mp.t()), self.T)              # [Kp, Np]
        else:
            f = w4d.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, -1)              # [Cin, (kh',kw',n)]
            self._put(name + ".wT", self._pad_cols(f), self.T)

    def _pack_dw(self, name, w4d):
        Cc = w4d.shape[0]
        m = w4d.reshape(Cc, 9).t()
        self._put(name + ".w9", m)
        self._put(name + ".w9f", m.flip(0))

    def pack(self):
        """Compute-dtype copies of the master weights (once per optimizer step).

        The layouts are defined by the torch expressions of `_pack_torch` (cast, transpose, tap permutation, flip, zero
        padding, concatenation).  On the GPU they are evaluated ONCE: `_build_pack_maps` replays the same expressions on
        a tensor of element indices, which yields for every packed element the master-weight element it copies; from then
        on a step re-packs with one `gather_cast` launch per dtype instead of ~600 small torch kernels."""
        if self._pack_maps is None:
            self._pack_torch()
            if self.dev.type == "cuda" and os.environ.get("CENET_B200_PACK_GATHER", "1") == "1":
                self._build_pack_maps()
            return
        for idx, dst in self._pack_maps:
            tops.gather_cast(self.pflat, idx, dst)
        self._pack_stem()

    def _pack_stem(self):
        """patch_embed1: with one input channel the three replicated channels fold into SUMMED filters (not a gather)"""
        w = self.P["backbone.patch_embed1.proj.weight"]
        if self.cfg["input_channels"] == 1:
            w = w.sum(1, keepdim=True)                                           # cat([x,x,x]) == summed filters
        self._pack_conv("backbone.patch_embed1.proj", w, im2col=True)

    def _build_pack_maps(self):
        real_P, n = self.P, self.n_flat
        ids = torch.arange(1, n + 1, dtype=torch.float64, device=self.dev)          # 0 is reserved for "zero padding"
        self.P = {k: ids[self.param_offsets[k]:self.param_offsets[k] + v.numel()].view(v.shape) for k, v in real_P.items()}
        self._trace = {}
        try:
            self._pack_torch(stem=False)
        finally:
            traced, self._trace, self.P = self._trace, None, real_P
        maps = []
        for dtype in (torch.bfloat16, torch.float32):
            names = [k for k in traced if self.w[k].dtype == dtype]
            if not names:
                continue
            offs, tot = {}, 0
            for k in names:
                offs[k] = tot
                tot += _rup(self.w[k].numel(), 64)
            dst = torch.zeros(tot, device=self.dev, dtype=dtype)
            idx = torch.zeros(tot, device=self.dev, dtype=torch.int32)
            for k in names:
                old, t = self.w[k], traced[k]
                if t.shape != old.shape:
                    raise RuntimeError(f"pack map of {k}: traced shape {tuple(t.shape)} != packed shape {tuple(old.shape)}")
                view = dst[offs[k]:offs[k] + old.numel()].view(old.shape)
                view.copy_(old)
                idx[offs[k]:offs[k] + old.numel()] = t.flatten().round().to(torch.int32)
                self.w[k] = view
            want = dst.clone()
            tops.gather_cast(self.pflat, idx, dst)
            if not torch.equal(want, dst):
                raise RuntimeError("cenet_b200: the gathered weight pack differs from the torch-evaluated layouts")
            maps.append((idx, dst))
        self._pack_maps = maps

    def _pack_torch(self, stem=True):
        P, cfg = self.P, self.cfg
        if stem:
            self._pack_stem()
        for s in range(4):
            pe = f"backbone.patch_embed{s+1}"
            if s > 0:
                self._pack_conv(pe + ".proj", P[pe + ".proj.weight"], im2col=True)
            for i in range(self.pvt["depths"][s]):
                b = f"backbone.block{s+1}.{i}"
                for n in ("attn.q", "attn.kv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                    self._pack_mat(f"{b}.{n}", P[f"{b}.{n}.weight"])
                if self.pvt["sr_ratios"][s] > 1:
                    self._pack_conv(f"{b}.attn.sr", P[f"{b}.attn.sr.weight"], im2col=True)
                self._pack_dw(f"{b}.mlp.dwconv.dwconv", P[f"{b}.mlp.dwconv.dwconv.weight"])
        for name, Cc in (("dec4", 512), ("dec3", 320), ("dec2", 128), ("dec1", 64)):
            p = f"decoder.{name}"
            m = p + ".mca"
            self._pack_mat(m + ".gate", P[m + ".gate.weight"].flatten(1))
            v = m + ".value"
            for i in range(3):
                d = f"{v}.dlps.{i}"
                self._pack_dw(d + ".depthwise", P[d + ".depthwise.weight"])
            # the three pointwise convs as ONE block-diagonal GEMM [Cc, 3*ap] over the padded depthwise buffer (zero rows for the
            # pooled slice): slice widths 20 / 40 / 100 / 160 are not multiples of 8, which put three forward GEMMs per block on
            # the scalar epilogue, their dgrads on the CUDA-core GEMM and their weight gradients on the mma.sync kernel
            from .networks.cenet import channel_slices
            sl_ = channel_slices(Cc)
            a_ = sl_[0][1] - sl_[0][0]
            ap_ = _rup(a_, 8)
            w0 = P[f"{v}.dlps.0.pointwise.weight"]
            bd = torch.zeros(Cc, 3 * ap_, device=w0.device, dtype=w0.dtype)
            for i in range(3):
                bd[sl_[i][0]:sl_[i][1], i * ap_:i * ap_ + a_] = P[f"{v}.dlps.{i}.pointwise.weight"].flatten(1)
            self._pack_mat(v + ".dlps_bd", bd)
            self._pack_mat(f"{v}.dlps.3.1", P[f"{v}.dlps.3.1.weight"].flatten(1))
            self._pack_mat(v + ".PW_conv", P[v + ".PW_conv.weight"].flatten(1))
            self._pack_mat(m + ".proj_2", P[m + ".proj_2.weight"].flatten(1))
            n = m + ".denoising_module"
            self._pack_mat(n + ".tpg", torch.cat([P[n + ".conv_theta.weight"], P[n + ".conv_phi.weight"],
                                                  P[n + ".conv_g.weight"]], 0).flatten(1))
            self._put(n + ".tpg.b", torch.cat([P[n + ".conv_theta.bias"], P[n + ".conv_phi.bias"], P[n + ".conv_g.bias"]], 0))
            self._pack_mat(n + ".conv_out", P[n + ".conv_out.weight"].flatten(1))
            q = p + ".mlp"
            self._pack_mat(q + ".fc1", P[q + ".fc1.weight"].flatten(1))
            self._pack_dw(q + ".dwconv", P[q + ".dwconv.weight"])
            self._pack_mat(q + ".fc2", P[q + ".fc2.weight"].flatten(1))
        for lvl in (3, 2, 1):
            self._pack_up(f"decoder.up{lvl}", cfg["dec_up_block"])
            d = f"decoder.skip_enhancer{lvl}.diffattn"
            self._pack_mat(d + ".qkv", torch.cat([P[d + ".q_proj.weight"], P[d + ".k_proj.weight"], P[d + ".v_proj.weight"]], 0))
            self._pack_mat(d + ".out_proj", P[d + ".out_proj.weight"])
            self._pack_mat(f"decoder.skip_enhancer{lvl}.mixer", P[f"decoder.skip_enhancer{lvl}.mixer.weight"].flatten(1))
        # head
        Cin = cfg["input_channels"]
        w1 = P["out.rb.0.conv1.conv.weight"]
        self._put("out.rb.0.stem.w1", w1.permute(0, 2, 3, 1).reshape(w1.shape[0], -1))
        self._put("out.rb.0.stem.w3", P["out.rb.0.conv3.conv.weight"].flatten(1))
        self._pack_conv("out.rb.0.conv2.conv", P["out.rb.0.conv2.conv.weight"])
        self._pack_conv("out.out.0.conv1.conv", P["out.out.0.conv1.conv.weight"])
        self._pack_conv("out.out.0.conv2.conv", P["out.out.0.conv2.conv.weight"])
        self._pack_up("out.up", cfg["out_up_block"])
        self._pack_mat("out.out.1.conv.conv", P["out.out.1.conv.conv.weight"].flatten(1))

    def _pack_up(self, p, kind):
        P = self.P
        if kind == "eucb":
            self._pack_dw(p + ".up_dwc.1", P[p + ".up_dwc.1.weight"])
            self._pack_mat(p + ".pwc.0", P[p + ".pwc.0.weight"].flatten(1))
        elif kind == "upcn":
            self._pack_conv(p + ".up.1", P[p + ".up.1.weight"])
        elif kind == "uptc":                                        # blocks.py:223-243: ConvTranspose2d(k, stride 2) as zero insertion +
            w = P[p + ".up.conv.weight"]                            # the stride-1 conv with the transposed, tap-flipped filter
            self._pack_conv(p + ".up.conv", w.permute(1, 0, 2, 3).flip(2, 3))
        elif kind == "uprb":                                        # blocks.py:188-204: bilinear x2 + UnetResBlock(k=3)
            q = p + ".up.1"
            self._pack_conv(q + ".conv1.conv", P[q + ".conv1.conv.weight"])
            self._pack_conv(q + ".conv2.conv", P[q + ".conv2.conv.weight"])
            if (q + ".conv3.conv.weight") in P:
                self._pack_mat(q + ".conv3.conv", P[q + ".conv3.conv.weight"].flatten(1))
        else:
            raise NotImplementedError(kind)

    # ------------------------------------------------------------------------------------------------ buffers / grads
    def buf(self, key, shape, dtype=None, zero=False):
        dtype = dtype or self.T
        k = (self._plan_key, key)
        t = self._bufs.get(k)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = self._bufs[k] = (torch.zeros if zero else torch.empty)(shape, device=self.dev, dtype=dtype)
        return t

    def _gkey(self, t):
        k = t.data_ptr()
        return self._galias.get(k, k)

    def G(self, t):
        """gradient buffer of activation buffer `t` (keyed by its storage address)"""
        k = self._gkey(t)
        g = self._g.get((self._plan_key, k))
        if g is None or g.numel() != t.numel() or g.dtype != t.dtype:
            g = self._g[(self._plan_key, k)] = torch.zeros(t.numel(), device=t.device, dtype=t.dtype)
        return g.view(t.shape)

    def alias_grad(self, t_new, t_old):
        """d(t_new) and d(t_old) share one buffer (t_new = t_old + branch)."""
        self._galias[t_new.data_ptr()] = self._gkey(t_old)

    def wr(self, t, off=0):
        """True if G(t) (column slice starting at `off`) already holds a gradient in this pass (the caller must then
        accumulate); marks it written."""
        k = (self._gkey(t), off)
        had = k in self._written
        self._written.add(k)
        return had

    def _tap(self, name, t, B, H, W, Cc):
        if self.taps is not None:
            self.taps[name] = t.reshape(B, H, W, Cc).permute(0, 3, 1, 2).float().clone()

    def _ws(self, n):
        """scratch for two-stage reductions (consumed inside the launching op, so one shared buffer is enough)"""
        return self.buf("ws.reduce", (max(n, 1 << 24),), torch.float32)

    def _split_ws(self):
        """fp32 scratch the forward GEMMs may use for split-K (few output tiles, long contraction: the SR convs)"""
        return self.buf("ws.splitk", (1 << 22,), torch.float32) if self.T == torch.bfloat16 else None

    def _wws(self):
        """workspace of the weight-gradient kernels: their own buffer when they run on the side stream"""
        return self._ws(0) if self.side is None else self.buf("ws.wgrad", (1 << 24,), torch.float32)

    def _wgrad(self, reads, fn):
        """launch weight-gradient kernels `fn` (which read the tensors `reads`) off the critical path"""
        if self.side is None or _hazards.cur is not self.side:
            fn()
        else:
            self.side.run(reads, fn)

    # Weight-gradient GEMMs leave split partials in `ws.wgrad`; ONE batched launch per gradient bucket reduces them all
    # (autograd: one AccumulateGrad per parameter; the first version of this engine: a finalize launch per GEMM).
    def _wg_ws(self):
        return self.buf("ws.wgrad.big", (1 << (27 if self.dev.type == "cuda" else 22),), torch.float32)

    def _wg_gemm(self, dy, x, dw, **kw):
        """tops.gemm_wgrad with the reduction deferred to the next `_wg_flush`; call inside `_wgrad` (side stream)"""
        ws = self._wg_ws()
        nk = kw["N"] * (kw["K"] + 1)
        if self._wg_off + 32 * nk > ws.numel():
            self._wg_flush()
        jobs, used = tops.gemm_wgrad_partial(dy, x, dw, ws=ws[self._wg_off:], **kw)
        self._wg_pending += jobs
        self._wg_off += _rup(used, 4)

    def _wg_flush(self):
        jobs, self._wg_pending, self._wg_off = self._wg_pending, [], 0
        if not jobs:
            return
        key = (self._plan_key, self._wg_flush_idx)
        self._wg_flush_idx += 1
        ent = self._wg_tables.get(key)
        if ent is None or ent[0] != jobs:
            if self.dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("weight-gradient reduction table changed between the warm-up step and the graph capture")
            tab, nj, nb = tops.wgrad_reduce_table(jobs)
            ent = self._wg_tables[key] = (jobs, tab.to(self.dev), nj, nb)
        # (registered as a READER of the workspace: a later main-stream kernel that writes partials into it waits for this launch)
        self._wgrad((self._wg_ws(),), lambda: tops.wgrad_reduce_batch(ent[1], ent[2], ent[3]))

    # ------------------------------------------------------------------------------------------------ primitives
    def lin(self, x, name, out, *, M=None, N=None, K=None, lda=None, a_off=0, ldc=None, c_off=0, bias=None,
            wname=None, drop=None, drop_div=1, res1=None, dmul=None, x_needs_grad=True, wgrads=None):
        """out = [res1 +] drop * (x W^T + b);  W = parameter `name`.weight ([N,K] or 1x1 conv).
        wgrads: list of (param name, row0, rows) when the packed weight is a concatenation of several parameters."""
        wn = wname or name
        W, WT = self.w[wn + ".w"], self.w[wn + ".wT"]
        M = x.shape[0] if M is None else M
        K = x.shape[1] if K is None else K
        N = out.shape[1] if N is None else N
        lda = x.stride(0) if lda is None else lda
        ldc = out.stride(0) if ldc is None else ldc
        if bias is None:
            bias = self.P.get(name + ".bias")
        elif bias is False:
            bias = None
        ops.gemm(x, W, out, M=M, N=N, K=K, lda=lda, ldw=W.stride(0), ldc=ldc, bias=bias, a_off=a_off, c_off=c_off,
                 post_rs=drop, post_rs_div=drop_div, res1=res1, ldr1=res1.stride(0) if res1 is not None else 0,
                 impl=self.gemm_impl, split_ws=self._split_ws())
        if res1 is not None:
            self.alias_grad(out, res1)
        if wgrads is None:
            wgrads = [(name, 0, N)]
        Kr = WT.shape[0]                                   # real K (x may carry zero pad columns up to K)
        assert dmul is None or a_off == 0

        def bwd():
            dy = self.G(out)
            if x_needs_grad:
                tgt = dmul if dmul is not None else x
                dx = self.G(tgt)
                acc = self.wr(tgt, a_off)
                Wf = self.P.get(wn + ".weight")
                if (self.T == torch.bfloat16 and dy.dtype == torch.float32 and N <= 16 and Kr % 8 == 0 and drop is None
                        and dmul is None and a_off == 0 and c_off == 0 and ldc == N and Wf is not None and Wf.numel() == N * Kr):
                    # fp32 gradient of a layer with a handful of outputs (the class logits): K = N_classes is too short for a
                    # GEMM tile -- one element-wise pass with the fp32 master weight in shared memory
                    tops.smallk_dgrad(dy, Wf, dx, rows=M, K=N, N=Kr, ldw=Kr, ldx=lda, acc=acc)
                else:
                    ops.gemm(dy, WT, dx, M=M, N=Kr, K=N, lda=ldc, ldw=WT.stride(0), ldc=lda, a_off=c_off, c_off=a_off,
                             row_scale=drop, rs_div=drop_div, mul=dmul, ldmul=dmul.stride(0) if dmul is not None else 0,
                             mul_act=ACT_GELU_GRAD if dmul is not None else ACT_NONE,
                             res1=dx if acc else None, ldr1=lda, r1_off=a_off, impl=self.gemm_impl)

            def wg():
                for pn, r0, nr in wgrads:
                    gb = self.GP.get(pn + ".bias") if (bias is not None) else None
                    self._wg_gemm(dy, x, self.GP[pn + ".weight"], M=M, N=nr, K=Kr, ldy=ldc, y_off=c_off + r0, ldx=lda,
                                  x_off=a_off, row_scale=drop, rs_div=drop_div, rs_binary=drop is not None, dbias=gb)
            self._wgrad((dy, x), wg)            # (drop is written only before the forward)
        self.tape.append(bwd)
        return out

    def _pw_blockdiag(self, dwb, catraw, v, sl, a, ap, M, Cc):
        """catraw[:, slice i] = dwb[:, i*ap : i*ap+a] W_i^T for the three dilated branches of MultiOrderDWConv (cfam.py:227-241) as
        one block-diagonal GEMM; the weight gradient is one tcgen05 launch whose reduction jobs pick the diagonal blocks."""
        W, WT = self.w[v + ".dlps_bd.w"], self.w[v + ".dlps_bd.wT"]
        ops.gemm(dwb, W, catraw, M=M, N=Cc, K=3 * ap, lda=3 * ap, ldw=W.stride(0), ldc=Cc, impl=self.gemm_impl)

        def bwd():
            dy = self.G(catraw)                       # (columns of the pooled slice are never written: zeros)
            dx = self.G(dwb)
            acc = self.wr(dwb)
            ops.gemm(dy, WT, dx, M=M, N=3 * ap, K=Cc, lda=Cc, ldw=WT.stride(0), ldc=3 * ap, res1=dx if acc else None,
                     ldr1=3 * ap, impl=self.gemm_impl)
            blocks = [(self.GP[f"{v}.dlps.{i}.pointwise.weight"], sl[i][0], a, i * ap, a) for i in range(3)]

            def wg():
                ws = self._wg_ws()
                if self._wg_off + 32 * Cc * 3 * ap > ws.numel():
                    self._wg_flush()
                jobs, used = tops.gemm_wgrad_blocks(dy, dwb, blocks, M=M, N=Cc, K=3 * ap, ldy=Cc, y_off=0, ldx=3 * ap, x_off=0,
                                                    ws=ws[self._wg_off:])
                self._wg_pending += jobs
                self._wg_off += _rup(used, 4)
            self._wgrad((dy, dwb), wg)
        self.tape.append(bwd)

    def ln(self, x, name, out, eps):
        g, b = self.P[name + ".weight"], self.P[name + ".bias"]
        ops.layernorm(x, out, g, b, eps)

        def bwd():
            dx = self.G(x)
            acc = self.wr(x)
            if self.dev.type == "cuda" and self.side is not None and _hazards.cur is self.side:
                # d(gamma) / d(beta) stay as partial rows in the weight-gradient workspace: the per-bucket batched reduction on the
                # side stream finalises them (53 finalize launches off the critical path).  The guard orders this main-stream
                # writer after a pending reduction that still reads the workspace (`_wg_flush` registers it as a reader).
                ws = self._wg_ws()
                if self._wg_off + 16 * 2 * x.shape[1] * 1184 > ws.numel():
                    self._wg_flush()
                jobs, used = tops.layernorm_bwd(self.G(out), x, g, eps, dx, acc, self.GP[name + ".weight"], self.GP[name + ".bias"],
                                                ws[self._wg_off:], defer=True)
                self._wg_pending += jobs
                self._wg_off += _rup(used, 4)
            else:
                tops.layernorm_bwd(self.G(out), x, g, eps, dx, acc, self.GP[name + ".weight"], self.GP[name + ".bias"],
                                   self._ws(0))
        self.tape.append(bwd)
        return out

    def bn_stats(self, x, rows, Cc, name, key, ld=None, off=0):
        """train-mode BatchNorm statistics of x[rows, C] (pitch ld, column offset off) -> per-channel scale/shift; updates
        running_mean / running_var / num_batches_tracked (momentum 0.1, unbiased running variance)."""
        ld = Cc if ld is None else ld
        st = dict(scale=self.buf(key + ".bn_s", (Cc,), torch.float32), shift=self.buf(key + ".bn_t", (Cc,), torch.float32),
                  mean=self.buf(key + ".bn_m", (Cc,), torch.float32), rstd=self.buf(key + ".bn_r", (Cc,), torch.float32),
                  rows=rows, C=Cc, name=name)
        tops.bn_stats(x, rows, Cc, self.P[name + ".weight"], self.P[name + ".bias"], self.BUF[name + ".running_mean"],
                      self.BUF[name + ".running_var"], self.BUF[name + ".num_batches_tracked"], self.momentum, 1e-5,
                      st["scale"], st["shift"], st["mean"], st["rstd"], self._ws(0), ldx=ld, x_off=off, frozen=self.frozen_stats)
        return st

    def bn_bwd(self, st, dy, y, a, *, act=ACT_NONE, slope=0.0, ldy=None, y_off=0, lda=None, a_off=0, dres=None,
               lddres=None, dres_off=0, dres_acc=False, da=None, da_acc=None):
        """launch the BN backward of y = act(BN(a) [+ other]) given dy; writes d(a), d(gamma), d(beta) [and dres = dy*act']"""
        Cc, rows, name = st["C"], st["rows"], st["name"]
        da_t = self.G(a) if da is None else da
        acc = self.wr(a, a_off) if da_acc is None else da_acc
        tops.bn_bwd(dy, y, a, st["mean"], st["rstd"], self.P[name + ".weight"], rows, Cc, da_t, self.GP[name + ".weight"],
                    self.GP[name + ".bias"], self._ws(0), act=act, slope=slope, acc_da=acc, dres=dres, acc_dres=dres_acc,
                    ldy=Cc if ldy is None else ldy, y_off=y_off, lda=Cc if lda is None else lda, a_off=a_off,
                    lddres=lddres or Cc, dres_off=dres_off, frozen=self.frozen_stats)

    def bn_act(self, x, rows, Cc, name, out, key, *, act=ACT_NONE, slope=0.0, ldx=None, x_off=0, ldo=None, o_off=0):
        """out = act(BN_train(x)) with its backward."""
        st = self.bn_stats(x, rows, Cc, name, key, ld=ldx, off=x_off)
        tops.affine_act(x, out, rows, Cc, sa=st["scale"], ta=st["shift"], act=act, slope=slope, lda=ldx, a_off=x_off,
                        ldo=ldo, o_off=o_off)

        def bwd():
            self.bn_bwd(st, self.G(out), out, x, act=act, slope=slope, ldy=ldo, y_off=o_off, lda=ldx, a_off=x_off)
        self.tape.append(bwd)
        return st

    def dwconv(self, x, out, name, B, H, W, Cc, *, bias=None, act=ACT_NONE, zout=None, dil=1, up2=False, ldx=None,
               x_off=0, ldy=None, y_off=0):
        """depthwise 3x3 (+bias)(+GELU with the pre-activation kept in zout); H, W are output sizes."""
        w9, w9f = self.w[name + ".w9"], self.w[name + ".w9f"]
        ops.dwconv3x3(x, out, w9, B, H, W, Cc, ldx=ldx, ldy=ldy, x_off=x_off, y_off=y_off, bias=bias, dil=dil, up2=up2,
                      act=act, zout=zout)
        pre = zout if zout is not None else out        # buffer whose gradient is d(pre-activation)
        ldp, p_off = (Cc, 0) if zout is not None else (ldy or Cc, y_off)

        def bwd():
            dz = self.G(pre)
            gb = self.GP[name + ".bias"] if bias is not None else None
            dx = self.G(x)
            if up2:
                full = self.buf("tmp.dwup." + name, (B * H * W, Cc))
                ops.dwconv3x3(dz, full, w9f, B, H, W, Cc, ldx=ldp, x_off=p_off, dil=dil)
                acc = self.wr(x)
                tops.sumpool2(full, dx, B, H // 2, W // 2, Cc, acc)
            else:
                self.wr(x, x_off)                      # single consumer per channel slice: plain write
                ops.dwconv3x3(dz, dx, w9f, B, H, W, Cc, ldx=ldp, x_off=p_off, ldy=ldx, y_off=x_off, dil=dil)
            self._wgrad((x, dz), lambda: tops.dwconv3x3_wgrad(x, dz, self.GP[name + ".weight"], gb, B, H, W, Cc, dil, up2,
                                                              ldx or Cc, x_off, ldp, p_off, self._wws()))
        self.tape.append(bwd)
        return out

    def conv(self, x4, name, out, k, key, gw=None, gw_fix=None):
        """dense stride-1 'same' conv without bias: out[M,Cout] raw.  x4: [B,H,W,Cin] contiguous.
        gw / gw_fix: weight-gradient destination [Cout,Cin,k,k] other than the parameter's own gradient + a fix-up run after it."""
        B, H, W, Cin = x4.shape
        Wm, WT = self.w[name + ".w"], self.w[name + ".wT"]
        Cout = out.shape[-1]
        ops.conv_nhwc(x4, Wm, out, k, 1, k // 2, impl=self.gemm_impl)

        def bwd():
            dy = self.G(out)
            dx = self.G(x4)
            acc = self.wr(x4)
            ops.conv_nhwc(dy.view(B, H, W, Cout), WT, dx.view(B * H * W, Cin), k, 1, k // 2, res1=dx if acc else None,
                          ldr1=Cin, impl=self.gemm_impl)
            dst = self.GP[name + ".weight"] if gw is None else gw
            if self.T == torch.bfloat16:                                                # taps gathered inside the kernel
                def wg():
                    tops.conv_wgrad(dy, x4, dst, k, self._wws())
                    if gw_fix is not None:
                        gw_fix()
                self._wgrad((dy, x4), wg)
            else:
                Kp = _rup(k * k * Cin, 8)
                col = self.buf(key + ".col", (B * H * W, Kp))
                ops.im2col(x4, col, B, H, W, Cin, k, 1, k // 2, H, W, Kp)
                tops.gemm_wgrad(dy, col, dst, M=B * H * W, N=Cout, K=k * k * Cin, ldy=Cout, y_off=0,
                                ldx=Kp, x_off=0, T=k * k, ws=self._ws(0))
                if gw_fix is not None:
                    gw_fix()
        self.tape.append(bwd)
        return out

    def conv_im2col(self, x, B, H, W, Cin, k, stride, pad, name, out, key, x_needs_grad=True, wgrad_fix=None):
        """strided / non-overlapping conv through im2col + GEMM (patch embeds pvtv2.py:164-165, SR conv pvtv2.py:68)"""
        Ho = (H + 2 * pad - k) // stride + 1
        Wo = (W + 2 * pad - k) // stride + 1
        Wm = self.w[name + ".w"]
        Kp = Wm.shape[1]
        Cout = out.shape[-1]
        M = B * Ho * Wo
        col = self.buf(key + ".col", (M, Kp))
        ops.im2col(x, col, B, H, W, Cin, k, stride, pad, Ho, Wo, Kp)
        ops.gemm(col, Wm, out, M=M, N=Cout, K=Kp, lda=Kp, ldw=Kp, ldc=Cout, bias=self.P[name + ".bias"], impl=self.gemm_impl,
                 split_ws=self._split_ws())

        def bwd():
            dy = self.G(out)
            gw = self.GP[name + ".weight"]
            tgt = gw if wgrad_fix is None else self.buf(key + ".gw1", (Cout, Cin, k, k), torch.float32)

            def wg():
                kw = dict(M=M, N=Cout, K=k * k * Cin, ldy=Cout, y_off=0, ldx=Kp, x_off=0, T=k * k, dbias=self.GP[name + ".bias"])
                if wgrad_fix is None:
                    self._wg_gemm(dy, col, tgt, **kw)
                else:                                                 # the fix-up reads the finished gradient right away
                    tops.gemm_wgrad(dy, col, tgt, ws=self._wws(), **kw)
                    wgrad_fix(tgt, gw)
            self._wgrad((dy, col), wg)
            if x_needs_grad:
                WTp = self.w[name + ".wTp"]
                dcol = self.buf(key + ".dcol", (M, Kp))
                ops.gemm(dy, WTp, dcol, M=M, N=Kp, K=Cout, lda=Cout, ldw=WTp.stride(0), ldc=Kp, impl=self.gemm_impl)
                dx = self.G(x)
                acc = self.wr(x)
                tops.col2im(dcol, dx, B, H, W, Cin, k, stride, pad, Ho, Wo, Kp, acc)
        self.tape.append(bwd)
        return Ho, Wo

    # ---- attention ----------------------------------------------------------------------------------------------
    def attention(self, Q, Km, V, O, *, B, maps, Nq, Nk, dqk, dv, vdiv, scale, key, ldq, qo, ldk, ko, ldv, vo, ldo, oo,
                  diff_qkv=False):
        """O[:, oo + m*dv : +dv] = softmax(scale * Q_m K_m^T) V_{m // vdiv} for m < maps, per image.
        Q_m = Q[:, qo + m*dqk : +dqk] etc.; all operands are row-major token matrices, images are consecutive row blocks."""
        flash = self.use_flash and (dqk, dv) in FLASH_DIMS
        if flash:
            lse = self.buf(key + ".lse", (B * maps * Nq,), torch.float32)
            if diff_qkv and dqk in tops.DIFFATTN_TC_HEAD_DIMS and Q.dtype == torch.bfloat16 and self.use_diffattn_tc:
                # the differential attention's own layout: all 2h maps on the tcgen05 kernel (diffattn_tc.cu, training epilogue)
                tops.diffattn_fwd_train(Q, O, lse, B, Nq, maps * dqk, maps // 2,
                                        self.buf(key + ".kmax", (B * maps,), torch.float32))
            else:
                tops.flash_fwd(Q, Km, V, O, lse, B, maps, Nq, Nk, dqk, dv, vdiv, scale, ldq, qo, ldk, ko, ldv, vo, ldo, oo)

            def bwd():
                dO = self.G(O)
                delta = self.buf(key + ".delta", (B * maps * Nq,), torch.float32)
                tops.flash_bwd(Q, Km, V, O, dO, lse, delta, self.G(Q), self.G(Km), self.G(V), B, maps, Nq, Nk, dqk, dv, vdiv,
                               scale, ldq, qo, ldk, ko, ldv, vo, ldo, oo, ws=self._ws(0))
                for t in (Q, Km, V):
                    self.wr(t)
            self.tape.append(bwd)
            return
        hv = maps // vdiv
        S = self.buf(key + ".S", (B * maps, Nq, Nk))
        ops.gemm(Q, Km, S, M=Nq, N=Nk, K=dqk, lda=ldq, ldw=ldk, ldc=Nk, alpha=scale, batch=B * maps, batch_inner=maps,
                 a_bs=(Nq * ldq, dqk), w_bs=(Nk * ldk, dqk), c_bs=(maps * Nq * Nk, Nq * Nk), a_off=qo, w_off=ko, impl=self.attn_gemm_impl)
        ops.softmax_rows_(S, B * maps * Nq, Nk, Nk)
        for j in range(vdiv):
            ops.gemm(S, V, O, M=Nq, N=dv, K=Nk, lda=Nk, ldw=ldv, ldc=ldo, batch=B * hv, batch_inner=hv,
                     a_bs=(maps * Nq * Nk, vdiv * Nq * Nk), a_off=j * Nq * Nk, w_bs=(Nk * ldv, dv), w_off=vo, w_nmajor=True,
                     c_bs=(Nq * ldo, vdiv * dv), c_off=oo + j * dv, impl=self.attn_gemm_impl)

        def bwd():
            dO = self.G(O)
            dQ, dK, dV = self.G(Q), self.G(Km), self.G(V)
            dP = self.buf(key + ".dP", (B * maps, Nq, Nk))
            for j in range(vdiv):
                ops.gemm(dO, V, dP, M=Nq, N=Nk, K=dv, lda=ldo, ldw=ldv, ldc=Nk, batch=B * hv, batch_inner=hv,
                         a_bs=(Nq * ldo, vdiv * dv), a_off=oo + j * dv, w_bs=(Nk * ldv, dv), w_off=vo,
                         c_bs=(maps * Nq * Nk, vdiv * Nq * Nk), c_off=j * Nq * Nk, impl=self.attn_gemm_impl)
                # dV_h (+)= P_m^T dO_m
                ops.gemm(S, dO, dV, M=Nk, N=dv, K=Nq, lda=Nk, a_mmajor=True, ldw=ldo, w_nmajor=True, ldc=ldv, batch=B * hv,
                         batch_inner=hv, a_bs=(maps * Nq * Nk, vdiv * Nq * Nk), a_off=j * Nq * Nk, w_bs=(Nq * ldo, vdiv * dv),
                         w_off=oo + j * dv, c_bs=(Nk * ldv, dv), c_off=vo, res1=dV if j > 0 else None, ldr1=ldv, r1_off=vo,
                         impl=self.attn_gemm_impl)
            tops.softmax_bwd_rows_(S, dP, B * maps * Nq, Nk)
            ops.gemm(dP, Km, dQ, M=Nq, N=dqk, K=Nk, lda=Nk, ldw=ldk, w_nmajor=True, ldc=ldq, alpha=scale, batch=B * maps,
                     batch_inner=maps, a_bs=(maps * Nq * Nk, Nq * Nk), w_bs=(Nk * ldk, dqk), w_off=ko, c_bs=(Nq * ldq, dqk),
                     c_off=qo, impl=self.attn_gemm_impl)
            ops.gemm(dP, Q, dK, M=Nk, N=dqk, K=Nq, lda=Nk, a_mmajor=True, ldw=ldq, w_nmajor=True, ldc=ldk, alpha=scale,
                     batch=B * maps, batch_inner=maps, a_bs=(maps * Nq * Nk, Nq * Nk), w_bs=(Nq * ldq, dqk), w_off=qo,
                     c_bs=(Nk * ldk, dqk), c_off=ko, impl=self.attn_gemm_impl)
            for t in (Q, Km, V):
                self.wr(t)
        self.tape.append(bwd)

    # ------------------------------------------------------------------------------------------------ encoder
    def _encoder(self, x_nhwc, B, H, W, Cin):
        P = self.P
        feats = []
        cur, curC = x_nhwc, Cin
        dp_i = 0
        for s in range(4):
            ops.tag = f"enc{s+1}"
            self._mark_bucket(s)
            Cc, heads, sr = self.pvt["embed_dims"][s], self.pvt["heads"][s], self.pvt["sr_ratios"][s]
            hid = Cc * self.pvt["mlp_ratios"][s]
            k, st = (7, 4) if s == 0 else (3, 2)
            pe = f"backbone.patch_embed{s+1}"
            traw = self.buf(f"enc{s}.traw", (B * ((H + 2 * (k // 2) - k) // st + 1) * ((W + 2 * (k // 2) - k) // st + 1), Cc))
            fix = None
            if s == 0 and Cin == 1:
                def fix(tmp, gw):                                                 # d(sum over the 3 replicated channels)
                    gw.copy_(tmp.expand_as(gw))
            H, W = self.conv_im2col(cur, B, H, W, curC, k, st, k // 2, pe + ".proj", traw, f"enc{s}.pe",
                                    x_needs_grad=(s > 0), wgrad_fix=fix)
            Mtok = B * H * W
            t = self.ln(traw, pe + ".norm", self.buf(f"enc{s}.t0", (Mtok, Cc)), 1e-5)
            Nk = (H // sr) * (W // sr)
            for i in range(self.pvt["depths"][s]):
                b = f"backbone.block{s+1}.{i}"
                kb = f"enc{s}.b{i}"
                drop = drop2 = None                                          # two independent draws per block:
                if self.drop_path and not self.frozen_stats and self.mod.backbone.drop_path_probs[dp_i] > 0:   # (pvtv2.py:146-147)
                    drop, drop2 = self.dp_scale[2 * dp_i], self.dp_scale[2 * dp_i + 1]
                dp_i += 1
                xn = self.ln(t, b + ".norm1", self.buf(kb + ".xn1", (Mtok, Cc)), 1e-6)
                q = self.lin(xn, b + ".attn.q", self.buf(kb + ".q", (Mtok, Cc)))
                kv = self.buf(kb + ".kv", (B * Nk, 2 * Cc))
                if sr > 1:
                    xr = self.buf(kb + ".xr", (B * Nk, Cc))
                    self.conv_im2col(xn, B, H, W, Cc, sr, sr, 0, b + ".attn.sr", xr, kb + ".sr")
                    xrn = self.ln(xr, b + ".attn.norm", self.buf(kb + ".xrn", (B * Nk, Cc)), 1e-5)
                    self.lin(xrn, b + ".attn.kv", kv)
                else:
                    self.lin(xn, b + ".attn.kv", kv)
                att = self.buf(kb + ".att", (Mtok, Cc))
                self.attention(q, kv, kv, att, B=B, maps=heads, Nq=H * W, Nk=Nk, dqk=64, dv=64, vdiv=1, scale=64 ** -0.5,
                               key=kb + ".sra", ldq=Cc, qo=0, ldk=2 * Cc, ko=0, ldv=2 * Cc, vo=Cc, ldo=Cc, oo=0)
                t1 = self.lin(att, b + ".attn.proj", self.buf(kb + ".t1", (Mtok, Cc)), res1=t, drop=drop, drop_div=H * W)
                xn2 = self.ln(t1, b + ".norm2", self.buf(kb + ".xn2", (Mtok, Cc)), 1e-6)
                h1 = self.lin(xn2, b + ".mlp.fc1", self.buf(kb + ".h1", (Mtok, hid)))
                z = self.buf(kb + ".z", (Mtok, hid))
                h2 = self.buf(kb + ".h2", (Mtok, hid))
                self.dwconv(h1, h2, b + ".mlp.dwconv.dwconv", B, H, W, hid, bias=P[b + ".mlp.dwconv.dwconv.bias"], act=ACT_GELU,
                            zout=z)
                t = self.lin(h2, b + ".mlp.fc2", self.buf(kb + ".t2", (Mtok, Cc)), res1=t1, drop=drop2, drop_div=H * W, dmul=z)
            f = self.ln(t, f"backbone.norm{s+1}", self.buf(f"enc{s}.out", (Mtok, Cc)), 1e-6)
            self._tap(f"backbone.stage{s+1}", f, B, H, W, Cc)
            feats.append((f, H, W, Cc))
            cur, curC = f, Cc
        return feats

    # ------------------------------------------------------------------------------------------------ CFAM
    def _cfam(self, x, B, H, W, Cc, p, key):
        """cfam.py:365-374 in train mode"""
        P, GP = self.P, self.GP
        ops.tag = key
        HW, M = H * W, B * H * W
        from .networks.cenet import channel_slices
        sl = channel_slices(Cc)
        m = p + ".mca"
        # ---- norm1 -> xb (MCA input and its shortcut)
        xb = self.buf(key + ".xb", (M, Cc))
        self.bn_act(x, M, Cc, p + ".norm1", xb, key + ".n1")
        # ---- CCU
        u = self.buf(key + ".ccu_u", (B, Cc, 3), torch.float32)
        arg = self.buf(key + ".ccu_arg", (B, Cc), torch.int32)
        tops.ccu_stats(xb, u, arg, B, HW, Cc, self._ws(0))
        gate = self.buf(key + ".ccu_gate", (B, Cc), torch.float32)
        csave = self.buf(key + ".ccu_save", (B, Cc, 8), torch.float32)
        bn1d = B > 1                                                          # cfam.py:260-261
        cb = m + ".ccu.bn"
        tops.ccu_mlp_fwd(u, P[m + ".ccu.fc1.weight"], P[m + ".ccu.fc2.weight"], P[cb + ".weight"] if bn1d else None,
                         P[cb + ".bias"] if bn1d else None, self.BUF[cb + ".running_mean"], self.BUF[cb + ".running_var"],
                         self.BUF[cb + ".num_batches_tracked"], self.momentum, 1e-5, gate, csave, B, Cc, frozen=self.frozen_stats)
        x1 = self.buf(key + ".x1", (M, Cc))
        ops.affine_gate(xb, x1, None, None, gate, B, HW, Cc)

        def ccu_bwd():
            dx1 = self.G(x1)
            dgate = self.buf(key + ".ccu_dgate", (B, Cc), torch.float32)
            tops.ccu_dgate(dx1, xb, dgate, B, HW, Cc, self._ws(0))
            du = self.buf(key + ".ccu_du", (B, Cc, 3), torch.float32)
            tops.ccu_mlp_bwd(dgate, u, P[m + ".ccu.fc1.weight"], P[m + ".ccu.fc2.weight"], P[cb + ".weight"] if bn1d else None,
                             P[cb + ".bias"] if bn1d else None, csave, du, GP[m + ".ccu.fc1.weight"], GP[m + ".ccu.fc2.weight"], GP[cb + ".weight"],
                             GP[cb + ".bias"], B, Cc, frozen=self.frozen_stats)
            dxb = self.G(xb)
            acc = self.wr(xb)
            tops.ccu_apply_bwd(dx1, xb, gate, u, arg, du, dxb, acc, B, HW, Cc)
        self.tape.append(ccu_bwd)
        g = self.lin(x1, m + ".gate", self.buf(key + ".g", (M, Cc)))
        # ---- MultiOrderDWConv(x1)
        v = m + ".value"
        ap = _rup(sl[0][1] - sl[0][0], 8)
        a = sl[0][1] - sl[0][0]
        dwraw = self.buf(key + ".dwraw", (M, 3 * ap), zero=True)
        dwb = self.buf(key + ".dwb", (M, 3 * ap), zero=True)
        catraw = self.buf(key + ".catraw", (M, Cc))
        cat = self.buf(key + ".cat", (M, Cc))

        def x1_slices_done():          # (runs AFTER the slice backward closures below): all channel slices of dx1 written
            self.wr(x1)
        self.tape.append(x1_slices_done)
        for i, rate in enumerate(_MCA_RATES[Cc]):
            a0, a1 = sl[i]
            d = f"{v}.dlps.{i}"
            self.dwconv(x1, dwraw, d + ".depthwise", B, H, W, a, dil=rate, ldx=Cc, x_off=a0, ldy=3 * ap, y_off=i * ap)
            self.bn_act(dwraw, M, a, d + ".depthwise_bn", dwb, f"{key}.dl{i}.dbn", act=ACT_RELU, ldx=3 * ap, x_off=i * ap,
                        ldo=3 * ap, o_off=i * ap)
        self._pw_blockdiag(dwb, catraw, v, sl, a, ap, M, Cc)
        for i in range(3):
            a0, a1 = sl[i]
            d = f"{v}.dlps.{i}"
            self.bn_act(catraw, M, a, d + ".pointwise_bn", cat, f"{key}.dl{i}.pbn", act=ACT_RELU, ldx=Cc, x_off=a0, ldo=Cc,
                        o_off=a0)
        # image-pooling branch: AdaptiveAvgPool(7) -> 1x1 -> BN -> LeakyReLU(0.01) -> bilinear x7 (align) -> bilinear (H,W)
        r0, r1 = sl[3]
        r = r1 - r0
        d = f"{v}.dlps.3"
        pooled = self.buf(key + ".pooled", (B * 49, r))
        tb = self._pool_tables(H, W)
        tops.resample(x1, pooled, B, H, W, 7, 7, r, tb["pool"], ldx=Cc, x_off=r0, ldy=r, y_off=0, acc=False)
        praw = self.buf(key + ".poolraw", (B * 49, r))
        pact = self.buf(key + ".poolact", (B * 49, r))

        def pool_in_bwd():
            tops.resample(self.G(pooled), self.G(x1), B, 7, 7, H, W, r, tb["pool_T"], ldx=r, x_off=0, ldy=Cc, y_off=r0, acc=False)
        self.tape.append(pool_in_bwd)
        self.lin(pooled, d + ".1", praw, bias=False)
        self.bn_act(praw, B * 49, r, d + ".2", pact, key + ".poolbn", act=ACT_LEAKY, slope=0.01)
        tops.resample(pact, cat, B, 7, 7, H, W, r, tb["up"], ldx=r, x_off=0, ldy=Cc, y_off=r0, acc=False)

        def pool_out_bwd():
            self.wr(pact)
            tops.resample(self.G(cat), self.G(pact), B, H, W, 7, 7, r, tb["up_T"], ldx=Cc, x_off=r0, ldy=r, y_off=0, acc=False)
        self.tape.append(pool_out_bwd)
        vv = self.lin(cat, v + ".PW_conv", self.buf(key + ".v", (M, Cc)))
        sv = self.buf(key + ".sv", (M, Cc))
        tops.silu_mul_fwd(g, vv, sv, M * Cc)

        def silu_bwd():
            tops.silu_mul_bwd(self.G(sv), g, vv, self.G(g), self.G(vv), M * Cc)
            self.wr(g); self.wr(vv)
        self.tape.append(silu_bwd)
        y = self.lin(sv, m + ".proj_2", self.buf(key + ".y", (M, Cc)), res1=xb)            # + shortcut BN(x)
        # ---- Nonlocal(y) -> xa = x + ls1 * ((1-w) y + w BN(conv_out(att)))
        n = m + ".denoising_module"
        tpg = self.lin(y, n + ".tpg", self.buf(key + ".tpg", (M, 3 * Cc)), bias=self.w[n + ".tpg.b"], wname=n + ".tpg",
                       wgrads=[(n + ".conv_theta", 0, Cc), (n + ".conv_phi", Cc, Cc), (n + ".conv_g", 2 * Cc, Cc)])
        att = self.buf(key + ".nl.att", (M, Cc))
        self.attention(tpg, tpg, tpg, att, B=B, maps=1, Nq=HW, Nk=HW, dqk=Cc, dv=Cc, vdiv=1, scale=Cc ** -0.5, key=key + ".nl",
                       ldq=3 * Cc, qo=0, ldk=3 * Cc, ko=Cc, ldv=3 * Cc, vo=2 * Cc, ldo=Cc, oo=0)
        nraw = self.lin(att, n + ".conv_out", self.buf(key + ".nl.raw", (M, Cc)))
        st = self.bn_stats(nraw, M, Cc, n + ".bn", key + ".nl.bn")
        xa = self.buf(key + ".xa", (M, Cc))
        ls1 = P[p + ".layer_scale_1"]
        tops.ls_combine_fwd(x, y, nraw, st["scale"], st["shift"], ls1, P[n + ".w"], xa, M, Cc)
        self.alias_grad(xa, x)

        def mix_bwd():
            dxa = self.G(xa)                               # == G(x): d(x) already holds d(xa)
            dy = self.G(y)
            dpz = self.buf(key + ".nl.dpz", (M, Cc))
            tops.ls_combine_bwd(dxa, y, nraw, st["scale"], st["shift"], ls1, P[n + ".w"], dy, self.wr(y), dpz,
                                GP[p + ".layer_scale_1"], GP[n + ".w"], M, Cc, self._ws(0))
            self.bn_bwd(st, dpz, None, nraw)
        self.tape.append(mix_bwd)
        # ---- Mlp
        q = p + ".mlp"
        xb2 = self.buf(key + ".xb2", (M, Cc))
        self.bn_act(xa, M, Cc, p + ".norm2", xb2, key + ".n2")
        h1 = self.lin(xb2, q + ".fc1", self.buf(key + ".h1", (M, 4 * Cc)))
        z = self.buf(key + ".z", (M, 4 * Cc))
        h2 = self.buf(key + ".h2", (M, 4 * Cc))
        self.dwconv(h1, h2, q + ".dwconv", B, H, W, 4 * Cc, bias=P[q + ".dwconv.bias"], act=ACT_GELU, zout=z)
        su = self.buf(key + ".srm_u", (M, 3), torch.float32)
        sarg = self.buf(key + ".srm_arg", (M,), torch.int32)
        tops.row_stats_arg(h2, su, sarg, M, 4 * Cc)
        gm = self.buf(key + ".srm_g", (M,), torch.float32)
        ssave = self.buf(key + ".srm_save", (M, 2), torch.float32)            # pre-GELU f, BN-normalised value
        sst = self.buf(key + ".srm_st", (4,), torch.float32)                  # mean, rstd of the 1-channel BN
        sb = q + ".srm.bn"
        tops.srm_fwd(su, P[q + ".srm.pwc.weight"], P[q + ".srm.dwc.weight"], P[sb + ".weight"], P[sb + ".bias"],
                     self.BUF[sb + ".running_mean"], self.BUF[sb + ".running_var"], self.BUF[sb + ".num_batches_tracked"],
                     self.momentum, 1e-5, gm, ssave, sst, B, H, W, self._ws(0), frozen=self.frozen_stats)
        mo = self.buf(key + ".mo", (M, Cc))
        W2, W2T = self.w[q + ".fc2.w"], self.w[q + ".fc2.wT"]
        ops.gemm(h2, W2, mo, M=M, N=Cc, K=4 * Cc, lda=4 * Cc, ldw=W2.stride(0), ldc=Cc, bias=P[q + ".fc2.bias"], row_scale=gm,
                 impl=self.gemm_impl)

        def fc2_srm_bwd():
            dmo = self.G(mo)
            dh3 = self.buf(key + ".dh3", (M, 4 * Cc))
            ops.gemm(dmo, W2T, dh3, M=M, N=4 * Cc, K=Cc, lda=Cc, ldw=W2T.stride(0), ldc=4 * Cc, impl=self.gemm_impl)
            if self.T == torch.bfloat16:
                # mo = gm * (h2 W^T) + b: fold the per-pixel gate into d(mo) once, then the weight gradient is a plain tcgen05 GEMM
                # (a per-row scale kept it on the mma.sync kernel) and the bias gradient a column sum of the unscaled d(mo)
                dmo_s = self.buf(key + ".dmo_s", (M, Cc))

                def wg():
                    tops.row_scale(dmo, gm, dmo_s, M, Cc)
                    self._wg_gemm(dmo_s, h2, GP[q + ".fc2.weight"], M=M, N=Cc, K=4 * Cc, ldy=Cc, y_off=0, ldx=4 * Cc, x_off=0)
                    tops.colsum(dmo, GP[q + ".fc2.bias"], rows=M, C=Cc, ld=Cc, ws=self._wws())
                self._wgrad((dmo, gm, h2), wg)
            else:
                tops.gemm_wgrad(dmo, h2, GP[q + ".fc2.weight"], M=M, N=Cc, K=4 * Cc, ldy=Cc, y_off=0, ldx=4 * Cc, x_off=0,
                                row_scale=gm, dbias=GP[q + ".fc2.bias"], bias_unscaled=True, ws=self._ws(0))
            dgm = self.buf(key + ".srm_dg", (M,), torch.float32)
            tops.row_dot(dh3, h2, dgm, M, 4 * Cc)
            dsu = self.buf(key + ".srm_du", (M, 3), torch.float32)
            tops.srm_bwd(dgm, su, gm, ssave, sst, P[q + ".srm.pwc.weight"], P[q + ".srm.dwc.weight"], P[sb + ".weight"],
                         P[sb + ".bias"], dsu,
                         GP[q + ".srm.pwc.weight"], GP[q + ".srm.dwc.weight"], GP[sb + ".weight"], GP[sb + ".bias"], B, H, W,
                         self._ws(0), frozen=self.frozen_stats)
            # d(h2) = dh3 * gm + statistics terms, then * gelu'(z) -> d(z)
            tops.srm_apply_bwd(dh3, h2, z, gm, su, sarg, dsu, self.G(z), M, 4 * Cc)
            self.wr(z)
        self.tape.append(fc2_srm_bwd)
        out = self.buf(key + ".out", (M, Cc))
        ls2 = P[p + ".layer_scale_2"]
        tops.ls_combine_fwd(xa, None, mo, None, None, ls2, None, out, M, Cc)
        self.alias_grad(out, xa)

        def ls2_bwd():
            dmo = self.G(mo)
            tops.ls_combine_bwd(self.G(out), None, mo, None, None, ls2, None, None, False, dmo, GP[p + ".layer_scale_2"], None,
                                M, Cc, self._ws(0))
            self.wr(mo)
        self.tape.append(ls2_bwd)
        self._tap(p, out, B, H, W, Cc)
        return out

    _TABLES = {}

    def _pool_tables(self, H, W):
        k = ("pool", H, W, str(self.dev))
        t = TrainEngine._TABLES.get(k)
        if t is None:
            import torch.nn.functional as F

            def up1d(n):
                e = torch.eye(7).view(1, 7, 7, 1)                                   # [1, C=7 (one-hot), 7, 1]
                y = F.interpolate(e, size=(49, 1), mode="bilinear", align_corners=True)
                if n != 49:
                    y = F.interpolate(y, size=(n, 1), mode="bilinear", align_corners=False)
                return y[0, :, :, 0].t().contiguous()                               # [n, 7]

            def pool1d(n):
                e = torch.eye(n).view(1, n, n, 1)
                return F.adaptive_avg_pool2d(e, (7, 1))[0, :, :, 0].t().contiguous()  # [7, n]
            Uh, Uw, Ph, Pw = up1d(H), up1d(W), pool1d(H), pool1d(W)
            t = dict(up=tops.make_tables(Uh, Uw, self.dev), up_T=tops.make_tables(Uh.t(), Uw.t(), self.dev),
                     pool=tops.make_tables(Ph, Pw, self.dev), pool_T=tops.make_tables(Ph.t(), Pw.t(), self.dev))
            TrainEngine._TABLES[k] = t
        return t

    def _zi_tables(self, H, W):
        """2x zero insertion [H,W] -> [2H,2W] (y[2i,2j] = x[i,j]) and its adjoint as resampling tap tables"""
        k = ("zi", H, W, str(self.dev))
        t = TrainEngine._TABLES.get(k)
        if t is None:
            def z(n):
                m = torch.zeros(2 * n, n)
                m[torch.arange(n) * 2, torch.arange(n)] = 1.0
                return m
            Zh, Zw = z(H), z(W)
            t = dict(zi=tops.make_tables(Zh, Zw, self.dev), zi_T=tops.make_tables(Zh.t().contiguous(), Zw.t().contiguous(), self.dev))
            TrainEngine._TABLES[k] = t
        return t

    def _up2_tables(self, H, W):
        k = ("up2ac", H, W, str(self.dev))
        t = TrainEngine._TABLES.get(k)
        if t is None:
            import torch.nn.functional as F

            def up1d(n):
                e = torch.eye(n).view(1, n, n, 1)
                return F.interpolate(e, size=(2 * n, 1), mode="bilinear", align_corners=True)[0, :, :, 0].t().contiguous()
            Uh, Uw = up1d(H), up1d(W)
            t = dict(up_T=tops.make_tables(Uh.t(), Uw.t(), self.dev))
            TrainEngine._TABLES[k] = t
        return t

    # ------------------------------------------------------------------------------------------------ up blocks
    def _up(self, x, B, H, W, Cin, Cout, p, kind, key, out=None, ldc=None, c_off=0):
        ops.tag = key
        Mo = B * 4 * H * W
        if out is None:
            out = self.buf(key + ".out", (Mo, Cout))
            ldc = Cout
        if kind == "uptc":
            t = self.buf(key + ".zi", (B, 2 * H, 2 * W, Cin))
            tb = self._zi_tables(H, W)
            tops.resample(x, t, B, H, W, 2 * H, 2 * W, Cin, tb["zi"], ldx=Cin, x_off=0, ldy=Cin, y_off=0, acc=False)

            def zi_bwd():                                             # adjoint of the zero insertion: dx[i,j] = dt[2i,2j]
                acc = self.wr(x)
                tops.resample(self.G(t), self.G(x), B, 2 * H, 2 * W, H, W, Cin, tb["zi_T"], ldx=Cin, x_off=0, ldy=Cin, y_off=0,
                              acc=acc)
            self.tape.append(zi_bwd)
            name = p + ".up.conv"
            gw = self.GP[name + ".weight"]                            # reference layout [Cin, Cout, k, k]
            ks = gw.shape[-1]
            raw = out if (ldc == Cout and c_off == 0) else self.buf(key + ".raw", (Mo, Cout))
            tmp = self.buf(key + ".gw_eq", (Cout, Cin, ks, ks), torch.float32)

            def fix():                                                # d(equivalent conv filter) -> d(ConvTranspose2d weight)
                gw.copy_(tmp.permute(1, 0, 2, 3).flip(2, 3))
            self.conv(t, name, raw, ks, key + ".conv", gw=tmp, gw_fix=fix)
            if raw is not out:                                        # channel slice of the head's concat buffer
                tops.affine_act(raw, out, Mo, Cout, ldo=ldc, o_off=c_off)

                def copy_bwd():
                    tops.affine_act(self.G(out), self.G(raw), Mo, Cout, lda=ldc, a_off=c_off)
                    self.wr(raw)
                self.tape.append(copy_bwd)
            return out
        if kind == "uprb":
            t = self.buf(key + ".up", (B, 2 * H, 2 * W, Cin))
            ops.upsample2x_ac(x, t, B, H, W, Cin)
            tb = self._up2_tables(H, W)

            def uprb_bwd():
                acc = self.wr(x)
                tops.resample(self.G(t), self.G(x), B, 2 * H, 2 * W, H, W, Cin, tb["up_T"], ldx=Cin, x_off=0, ldy=Cin, y_off=0,
                              acc=acc)
            self.tape.append(uprb_bwd)
            self._resblock(t, B, 2 * H, 2 * W, Cin, Cout, 3, p + ".up.1", key, out, ldc, c_off)
            return out
        if kind == "eucb":
            raw = self.buf(key + ".dwraw", (Mo, Cin))
            self.dwconv(x, raw, p + ".up_dwc.1", B, 2 * H, 2 * W, Cin, up2=True)
            act = self.buf(key + ".dwact", (Mo, Cin))
            self.bn_act(raw, Mo, Cin, p + ".up_dwc.2", act, key + ".bn", act=ACT_LEAKY, slope=0.2)
            self.lin(act, p + ".pwc.0", out, M=Mo, N=Cout, K=Cin, ldc=ldc, c_off=c_off)
        else:
            t = self.buf(key + ".up", (B, 2 * H, 2 * W, Cin))
            ops.upsample2x_ac(x, t, B, H, W, Cin)
            tb = self._up2_tables(H, W)

            def up_bwd():
                acc = self.wr(x)
                tops.resample(self.G(t), self.G(x), B, 2 * H, 2 * W, H, W, Cin, tb["up_T"], ldx=Cin, x_off=0, ldy=Cin, y_off=0,
                              acc=acc)
            self.tape.append(up_bwd)
            craw = self.buf(key + ".craw", (Mo, Cout))
            self.conv(t, p + ".up.1", craw, 3, key + ".conv")
            self.bn_act(craw, Mo, Cout, p + ".up.2", out, key + ".bn", act=ACT_LEAKY, slope=0.2, ldo=ldc, o_off=c_off)
        return out

    # ------------------------------------------------------------------------------------------------ DSEB
    def _dseb(self, skip, dec, B, H, W, Cc, p, heads, depth, key):
        """dseb.py:153-165 (+ the decoder's `d + s`, decoders.py:95): returns mixer(z) + skip + dec"""
        P, GP = self.P, self.GP
        ops.tag = key
        add = self.cfg.get("skip_mode", "cat") == "add"               # dseb.py:155: y = dec + skip instead of cat([dec, skip])
        HW, E, M = H * W, (Cc if add else 2 * Cc), B * H * W
        y = self.buf(key + ".y", (B, E, H, W))
        if add:
            ysum = self.buf(key + ".ysum", (M, Cc))
            tops.add_(ysum, dec, M * Cc, False)
            tops.add_(ysum, skip, M * Cc, True)
            ops.nhwc_to_nchw(ysum, y, B, HW, Cc, E, 0)

            def add_bwd():
                dsum = self.G(ysum)
                tops.nchw_to_nhwc_slice(self.G(y), dsum, B, HW, Cc, E, 0, False)
                for t in (dec, skip):
                    tops.add_(self.G(t), dsum, M * Cc, self.wr(t))
            self.tape.append(add_bwd)
        else:
            ops.nhwc_to_nchw(dec, y, B, HW, Cc, E, 0)
            ops.nhwc_to_nchw(skip, y, B, HW, Cc, E, Cc)

            def cat_bwd():
                dy = self.G(y)
                tops.nchw_to_nhwc_slice(dy, self.G(dec), B, HW, Cc, E, 0, self.wr(dec))
                tops.nchw_to_nhwc_slice(dy, self.G(skip), B, HW, Cc, E, Cc, self.wr(skip))
            self.tape.append(cat_bwd)
        tok = y.view(M, E)
        # ---- differential attention on the reinterpreted buffer
        d = p + ".diffattn"
        hd = E // heads // 2
        li = lambda_init(depth)
        lam = self.buf(key + ".lam", (4,), torch.float32)
        lq1, lk1, lq2, lk2 = (P[d + ".lambda_q1"], P[d + ".lambda_k1"], P[d + ".lambda_q2"], P[d + ".lambda_k2"])
        tops.lambda_fwd(lq1, lk1, lq2, lk2, hd, li, lam)
        qkv = self.lin(tok, d + ".qkv", self.buf(key + ".qkv", (M, 3 * E)), bias=False, wname=d + ".qkv",
                       wgrads=[(d + ".q_proj", 0, E), (d + ".k_proj", E, E), (d + ".v_proj", 2 * E, E)])
        Om = self.buf(key + ".Om", (M, 2 * E))
        self.attention(qkv, qkv, qkv, Om, B=B, maps=2 * heads, Nq=HW, Nk=HW, dqk=hd, dv=2 * hd, vdiv=2, scale=hd ** -0.5,
                       key=key + ".da", ldq=3 * E, qo=0, ldk=3 * E, ko=E, ldv=3 * E, vo=2 * E, ldo=2 * E, oo=0, diff_qkv=True)
        o = self.buf(key + ".o", (M, E))
        tops.diff_rmsnorm_fwd(Om, lam, o, M, heads, 2 * hd, 1e-5, 1.0 - li)

        def dr_bwd():
            dlam = self.buf(key + ".dlam", (4,), torch.float32)
            tops.diff_rmsnorm_bwd(self.G(o), Om, lam, self.G(Om), dlam, M, heads, 2 * hd, 1e-5, 1.0 - li, self._ws(0))
            self.wr(Om)
            tops.lambda_bwd(dlam, lq1, lk1, lq2, lk2, hd, GP[d + ".lambda_q1"], GP[d + ".lambda_k1"], GP[d + ".lambda_q2"],
                            GP[d + ".lambda_k2"])
        self.tape.append(dr_bwd)
        gate = self.lin(o, d + ".out_proj", self.buf(key + ".gate", (M, E)), bias=False)
        # ---- z = 2y + w*edge(y) + gate*y  (FEA + y + diff*y)
        z = self.buf(key + ".z", (B, E, H, W))
        fw = P[p + ".boundary.w"]
        scales = self.cfg["scale_factors"]
        ops.fea_combine(y, gate, z, fw.reshape(-1), B, E, H, W, scales)
        mats = self._fea_mats(H, W, scales)

        def fea_bwd():
            acc = self.wr(y)
            tops.fea_bwd(y, gate, self.G(z), fw.reshape(-1), self.G(y), acc, self.G(gate), GP[p + ".boundary.w"], B, E, H, W,
                         mats, len(scales), self._ws(0), ident_mask=sum(1 << i for i, sf in enumerate(scales) if float(sf) == 1.0))
            self.wr(gate)
        self.tape.append(fea_bwd)
        zt = self.buf(key + ".zt", (M, E))
        ops.nchw_to_nhwc(z, zt, B, HW, E)

        def zt_bwd():
            ops.nhwc_to_nchw(self.G(zt), self.G(z), B, HW, E, E, 0)
            self.wr(z)
        self.tape.append(zt_bwd)
        out = self.buf(key + ".out", (M, Cc))
        Wm = self.w[p + ".mixer.w"]
        ops.gemm(zt, Wm, out, M=M, N=Cc, K=E, lda=E, ldw=Wm.stride(0), ldc=Cc, res1=skip, ldr1=Cc, res2=dec, ldr2=Cc,
                 impl=self.gemm_impl)

        def mix_bwd():
            dout = self.G(out)
            WT = self.w[p + ".mixer.wT"]
            ops.gemm(dout, WT, self.G(zt), M=M, N=E, K=Cc, lda=Cc, ldw=WT.stride(0), ldc=E, impl=self.gemm_impl)
            self.wr(zt)
            self._wgrad((dout, zt), lambda: self._wg_gemm(dout, zt, GP[p + ".mixer.weight"], M=M, N=Cc, K=E, ldy=Cc, y_off=0,
                                                          ldx=E, x_off=0))
            for t in (skip, dec):                                   # residual operands: d += dout
                tops.add_(self.G(t), dout, M * Cc, self.wr(t))
        self.tape.append(mix_bwd)
        if self.taps is not None:
            self.taps[p] = (out.float() - dec.float()).reshape(B, H, W, Cc).permute(0, 3, 1, 2).clone()
        return out

    def _fea_mats(self, H, W, scales):
        """dense per-axis operators A_s = Up_s Down_s (bilinear, align_corners=False) of FEA (dseb.py:40-50), [ns,2,n,n]"""
        k = ("fea", H, W, tuple(scales), str(self.dev))
        t = TrainEngine._TABLES.get(k)
        if t is None:
            import torch.nn.functional as F
            out = []
            for s in scales:
                per_axis = []
                for n in (H, W):
                    e = torch.eye(n).view(1, n, n, 1)                                   # channel = source index
                    dn = F.interpolate(e, scale_factor=(s, 1.0), mode="bilinear")
                    upm = F.interpolate(dn, size=(n, 1), mode="bilinear")[0, :, :, 0].t().contiguous()   # [n_out, n_src]
                    m = torch.zeros(max(H, W), max(H, W))
                    m[:n, :n] = upm
                    per_axis.append(m)
                out.append(torch.stack(per_axis))
            t = torch.stack(out).to(self.dev).contiguous()
            TrainEngine._TABLES[k] = t
        return t

    # ------------------------------------------------------------------------------------------------ head
    def _resblock_tail(self, c2raw, rraw, st3name, out, Mtok, Cc, p, key, ldo=None, o_off=0):
        """out = LeakyReLU(BN2(c2raw) + (BN3(rraw) | rraw)); out may be a channel slice (pitch ldo, offset o_off)"""
        st2 = self.bn_stats(c2raw, Mtok, Cc, p + ".norm2", key + ".bn2")
        st3 = self.bn_stats(rraw, Mtok, Cc, st3name, key + ".bn3") if st3name else None
        tops.affine_act(c2raw, out, Mtok, Cc, sa=st2["scale"], ta=st2["shift"], b=rraw,
                        sb=st3["scale"] if st3 else None, tb=st3["shift"] if st3 else None, act=ACT_LEAKY, slope=0.01,
                        ldo=ldo, o_off=o_off)

        def bwd():
            dy = self.G(out)
            if st3:
                self.bn_bwd(st3, dy, out, rraw, act=ACT_LEAKY, slope=0.01, ldy=ldo, y_off=o_off)
                self.bn_bwd(st2, dy, out, c2raw, act=ACT_LEAKY, slope=0.01, ldy=ldo, y_off=o_off)
            else:
                acc = self.wr(rraw)
                self.bn_bwd(st2, dy, out, c2raw, act=ACT_LEAKY, slope=0.01, ldy=ldo, y_off=o_off, dres=self.G(rraw), dres_acc=acc)
        self.tape.append(bwd)

    def _resblock(self, x4, B, H, W, Cin, Cout, k, p, key, out, ldo=None, o_off=0):
        """UnetResBlock (unet.py:201-214) in train mode: conv k x k -> BN -> LeakyReLU -> conv k x k -> BN; residual through
        1x1 conv + BN when Cin != Cout; add; LeakyReLU(0.01).  x4 [B,H,W,Cin] contiguous."""
        Mtok = B * H * W
        c1raw = self.buf(key + ".c1raw", (Mtok, Cout))
        self.conv(x4, p + ".conv1.conv", c1raw, k, key + ".c1")
        a1 = self.buf(key + ".a1", (B, H, W, Cout))
        self.bn_act(c1raw, Mtok, Cout, p + ".norm1", a1.view(Mtok, Cout), key + ".bn1", act=ACT_LEAKY, slope=0.01)
        c2raw = self.buf(key + ".c2raw", (Mtok, Cout))
        self.conv(a1, p + ".conv2.conv", c2raw, k, key + ".c2")
        x2 = x4.view(Mtok, Cin)
        if (p + ".conv3.conv.weight") in self.P:
            rraw = self.buf(key + ".rraw", (Mtok, Cout))
            self.lin(x2, p + ".conv3.conv", rraw, bias=False)
            self._resblock_tail(c2raw, rraw, p + ".norm3", out, Mtok, Cout, p, key, ldo, o_off)
        else:
            self._resblock_tail(c2raw, x2, None, out, Mtok, Cout, p, key, ldo, o_off)
        return out

    def _run_forward(self, x_in, B, H, W, logits):
        cfg, P, GP = self.cfg, self.P, self.GP
        Cin, ncls = cfg["input_channels"], cfg["num_classes"]
        xc = self.buf("x", (B * H * W, Cin))
        ops.tag = "input"
        if Cin == 1:
            ops.affine_gate(x_in, xc, None, None, None, B, H * W, 1)
        else:
            ops.nchw_to_nhwc(x_in, xc, B, H * W, Cin)
        feats = self._encoder(xc, B, H, W, Cin)
        (x1, H1, W1, C1), (x2, H2, W2, C2), (x3, H3, W3, C3), (x4, H4, W4, C4) = feats
        self._mark_bucket(4)
        d = self._cfam(x4, B, H4, W4, C4, "decoder.dec4", "dec4")
        heads = cfg["diffatt_num_heads"]
        for lvl, (sk, Hs, Ws, Cs, ), hi, Cprev, depth in ((3, feats[2], 0, C4, 4), (2, feats[1], 1, C3, 3), (1, feats[0], 2, C2, 2)):
            up = self._up(d, B, Hs // 2, Ws // 2, Cprev, Cs, f"decoder.up{lvl}", cfg["dec_up_block"], f"up{lvl}")
            self._tap(f"decoder.up{lvl}", up, B, Hs, Ws, Cs)
            xin = self._dseb(sk, up, B, Hs, Ws, Cs, f"decoder.skip_enhancer{lvl}", heads[hi], depth, f"se{lvl}")
            d = self._cfam(xin, B, Hs, Ws, Cs, f"decoder.dec{lvl}", f"dec{lvl}")
        # ---- OutHead (out.py:69-75), train-mode BN
        om = C1 // 2
        Hh, Wh = H // 2, W // 2
        Mf, Mh = B * H * W, B * Hh * Wh
        self._mark_bucket(5)
        ops.tag = "head.rb"
        o1raw = self.buf("head.rb.o1raw", (B, H, W, om))
        rraw = self.buf("head.rb.rraw", (Mf, om))
        ops.stem5x5(xc, self.w["out.rb.0.stem.w1"], self.zero32, self.w["out.rb.0.stem.w3"], self.zero32, o1raw, rraw, B, H, W,
                    Cin, 1.0)

        def stem_bwd():
            Kp = _rup(25 * Cin, 8)
            col = self.buf("head.rb.col", (Mf, Kp))
            do1, drr = self.G(o1raw), self.G(rraw)

            def wg():
                ops.im2col(xc, col, B, H, W, Cin, 5, 1, 2, H, W, Kp)    # (only the weight gradient reads it: off the main stream)
                self._wg_gemm(do1, col, GP["out.rb.0.conv1.conv.weight"], M=Mf, N=om, K=25 * Cin, ldy=om, y_off=0, ldx=Kp,
                              x_off=0, T=25)
                if Cin == 1 and x_in.dtype == torch.float32:
                    # 1x1 conv over ONE input channel: dW[n] = sum_m dY[m, n] * x[m] is a column sum weighted by the image
                    tops.colsum(drr, GP["out.rb.0.conv3.conv.weight"], rows=Mf, C=om, ld=om, row_scale=x_in.view(-1), ws=self._wws())
                else:
                    self._wg_gemm(drr, xc, GP["out.rb.0.conv3.conv.weight"], M=Mf, N=om, K=Cin, ldy=om, y_off=0, ldx=Cin, x_off=0)
            self._wgrad((do1, col, drr, xc, x_in), wg)
        self.tape.append(stem_bwd)
        o1 = self.buf("head.rb.o1", (B, H, W, om))
        self.bn_act(o1raw.view(Mf, om), Mf, om, "out.rb.0.norm1", o1.view(Mf, om), "head.rb.bn1", act=ACT_LEAKY, slope=0.01)
        c2raw = self.buf("head.rb.c2raw", (Mf, om))
        self.conv(o1, "out.rb.0.conv2.conv", c2raw, 5, "head.rb.c2")
        rb = self.buf("head.rb.out", (Mf, om))
        self._resblock_tail(c2raw, rraw, "out.rb.0.norm3", rb, Mf, om, "out.rb.0", "head.rb")
        madd = cfg.get("out_merge_mode", "cat") == "add"            # out.py:58-64: up(dec) + w*rb(x) instead of cat
        mix = om if madd else 2 * om
        z = self.buf("head.z", (Mh, mix))
        if madd:
            rbp = self.buf("head.rbp", (Mh, om))
            ops.maxpool2_scale(rb, rbp, om, 0, P["out.w"].reshape(-1), B, H, W, om)

            def pool_bwd():
                tops.maxpool2_scale_bwd(self.G(rbp), om, 0, rb, P["out.w"].reshape(-1), self.G(rb), GP["out.w"], B, H, W, om,
                                        self._ws(0))
                self.wr(rb)
            self.tape.append(pool_bwd)
            zu = self._up(d, B, H1, W1, C1, om, "out.up", cfg["out_up_block"], "head.up")
            tops.add_(z, zu, Mh * om, False)
            tops.add_(z, rbp, Mh * om, True)

            def merge_bwd():                                          # z = up + rbp: both operands receive dz
                dz = self.G(z)
                tops.add_(self.G(zu), dz, Mh * om, self.wr(zu))
                tops.add_(self.G(rbp), dz, Mh * om, self.wr(rbp))
            self.tape.append(merge_bwd)
        else:
            ops.maxpool2_scale(rb, z, 2 * om, om, P["out.w"].reshape(-1), B, H, W, om)

            def pool_bwd():
                tops.maxpool2_scale_bwd(self.G(z), 2 * om, om, rb, P["out.w"].reshape(-1), self.G(rb), GP["out.w"], B, H, W, om,
                                        self._ws(0))
                self.wr(rb)
            self.tape.append(pool_bwd)
            self._up(d, B, H1, W1, C1, om, "out.up", cfg["out_up_block"], "head.up", out=z, ldc=2 * om, c_off=0)
        ops.tag = "head.out"
        p = "out.out.0"
        z4 = z.view(B, Hh, Wh, mix)
        r1raw = self.buf("head.out.r1raw", (Mh, mix))
        self.conv(z4, p + ".conv1.conv", r1raw, 3, "head.out.c1")
        a1 = self.buf("head.out.a1", (B, Hh, Wh, mix))
        self.bn_act(r1raw, Mh, mix, p + ".norm1", a1.view(Mh, mix), "head.out.bn1", act=ACT_LEAKY, slope=0.01)
        r2raw = self.buf("head.out.r2raw", (Mh, mix))
        self.conv(a1, p + ".conv2.conv", r2raw, 3, "head.out.c2")
        o = self.buf("head.out.o", (Mh, mix))
        self._resblock_tail(r2raw, z, None, o, Mh, mix, p, "head.out")
        ops.tag = "head.logits"
        yh = self.buf("head.y", (Mh, ncls), torch.float32)
        self.lin(o, "out.out.1.conv.conv", yh)
        ops.head_upsample_argmax(yh, logits, None, B, Hh, Wh, ncls)

        def head_up_bwd():
            tops.head_upsample_bwd(self.G(logits), self.G(yh), B, Hh, Wh, ncls)
            self.wr(yh)
        self.tape.append(head_up_bwd)

    # ------------------------------------------------------------------------------------------------ step
    def _begin(self, B, H, W):
        self._plan_key = (B, H, W)
        self.tape = TrainEngine._Tape()
        self._written = set()
        self._galias = {}

    def _sample_drop_path(self, B):
        probs = self.mod.backbone.drop_path_probs
        n = 2 * len(probs)                                          # rows 2i / 2i+1: attention / MLP branch of block i
        ds = self.buf("dp_scale", (n, B), torch.float32)
        if self.drop_path and not self.frozen_stats:
            # timm DropPath: bernoulli(keep) / keep per sample and branch; one kernel of the library with a device-side counter
            # (fresh masks on every CUDA-graph replay); the seed comes from torch's generator, so `torch.manual_seed` still
            # decides the sequence
            if self._dp_counter is None:
                self._dp_counter = torch.zeros(1, device=self.dev, dtype=torch.int64)
                self._dp_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            tops.droppath_mask(ds, self.dp_keep.reshape(-1).contiguous(), n, B, self._dp_seed, self._dp_counter)
        self.dp_scale = ds

    def _touch_weights(self):
        """tell eval-mode engines of the same module that parameters / BatchNorm statistics were rewritten through raw
        pointers (no `_version` bump): Engine._weights_version() reads this counter and re-packs"""
        self.mod._cenet_weights_gen = getattr(self.mod, "_cenet_weights_gen", 0) + 1

    def forward(self, x_in, B, H, W, logits):
        self._begin(B, H, W)
        self._sample_drop_path(B)
        self._run_forward(x_in, B, H, W, logits)

    class _Tape(list):
        """backward closures with the profiling region label that was current when they were recorded"""

        def append(self, fn):
            list.append(self, fn if isinstance(fn, tuple) else (getattr(ops, "tag", ""), fn))

    def _mark_bucket(self, g_):
        """tape marker: when the backward pass reaches it, every parameter gradient of group g_ is final"""
        self.tape.append(("bucket", g_))

    def backward(self, logits):
        """d(logits) must already be in G(logits)."""
        self.wr(logits)
        _hazards.cur = self.side
        self._wg_flush_idx = 0
        try:
            for tag, fn in reversed(self.tape):
                if tag == "bucket":
                    self._wg_flush()                        # one batched reduction of the bucket's weight-gradient partials
                    if self.on_bucket is not None:
                        if self.side is not None:
                            self.side.join()                # the bucket's weight gradients must be final
                        self.on_bucket(fn)
                    continue
                ops.tag = tag + ".bwd"
                fn()
            self._wg_flush()
            if self.side is not None:
                self.side.join()
        finally:
            _hazards.cur = None

    on_bucket = None        # callable(group) -> launches the gradient all-reduce of that bucket (see replicas.py)

    def _step_body(self, x_in, labels, B, H, W, ncls, loss_out, w_dice, w_ce, lr, betas, eps, wd, optimize, w_boundary=0.0):
        self.pack()
        logits = self.buf("logits", (B, ncls, H, W), torch.float32)
        self.forward(x_in, B, H, W, logits)
        ws = self.buf("loss.ws", ((4 * ncls + 1) * ops.loss_nblocks(B * H * W) + 5 * ncls + 4,), torch.float32)
        if w_boundary:                                  # Criterion with a BoundaryDoULoss term (utils/core.py:83-131,161-188)
            ops.seg_loss(logits, labels, loss_out, self.G(logits), ws, B, ncls, H, W, w_dice, w_ce, w_boundary, 1.0)
        else:
            ops.dice_ce(logits, labels, loss_out, self.G(logits), ws, B, ncls, H * W, w_dice, w_ce, 1.0)
        self.backward(logits)
        if self.grad_hook is not None:
            self.grad_hook(self.gflat)
        if optimize:
            for lo, hi in self._trainable_ranges():
                tops.adamw(self.pflat[lo:hi], self.gflat[lo:hi], self.adam_m[lo:hi], self.adam_v[lo:hi], hi - lo, self.hyper)

    def _trainable_ranges(self):
        """contiguous [lo, hi) ranges of the flat parameter buffer whose parameters have requires_grad=True: a frozen
        backbone (`freeze_bb=True`, encoder.py:82-84) is neither updated nor weight-decayed, as with torch.optim over
        `filter(requires_grad)`; all-trainable == one range == one AdamW launch"""
        out = []
        for n, p in sorted(self.mod.named_parameters(), key=lambda np_: self.param_offsets[np_[0]]):
            if not p.requires_grad:
                continue
            lo = self.param_offsets[n]
            hi = lo + _rup(p.numel(), 4)
            if out and out[-1][1] == lo:
                out[-1][1] = hi
            else:
                out.append([lo, hi])
        return [(a, min(b, self.n_flat)) for a, b in out]

    grad_hook = None

    def _capture(self, args):
        """Capture one step as CUDA graph segments.  Without gradient synchronisation it is a single graph; with it
        (on_bucket / grad_hook set by replicas.GradSync) the capture is cut at every bucket marker so that the NCCL
        all-reduces are issued eagerly BETWEEN graph replays and overlap with the next backward segment."""
        segs = []
        pool = torch.cuda.graph_pool_handle()
        cur = {}
        self._user_on_bucket, self._user_grad_hook = self.on_bucket, self.grad_hook

        def begin():
            cur["g"] = torch.cuda.CUDAGraph()
            cur["n0"] = ops.launch_count()
            cur["g"].capture_begin(pool=pool)

        def end():
            if ops.launch_count() == cur["n0"]:                      # two markers back to back (e.g. the last bucket and the
                import warnings                                      # gradient hook): nothing was captured, drop the segment
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    cur["g"].capture_end()
                return
            cur["g"].capture_end()
            segs.append(("graph", cur["g"]))

        def cut_bucket(g_):
            end(); segs.append(("bucket", g_)); begin()

        def cut_hook(_gflat):
            end(); segs.append(("hook", None)); begin()
        if self._user_on_bucket is not None:
            self.on_bucket = cut_bucket
        if self._user_grad_hook is not None:
            self.grad_hook = cut_hook
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                begin()
                self._step_body(*args)
                end()
        finally:
            self.on_bucket, self.grad_hook = self._user_on_bucket, self._user_grad_hook
        torch.cuda.current_stream().wait_stream(side)
        return segs

    # ---- autograd-boundary entry points (used by networks.CENet.forward in train() mode) ----------------------------------
    # `net(x)` ... `loss.backward()` of the unchanged reference loop.  Per input shape: the first call runs eagerly (allocates
    # every buffer), the second call captures the forward (weight re-pack + forward kernels) and, at its backward, the
    # backward kernels as two CUDA graphs; later calls replay them.  Gradient-synchronised replicas (on_bucket / grad_hook)
    # and tap recording stay eager.
    def _ab_graphable(self):
        return (self.use_graph and self.taps is None and self.dev.type == "cuda" and self.on_bucket is None
                and self.grad_hook is None)

    def _ab_capture(self, fn):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            g.capture_begin()
            try:
                fn()
            finally:
                g.capture_end()
        cur.wait_stream(side)
        return g

    def forward_logits(self, x):
        """train-mode forward; returns the engine-owned logits buffer [B,ncls,H,W] (valid until the next call)"""
        if x.device != self.dev:
            raise RuntimeError(f"input on {x.device}, engine on {self.dev}")
        if x.dim() != 4 or x.shape[1] != self.cfg["input_channels"]:
            raise ValueError(f"expected [B,{self.cfg['input_channels']},H,W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 32 or W % 32:
            raise ValueError("H and W must be multiples of 32")
        self._plan_key = (B, H, W)
        x_in = self.buf("x_in", (B, x.shape[1], H, W), torch.float32)
        x_in.copy_(x.detach().float())
        logits = self.buf("logits", (B, self.cfg["num_classes"], H, W), torch.float32)

        def body():
            self.pack()
            self.forward(x_in, B, H, W, logits)
        st = self._ab_graphs.get((B, H, W)) if self._ab_graphable() else None
        if not self._ab_graphable():
            body()
        elif st is None:                                            # first call for this shape: eager warm-up
            body()
            self._ab_graphs[(B, H, W)] = {}
        elif "fwd" not in st:                                       # second call: capture (records the tape), then run it
            torch.cuda.current_stream().synchronize()
            st["fwd"] = self._ab_capture(body)
            st["tape"], st["written"], st["galias"] = self.tape, self._written, self._galias
            st["fwd"].replay()
        else:
            self.tape, self._written, self._galias = st["tape"], st["written"], st["galias"]
            st["fwd"].replay()
        self._last_logits = logits
        if not self.frozen_stats:
            self._touch_weights()                                   # BatchNorm running statistics moved
        return logits

    def backward_from(self, dlogits):
        """run the recorded backward with d(loss)/d(logits); parameter gradients land in self.GP / self.gflat"""
        logits = self._last_logits
        B, H, W = logits.shape[0], logits.shape[2], logits.shape[3]
        self._plan_key = (B, H, W)
        self.G(logits).copy_(dlogits)
        st = self._ab_graphs.get((B, H, W)) if self._ab_graphable() else None
        if st is None or "fwd" not in st:
            self.backward(logits)
            if self.grad_hook is not None:                          # replicas.GradSync.finish: wait for the bucket all-reduces
                self.grad_hook(self.gflat)
        elif "bwd" not in st:
            torch.cuda.current_stream().synchronize()
            st["bwd"] = self._ab_capture(lambda: self.backward(logits))
            st["bwd"].replay()
        else:
            st["bwd"].replay()

    def train_step(self, x, labels, *, w_dice=0.5, w_ce=0.5, w_boundary=0.0, lr=1e-4, betas=(0.9, 0.999), eps=1e-8,
                   weight_decay=1e-4, optimize=True):
        """One fused iteration: forward, Criterion (utils/core.py:179-188: w_dice*Dice + w_ce*CE + w_boundary*BoundaryDoU; the
        ACDC / Synapse scripts run `boundary` alone, the skin script `dice,ce`), backward, AdamW (core.py:16-18).
        Returns the device tensor [1+ncls]: loss, per-class dice.  x [B,Cin,H,W] fp32, labels [B,H,W] int64."""
        if x.device != self.dev:
            raise RuntimeError(f"input on {x.device}, engine on {self.dev}")
        B, _, H, W = x.shape
        if H % 32 or W % 32:
            raise ValueError("H and W must be multiples of 32")
        ncls = self.cfg["num_classes"]
        self._plan_key = (B, H, W)
        x_in = self.buf("x_in", (B, x.shape[1], H, W), torch.float32)
        lab = self.buf("labels_in", (B, H, W), torch.int64)
        x_in.copy_(x)
        lab.copy_(labels)
        loss_out = self.buf("loss_out", (1 + ncls,), torch.float32)
        self.step_count += 1
        hp = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, float(self.step_count), 0.0, 0.0], dtype=torch.float32)
        self.hyper.copy_(hp if self.dev.type != "cuda" else hp.pin_memory(), non_blocking=True)
        args = (x_in, lab, B, H, W, ncls, loss_out, w_dice, w_ce, lr, betas, eps, weight_decay, optimize, w_boundary)
        # hooks are part of the key: a step captured before replicas.GradSync was attached must not be replayed after
        key = (B, H, W, optimize, w_dice, w_ce, w_boundary, self.on_bucket is not None, self.grad_hook is not None)
        self._touch_weights()                                         # AdamW / BatchNorm statistics write raw pointers
        if not self.use_graph or self.taps is not None or self.dev.type != "cuda":
            self._step_body(*args)
            return loss_out
        segs = self._graphs.get(key)
        if segs is None:
            n0 = ops.launch_count()
            self._step_body(*args)                                    # eager warm-up step (allocates every buffer)
            self.launches_per_step = ops.launch_count() - n0
            torch.cuda.current_stream().synchronize()
            self._graphs[key] = self._capture(args)
            return loss_out                                           # the warm-up step WAS this iteration
        for kind, what in segs:
            if kind == "graph":
                what.replay()
            elif kind == "bucket":
                self._user_on_bucket(what)
            else:
                self._user_grad_hook(self.gflat)
        return loss_out
