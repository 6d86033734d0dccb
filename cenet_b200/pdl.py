"""Programmatic dependent launch (PDL) for captured CUDA graphs.

Every kernel of libcenet_b200 starts with `griddepcontrol.launch_dependents; griddepcontrol.wait;` (csrc/common.cuh,
`pdl_prologue`).  Under the full dependencies that stream capture records these are no-ops.  `relax(graph)` rewrites every
edge A -> B of a captured graph where A and B are both kernels of this library and B has no other predecessor into a
PROGRAMMATIC edge (A's programmatic port, CU_GRAPH_DEPENDENCY_TYPE_PROGRAMMATIC): B may then be scheduled while A is still
running and blocks in its prologue until A has completed and flushed.  Ordering of memory operations is unchanged (every
kernel waits before its first access; completion is transitive along chains); what disappears is the launch latency and
ramp-up between the ~330 (inference) / ~1800 (training step) mostly small kernels.

Edges that touch foreign kernels (torch copies, NCCL), memcpy / memset nodes or joins stay full dependencies.
Needs torch >= 2.8 (`CUDAGraph(keep_graph=True)`, `raw_cuda_graph`, `instantiate`) and cuda-python; callers fall back to the
unmodified graph when either is missing.  The pass is opt-in (CENET_B200_PDL=1), see `enabled`."""
from __future__ import annotations

import os


def enabled() -> bool:
    """OFF by default: measured on B200 (round 1) with every kernel->kernel edge programmatic and the trigger at the top of
    each kernel, the inference forward went 13.52 -> 13.68 ms and the training step 32.20 -> 33.30 ms: inside an
    instantiated graph the kernel->kernel latency is already small and the early-resident dependents disturb the tail of
    the running kernel.  CENET_B200_PDL=1 enables the pass (results are bit-identical either way)."""
    return os.environ.get("CENET_B200_PDL", "0") == "1"


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


def _ours(name: bytes) -> bool:
    """kernels of libcenet_b200 (all end in `_kernel`, none lives in at::native)"""
    return b"_kernel" in name and b"2at" not in name and b"nccl" not in name.lower()


def relax(graph) -> dict:
    """Rewrite eligible kernel->kernel edges of `graph` (a torch.cuda.CUDAGraph captured with keep_graph=True and not yet
    instantiated) into programmatic edges.  Returns counters."""
    from cuda.bindings import driver as drv
    g = drv.CUgraph(int(graph.raw_cuda_graph()))
    _, n = _check(drv.cuGraphGetNodes(g, 0))
    nodes, n = _check(drv.cuGraphGetNodes(g, n))
    kind = {}
    for nd in nodes:
        t = _check(drv.cuGraphNodeGetType(nd))
        ours = False
        if t == drv.CUgraphNodeType.CU_GRAPH_NODE_TYPE_KERNEL:
            p = _check(drv.cuGraphKernelNodeGetParams(nd))
            name = None
            for getter, handle in ((drv.cuFuncGetName, getattr(p, "func", None)), (drv.cuKernelGetName, getattr(p, "kern", None))):
                try:
                    if handle is not None and int(handle) != 0:
                        name = _check(getter(handle))
                        if name:
                            break
                except Exception:
                    name = None
            ours = bool(name) and _ours(name if isinstance(name, bytes) else str(name).encode())
        kind[int(nd)] = ours
    _, _, _, ne = _check(drv.cuGraphGetEdges_v2(g, 0))
    src, dst, data, ne = _check(drv.cuGraphGetEdges_v2(g, ne))
    indeg = {}
    for d in dst:
        indeg[int(d)] = indeg.get(int(d), 0) + 1
    rem_f, rem_t, rem_d, add_d = [], [], [], []
    # (stream capture records default full edges only; the edge-data array returned by the binding is not reliable, so
    #  the removal is issued with freshly zeroed = default edge data)
    for a, b in zip(src, dst):
        if kind.get(int(a)) and kind.get(int(b)) and indeg[int(b)] == 1:
            rem_f.append(a); rem_t.append(b); rem_d.append(drv.CUgraphEdgeData())
            nd = drv.CUgraphEdgeData()
            nd.from_port = 1                      # CU_GRAPH_KERNEL_NODE_PORT_PROGRAMMATIC
            nd.to_port = 0
            nd.type = drv.CUgraphDependencyType.CU_GRAPH_DEPENDENCY_TYPE_PROGRAMMATIC
            add_d.append(nd)
    if rem_f:
        _check(drv.cuGraphRemoveDependencies_v2(g, rem_f, rem_t, rem_d, len(rem_f)))
        _check(drv.cuGraphAddDependencies_v2(g, rem_f, rem_t, add_d, len(rem_f)))
    return dict(nodes=len(nodes), kernels_ours=sum(1 for v in kind.values() if v), edges=len(src), programmatic=len(rem_f))


def capture(fn, pool=None):
    """Capture `fn()` on the current stream into a CUDA graph; with PDL enabled the graph is relaxed before instantiation.
    Returns (graph, stats or None)."""
    import torch
    stats = None
    if enabled():
        try:
            g = torch.cuda.CUDAGraph(keep_graph=True)
            g.capture_begin(**({"pool": pool} if pool is not None else {}))
            try:
                fn()
            finally:
                g.capture_end()
            try:
                stats = relax(g)
            except Exception as e:                       # driver / binding mismatch: keep the plain graph
                stats = dict(error=repr(e))
            g.instantiate()
            return g, stats
        except TypeError:
            pass                                         # torch without keep_graph
    g = torch.cuda.CUDAGraph()
    g.capture_begin(**({"pool": pool} if pool is not None else {}))
    try:
        fn()
    finally:
        g.capture_end()
    return g, stats
