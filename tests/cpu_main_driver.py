"""TEST INFRASTRUCTURE: CPU dry run of an unchanged reference main (`python -P cpu_main_driver.py <main_x.py> <args>`).

The mains call `.cuda()` unconditionally and cenet_b200 has no CPU path, so in this GPU-less container the host side of the
drop-in is exercised by (1) turning `.cuda()` into the identity, (2) replacing the C-ABI ops of the launch plans with the torch
emulations tests/fake_ops.py / tests/fake_train_ops.py, (3) lifting CENet.forward's "CUDA tensors only" check.  Nothing here
ships: the product raises on a CPU tensor.  The GPU run of the same scripts is tests/test_gpu_mains.py.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.append(p)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
nn.Module.cuda = lambda self, *a, **k: self
torch.cuda.manual_seed = lambda *a, **k: None

import fake_ops  # noqa: E402
import fake_train_ops  # noqa: E402
import cenet_b200.engine as E  # noqa: E402
import cenet_b200.train as T  # noqa: E402
from cenet_b200 import _lib  # noqa: E402
from cenet_b200.networks import cenet as C  # noqa: E402

E.ops = fake_ops
T.ops = fake_ops
T.tops = fake_train_ops
_lib.load = lambda: None
os.environ["CENET_B200_PRECISION"] = "fp32"
os.environ["CENET_B200_GRAPH"] = "0"

_orig_forward = C.CENet.forward


class _FakeCuda:
    """view of a CPU tensor that answers is_cuda=True for the one check in CENet.forward"""


def _forward(self, x):
    if self.training:
        if torch.is_grad_enabled():
            return C._TrainForward.apply(self, x, *list(self.parameters()))
        return self.train_engine(x.device).forward_logits(x).clone()
    return self._engine(x).forward(x)


C.CENet.forward = _forward

# the emulations of the backward kernels use torch autograd internally; autograd.Function.backward runs with grad mode off
_orig_bwd = C._TrainForward.backward


def _bwd(ctx, dlogits):
    with torch.enable_grad():
        return _orig_bwd(ctx, dlogits)


C._TrainForward.backward = staticmethod(_bwd)

script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
runpy.run_path(script, run_name="__main__")
