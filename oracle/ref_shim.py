"""TEST INFRASTRUCTURE ONLY -- stand-ins for the two third-party packages the reference imports.

The reference (`/root/reference/src/networks`) imports a handful of helpers from `timm==1.0.16` and
`monai==1.4.0` (requirements.txt:21, :8).  Neither package is installed in this image and there is no
network, so `install()` registers minimal modules under those names in `sys.modules` so that the reference
can be imported IN THIS CONTAINER to generate golden vectors (tests/golden/make_golden.py).

Only two pieces of arithmetic live here, both restated from the packages' published behaviour:
  * DropPath (timm.layers.drop.DropPath): train-only per-sample mask `bernoulli(1-p)/(1-p)`, shape [B,1,..,1].
  * Convolution(conv_only=True) (monai.networks.blocks.convolutions): an nn.Sequential whose single child
    `conv` is an nn.Conv2d / nn.ConvTranspose2d  (reference call sites: modules/unet.py:67-81).
Everything else is init helpers or factories.  Parity of these two stand-ins is NOT pinned by any reference
test ("parity unpinned" for DropPath; golden fixtures are generated with drop_prob = 0 / eval mode).

Nothing in the product package (cenet_b200/) imports this file.
"""
import sys
import types
import math
import collections.abc
from itertools import repeat

import torch
import torch.nn as nn


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def to_2tuple(v):
    if isinstance(v, collections.abc.Iterable) and not isinstance(v, str):
        return tuple(v)
    return tuple(repeat(v, 2))


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def trunc_normal_tf_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    with torch.no_grad():
        nn.init.trunc_normal_(tensor, 0.0, 1.0, a, b)
        tensor.mul_(std).add_(mean)
    return tensor


def named_apply(fn, module, name="", depth_first=True, include_root=False):
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = ".".join((name, child_name)) if name else child_name
        named_apply(fn=fn, module=child, name=child_name, depth_first=depth_first, include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module


def register_model(fn):
    return fn


class Convolution(nn.Sequential):
    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, act=None, norm=None,
                 dropout=None, bias=True, conv_only=False, is_transposed=False, padding=None,
                 output_padding=None, **kw):
        super().__init__()
        assert spatial_dims == 2 and conv_only, "shim covers only the conv_only 2-D use of the reference"
        if is_transposed:
            conv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=strides, padding=padding,
                                      output_padding=output_padding, bias=bias)
        else:
            conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=strides, padding=padding, bias=bias)
        self.add_module("conv", conv)


class _Act:
    PRELU = "prelu"
    LEAKYRELU = "leakyrelu"
    RELU = "relu"


class _Norm:
    INSTANCE = "instance"
    BATCH = "batch"


def get_act_layer(name):
    if isinstance(name, (tuple, list)):
        kind, args = name[0], dict(name[1])
    else:
        kind, args = name, {}
    kind = kind.lower()
    if kind == "leakyrelu":
        return nn.LeakyReLU(**args)
    if kind == "relu":
        return nn.ReLU(**args)
    if kind == "prelu":
        return nn.PReLU(**args)
    raise NotImplementedError(kind)


def get_norm_layer(name, spatial_dims=2, channels=1):
    kind = (name[0] if isinstance(name, (tuple, list)) else name).lower()
    if kind == "batch":
        return nn.BatchNorm2d(channels)
    if kind == "instance":
        return nn.InstanceNorm2d(channels)
    raise NotImplementedError(kind)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-ins (idempotent)."""
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_cenet_shim", False):
        return
    drop = _mod("timm.layers.drop", DropPath=DropPath)
    winit = _mod("timm.layers.weight_init", trunc_normal_=trunc_normal_, trunc_normal_tf_=trunc_normal_tf_)
    layers = _mod("timm.layers", DropPath=DropPath, to_2tuple=to_2tuple, trunc_normal_=trunc_normal_,
                  trunc_normal_tf_=trunc_normal_tf_, drop=drop, weight_init=winit)
    models = _mod("timm.models", register_model=register_model, named_apply=named_apply)
    _mod("timm", layers=layers, models=models, _cenet_shim=True)

    conv = _mod("monai.networks.blocks.convolutions", Convolution=Convolution)
    blocks = _mod("monai.networks.blocks", convolutions=conv)
    fact = _mod("monai.networks.layers.factories", Act=_Act, Norm=_Norm)
    utils = _mod("monai.networks.layers.utils", get_act_layer=get_act_layer, get_norm_layer=get_norm_layer)
    lay = _mod("monai.networks.layers", factories=fact, utils=utils)
    nets = _mod("monai.networks", blocks=blocks, layers=lay)
    _mod("monai", networks=nets)


def import_reference(src="/root/reference/src"):
    """Import the reference's `networks` package (this container only; /root/reference is absent on GPU boxes)."""
    install()
    if src not in sys.path:
        sys.path.insert(0, src)
    import importlib
    for k in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
        del sys.modules[k]
    return importlib.import_module("networks")
