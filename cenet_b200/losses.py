"""Drop-in for the reference's `utils.core.Criterion` (utils/core.py:161-188) on the fused loss kernels.

    crit = Criterion(num_classes, args)          # args.loss_type "dice,ce" / "boundary" / ..., args.loss_weights "0.5,0.5"
    loss = crit(outputs, labels); loss.backward()

One reduction pass + one gradient pass over the logits for the whole weighted sum (DiceLoss core.py:57-80, CrossEntropyLoss,
BoundaryDoULoss core.py:83-131) instead of the reference's per-class Python loops, per-sample conv2d calls and host syncs.
The gradient w.r.t. the logits is produced in the forward call and handed to autograd in backward.  CUDA only."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _FusedCriterion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, w_dice, w_ce, w_boundary):
        if not logits.is_cuda:
            raise RuntimeError("cenet_b200.losses.Criterion has no CPU path (the CPU oracle lives in oracle/ and is test-only)")
        B, ncls, H, W = logits.shape
        lg = logits.detach().float().contiguous()
        lab = labels.detach().reshape(B, H, W).long().contiguous()
        out = torch.empty(1 + ncls, device=lg.device, dtype=torch.float32)
        grad = torch.empty_like(lg)
        ws = ops.seg_loss_ws(B, ncls, H, W, lg.device)
        ops.seg_loss(lg, lab, out, grad, ws, B, ncls, H, W, w_dice, w_ce, w_boundary)
        ctx.save_for_backward(grad)
        ctx.in_dtype = logits.dtype
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).to(ctx.in_dtype), None, None, None, None


class Criterion(nn.Module):
    """Same constructor / call as utils.core.Criterion: `args.loss_type` and `args.loss_weights` are comma-separated lists."""

    def __init__(self, num_classes, args):
        super().__init__()
        self.num_classes = int(num_classes)
        self.lnames = args.loss_type.split(",")
        self.weights = [float(w) for w in args.loss_weights.split(",")]
        w = {"dice": 0.0, "ce": 0.0, "boundary": 0.0}
        for name, wt in zip(self.lnames, self.weights):
            if name not in w:
                raise NotImplementedError(f"Loss {name} not implemented")        # core.py:176 message
            w[name] += wt
        self.w = w

    def forward(self, outputs, labels):
        assert outputs.shape[1] == self.num_classes, \
            'predict {} & target {} shape do not match'.format(outputs.size(), labels.size())           # core.py:73 / 122
        return _FusedCriterion.apply(outputs, labels, self.w["dice"], self.w["ce"], self.w["boundary"])
