"""Batched per-volume inference + on-device Dice counts.

The reference evaluates a test volume slice by slice (utils/metrics_eval.py:37-84, utils_synapse.py:50-98): CPU
`zoom(order=3)` to the patch size, one B=1 forward, `argmax(softmax)`, D2H, CPU `zoom(order=0)` back, then medpy metrics per
class on the host.  Here all slices of a volume go through the network as ONE batch (with the per-slice semantics of the
reference's B=1 loop: CCU skips its BatchNorm1d, cfam.py:260-261), the fused head writes int64 label maps, and one kernel
(`cenet_volume_labels_counts`) resizes them back (nearest, scipy's index convention) and accumulates, per class, the three
integers medpy's `dc` is made of.  Dice = 2*I/(P+L) from those counts is bit-identical to `medpy.metric.binary.dc` on the
same prediction; HD95 / Jaccard / ASSD need surface distances and stay on the host (medpy), as in the reference.

    pred, counts = predict_volume(net, image[D,H,W], patch_size=(224, 224), label=label, normalize=None | (0.5, 0.5))
    metrics = test_single_volume(image, label, net, classes, patch_size)      # drop-in for the reference function
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def nearest_index_table(n_in: int, n_out: int) -> np.ndarray:
    """source index of every output sample of `scipy.ndimage.zoom(x, n_out / n_in, order=0)` along one axis.

    scipy (grid_mode=False) maps output index o to the input coordinate o * (n_in - 1) / (n_out - 1) and, for order 0,
    takes floor(coordinate + 0.5) (ni_interpolation.c: spline order 0 -> nearest with round-half-up); n_out == 1 samples
    coordinate 0.  A coordinate that lands outside [0, n_in - 1] -- in double arithmetic, which happens to the LAST sample
    of some size pairs (223 * (511 / 223) = 511.00000000000006) -- is filled with the constant 0 (`mode='constant'`,
    the default): the table holds -1 there and the kernel writes label 0.  Checked against scipy in tests/test_volume_host.py."""
    if n_out == 1 or n_in == 1:
        return np.zeros(n_out, dtype=np.int32)
    scale = (n_in - 1) / (n_out - 1)
    coord = np.arange(n_out, dtype=np.float64) * scale
    idx = np.floor(coord + 0.5).astype(np.int64)
    idx = np.clip(idx, 0, n_in - 1)
    idx[(coord < 0) | (coord > n_in - 1)] = -1
    return idx.astype(np.int32)


def _zoom_in(slices: np.ndarray, patch_size):
    """the reference's input resize, unchanged: scipy cubic-spline zoom on the host (metrics_eval.py:45-46)"""
    from scipy.ndimage import zoom
    D, H, W = slices.shape
    if (H, W) == tuple(patch_size):
        return slices
    return np.stack([zoom(slices[d], (patch_size[0] / H, patch_size[1] / W), order=3) for d in range(D)])


@torch.no_grad()
def predict_volume(net, image, patch_size=(224, 224), label=None, normalize=None, max_batch=64, device=None):
    """image: [D,H,W] float array / tensor (one grey-scale volume).  Returns (prediction int64 [D,H,W] on the device,
    counts int64 [3, num_classes] on the device or None): counts[0,c] = |pred==c & label==c|, [1,c] = |pred==c|,
    [2,c] = |label==c|.  normalize=(mean, std) applies torchvision's Normalize after the resize (utils_synapse.py:60-64)."""
    img = image.detach().cpu().numpy() if torch.is_tensor(image) else np.asarray(image)
    if img.ndim != 3:
        raise ValueError(f"expected a [D,H,W] volume, got shape {img.shape}")
    dev = torch.device(device) if device is not None else next(net.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("cenet_b200 has no CPU path: move the module to a B200 first")
    D, H, W = img.shape
    ph, pw = int(patch_size[0]), int(patch_size[1])
    x = torch.from_numpy(np.ascontiguousarray(_zoom_in(img.astype(np.float32, copy=False), (ph, pw)), dtype=np.float32))
    x = x.pin_memory().to(dev, non_blocking=True).unsqueeze(1)
    if normalize is not None:
        x = (x - float(normalize[0])) / float(normalize[1])
    ncls = net.cfg["num_classes"]
    was_training = net.training
    net.eval()
    eng = net._engine(x)
    saved = eng.ccu_bn1d
    eng.ccu_bn1d = False                                          # the reference runs every slice at B = 1
    try:
        pred_patch = torch.empty((D, ph, pw), device=dev, dtype=torch.int64)
        for lo in range(0, D, max_batch):
            hi = min(D, lo + max_batch)
            eng.forward(x[lo:hi].contiguous(), labels=True, out=pred_patch[lo:hi])
    finally:
        eng.ccu_bn1d = saved
        net.train(was_training)
    iy = torch.from_numpy(nearest_index_table(ph, H)).to(dev)
    ix = torch.from_numpy(nearest_index_table(pw, W)).to(dev)
    pred = torch.empty((D, H, W), device=dev, dtype=torch.int64)
    counts = lab = None
    if label is not None:
        lab = label if torch.is_tensor(label) else torch.from_numpy(np.ascontiguousarray(label))
        if lab.dtype not in (torch.float32, torch.int64, torch.uint8):
            lab = lab.to(torch.float32 if lab.is_floating_point() else torch.int64)
        lab = lab.to(dev).contiguous()
        counts = torch.empty(3 * ncls, device=dev, dtype=torch.int64)
    ops.volume_labels_counts(pred_patch, iy, ix, lab, pred, counts, ncls)
    return pred, (counts.view(3, ncls) if counts is not None else None)


def dice_from_counts(counts, cls: int):
    """medpy.metric.binary.dc(pred == cls, label == cls) from the integer counts (identical arithmetic: 2*I / float(P+L))"""
    i, p, l = (int(counts[k, cls]) for k in range(3))
    return 2.0 * i / float(p + l) if (p + l) > 0 else 0.0


def test_single_volume(image, label, net, classes, patch_size=(256, 256), test_save_path=None, case=None, z_spacing=1,
                       epoch=0, normalize=None, surface_metrics=True):
    """Drop-in for `utils.metrics_eval.test_single_volume` (same arguments, same return value: one (dice, hd95, jaccard,
    asd) tuple per foreground class).  Dice comes from the device counts; the surface metrics are computed by medpy on the
    host exactly as in the reference when `surface_metrics` (they need the prediction on the CPU anyway)."""
    image = image.squeeze(0) if image.dim() == 4 else image
    label = label.squeeze(0) if label.dim() == 4 else label
    pred, counts = predict_volume(net, image, patch_size, label=label, normalize=normalize)
    counts = counts.cpu()
    out = []
    pred_np = lab_np = None
    for c in range(1, classes):
        i, p, l = (int(counts[k, c]) for k in range(3))
        if p > 0 and l > 0:
            dice = dice_from_counts(counts, c)
            hd = jc = asd = 0.0
            if surface_metrics:
                from medpy import metric
                if pred_np is None:
                    pred_np, lab_np = pred.cpu().numpy(), np.asarray(label.cpu() if torch.is_tensor(label) else label)
                a, b = pred_np == c, lab_np == c
                hd, jc, asd = metric.binary.hd95(a, b), metric.binary.jc(a, b), metric.binary.assd(a, b)
            out.append((dice, hd, jc, asd))
        elif p > 0 and l == 0:
            out.append((1, 0, 1, 0))                              # metrics_eval.py:27-28
        else:
            out.append((0, 0, 0, 0))
    return out


test_single_volume.__test__ = False          # not a pytest test
