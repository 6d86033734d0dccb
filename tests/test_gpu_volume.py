"""Batched volume inference (cenet_b200.volume) against the reference's own per-slice loop.

`utils.metrics_eval.test_single_volume` (vendored byte copy under baseline/_ref, third-party imports from tests/stubs) runs the
drop-in module slice by slice at B=1 exactly as main_acdc.py does; `cenet_b200.volume.test_single_volume` pushes the whole
volume through as one batch and counts on the device.  Checked: prediction maps identical voxel for voxel except for argmax
near-ties (batch-size dependent accumulation order), Dice from the device's INTEGER counts bit-identical to medpy's `dc` on
the same prediction, counts bit-exact against numpy, same return structure."""
import os
import sys

import numpy as np
import pytest
import torch

import mains_harness as H
from oracle import fixtures

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(name="acdc"):
    from cenet_b200.networks import CENet
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(1234)
    m = CENet(**kw)
    m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
    return m.to(DEV).eval(), kw


def _volume(D, Hh, Ww, ncls, seed=3):
    rng = np.random.default_rng(seed)
    img, lab = H._blobs(rng, D, Hh, Ww, ncls)
    return img, lab.astype(np.float32)


@pytest.mark.parametrize("shape", [(5, 80, 72), (3, 224, 224), (4, 256, 313)])
def test_counts_are_exact_and_dice_matches_medpy(shape):
    from cenet_b200 import volume
    sys.path.insert(0, H.STUBS)
    from medpy.metric.binary import dc
    net, kw = _net()
    img, lab = _volume(*shape, kw["num_classes"])
    pred, counts = volume.predict_volume(net, img, (224, 224), label=lab)
    pred_np, c = pred.cpu().numpy(), counts.cpu().numpy()
    assert pred_np.shape == lab.shape
    for k in range(kw["num_classes"]):
        a, b = pred_np == k, lab == k
        assert c[0, k] == np.count_nonzero(a & b) and c[1, k] == np.count_nonzero(a) and c[2, k] == np.count_nonzero(b)
        if c[1, k] + c[2, k] > 0:
            assert volume.dice_from_counts(counts.cpu(), k) == dc(a, b)          # bit-identical floats
    # uint8 / int64 label inputs take the same path
    _, c2 = volume.predict_volume(net, img, (224, 224), label=lab.astype(np.int64))
    _, c3 = volume.predict_volume(net, img, (224, 224), label=lab.astype(np.uint8))
    assert torch.equal(c2, counts) and torch.equal(c3, counts)


@pytest.mark.skipif(not H.have_reference(), reason="baseline/_ref not vendored")
def test_batched_volume_matches_reference_per_slice_loop():
    from cenet_b200 import volume
    for p in (H.STUBS, H.REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_metrics_eval", os.path.join(H.REF_SRC, "utils", "metrics_eval.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)                                  # the reference's own file, unmodified
    net, kw = _net()
    img, lab = _volume(6, 80, 72, kw["num_classes"])
    image_t, label_t = torch.from_numpy(img)[None], torch.from_numpy(lab)[None]        # DataLoader batch of one volume
    want = ref.test_single_volume(image_t, label_t, net, classes=kw["num_classes"], patch_size=[224, 224])
    got = volume.test_single_volume(image_t, label_t, net, classes=kw["num_classes"], patch_size=[224, 224])
    assert len(got) == len(want) == kw["num_classes"] - 1
    # the per-slice loop's prediction, rebuilt the way the reference builds it, vs the batched one
    pred_b, _ = volume.predict_volume(net, img, (224, 224))
    from scipy.ndimage import zoom
    pred_s = np.zeros_like(lab)
    with torch.no_grad():
        for d in range(img.shape[0]):
            sl = zoom(img[d], (224 / 80, 224 / 72), order=3)
            out = torch.argmax(torch.softmax(net(torch.from_numpy(sl)[None, None].float().to(DEV)), 1), 1).squeeze(0).cpu().numpy()
            pred_s[d] = zoom(out, (80 / 224, 72 / 224), order=0)
    agree = (pred_b.cpu().numpy() == pred_s).mean()
    assert agree >= 0.9999, agree
    for g, w in zip(got, want):
        assert abs(g[0] - w[0]) < 2e-3 and abs(g[2] - w[2]) < 2e-3, (g, w)       # dice / jaccard (near-tie voxels only)


def test_volume_batches_larger_than_max_batch_and_synapse_normalisation():
    from cenet_b200 import volume
    net, kw = _net("synapse")
    img, lab = _volume(7, 96, 80, kw["num_classes"], seed=9)
    p1, c1 = volume.predict_volume(net, img, (224, 224), label=lab, normalize=(0.5, 0.5), max_batch=64)
    p2, c2 = volume.predict_volume(net, img, (224, 224), label=lab, normalize=(0.5, 0.5), max_batch=3)
    assert (p1 == p2).float().mean().item() >= 0.9999
    assert int(c1[2].sum()) == lab.size
