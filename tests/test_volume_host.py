"""Host pieces of cenet_b200.volume: the nearest-neighbour index tables must reproduce scipy's `zoom(order=0)` (the resize
the reference applies to every predicted slice, metrics_eval.py:54-55) for the size pairs an evaluation meets."""
import numpy as np
import pytest
from scipy.ndimage import zoom

from cenet_b200.volume import dice_from_counts, nearest_index_table


@pytest.mark.parametrize("n_in,n_out", [(224, 512), (224, 256), (224, 80), (224, 72), (224, 224), (224, 313), (96, 41), (64, 1),
                                        (224, 447), (512, 224)])
def test_nearest_table_matches_scipy_zoom_order0(n_in, n_out):
    src = np.arange(n_in, dtype=np.float64)
    want = zoom(src, n_out / n_in, order=0)
    if want.shape[0] != n_out:                       # scipy rounds the output length; only equal lengths are comparable
        pytest.skip("scipy picks a different output length for this factor")
    t = nearest_index_table(n_in, n_out)
    got = np.where(t >= 0, src[np.maximum(t, 0)], 0.0)
    assert np.array_equal(got, want)


def test_2d_label_map_roundtrip():
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 9, (224, 224)).astype(np.int64)
    want = zoom(lab, (80 / 224, 72 / 224), order=0)
    ty, tx = nearest_index_table(224, 80), nearest_index_table(224, 72)
    got = lab[np.maximum(ty, 0)][:, np.maximum(tx, 0)]
    got[ty < 0, :] = 0
    got[:, tx < 0] = 0
    assert np.array_equal(got, want)


def test_dice_from_counts_is_medpy_arithmetic():
    import torch
    c = torch.tensor([[0, 10], [0, 30], [0, 25]])
    assert dice_from_counts(c, 1) == 2.0 * 10 / float(30 + 25)
    assert dice_from_counts(torch.zeros(3, 2, dtype=torch.int64), 1) == 0.0
