"""TEST INFRASTRUCTURE: restatement of medpy 0.5.2 `metric.binary` (published algorithm; parity unpinned).
dc = 2|A&B| / (|A|+|B|) on boolean arrays is the integer-count contract cenet_b200.volume reproduces on the device."""
import numpy as np
from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure


def dc(result, reference):
    result = np.atleast_1d(np.asarray(result).astype(bool))
    reference = np.atleast_1d(np.asarray(reference).astype(bool))
    inter = np.count_nonzero(result & reference)
    a, b = np.count_nonzero(result), np.count_nonzero(reference)
    try:
        return 2.0 * inter / float(a + b)
    except ZeroDivisionError:
        return 0.0


def jc(result, reference):
    result = np.atleast_1d(np.asarray(result).astype(bool))
    reference = np.atleast_1d(np.asarray(reference).astype(bool))
    inter = np.count_nonzero(result & reference)
    union = np.count_nonzero(result | reference)
    return float(inter) / float(union)


def _surface_distances(result, reference, voxelspacing=None, connectivity=1):
    result = np.atleast_1d(np.asarray(result).astype(bool))
    reference = np.atleast_1d(np.asarray(reference).astype(bool))
    if not result.any() or not reference.any():
        raise RuntimeError("empty mask")
    fp = generate_binary_structure(result.ndim, connectivity)
    rb = result ^ binary_erosion(result, structure=fp, iterations=1)
    fb = reference ^ binary_erosion(reference, structure=fp, iterations=1)
    dt = distance_transform_edt(~fb, sampling=voxelspacing)
    return dt[rb]


def hd95(result, reference, voxelspacing=None, connectivity=1):
    a = _surface_distances(result, reference, voxelspacing, connectivity)
    b = _surface_distances(reference, result, voxelspacing, connectivity)
    return float(np.percentile(np.hstack((a, b)), 95))


def assd(result, reference, voxelspacing=None, connectivity=1):
    return float(np.mean((_surface_distances(result, reference, voxelspacing, connectivity).mean(),
                          _surface_distances(reference, result, voxelspacing, connectivity).mean())))
