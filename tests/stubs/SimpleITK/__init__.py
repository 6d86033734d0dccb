"""TEST INFRASTRUCTURE: SimpleITK stand-in (metrics_eval.py:74-82): volumes are written as .npy next to the asked path."""
import numpy as np


class _Image:
    def __init__(self, arr):
        self.arr = np.asarray(arr)
        self.spacing = None

    def SetSpacing(self, s):
        self.spacing = tuple(s)


def GetImageFromArray(arr):
    return _Image(arr)


def WriteImage(img, path):
    np.save(path + ".npy", img.arr)
