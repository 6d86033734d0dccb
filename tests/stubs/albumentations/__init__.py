"""TEST INFRASTRUCTURE: albumentations stand-in (datasets/skin/dataset_ph2.py:244-258): identity pipeline."""
from . import pytorch  # noqa: F401


class _T:
    def __init__(self, *a, **k):
        pass


class Compose(_T):
    def __init__(self, transforms=None, *a, **k):
        self.transforms = transforms

    def __call__(self, **data):
        return dict(data)


class Rotate(_T): pass
class HorizontalFlip(_T): pass
class VerticalFlip(_T): pass
class RandomBrightnessContrast(_T): pass
class GaussianBlur(_T): pass
class ElasticTransform(_T): pass
class Resize(_T): pass
class Normalize(_T): pass
class ShiftScaleRotate(_T): pass
class ColorJitter(_T): pass
class GaussNoise(_T): pass
class RandomRotate90(_T): pass
class Flip(_T): pass
class Transpose(_T): pass


def __getattr__(name):
    if name and name[0].isupper():
        return _T
    raise AttributeError(name)
