"""TEST INFRASTRUCTURE: identity augmenters with imgaug's constructor names."""


class _Aug:
    def __init__(self, *a, **k):
        pass

    def to_deterministic(self):
        return self

    def augment_image(self, img):
        return img

    def augment_segmentation_maps(self, segmap):
        return segmap


class SomeOf(_Aug): pass
class Flipud(_Aug): pass
class Fliplr(_Aug): pass
class AdditiveGaussianNoise(_Aug): pass
class GaussianBlur(_Aug): pass
class LinearContrast(_Aug): pass
class Affine(_Aug): pass
class PiecewiseAffine(_Aug): pass
class Sequential(_Aug): pass
