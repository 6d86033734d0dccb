"""TEST INFRASTRUCTURE: imgaug stand-in (dataset_synapse.py:9-10,35-38,87-98): identity augmentation pipeline."""
from . import augmenters  # noqa: F401


class SegmentationMapOnImage:
    def __init__(self, arr, nb_classes=None, shape=None):
        self.arr = arr

    def get_arr_int(self):
        return self.arr


SegmentationMapsOnImage = SegmentationMapOnImage
