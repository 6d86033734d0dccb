"""TEST INFRASTRUCTURE: albumentations.pytorch stand-in (import only; the reference has ToTensorV2 commented out)."""


class ToTensorV2:
    def __init__(self, *a, **k):
        pass
