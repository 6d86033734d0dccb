"""TEST INFRASTRUCTURE: tensorboardX stand-in (main_acdc.py:13,99): scalars are appended to <logdir>/scalars.txt."""
import os


class SummaryWriter:
    def __init__(self, logdir=None, *a, **k):
        self.logdir = logdir
        if logdir:
            os.makedirs(logdir, exist_ok=True)

    def add_scalar(self, tag, value, step=None, *a, **k):
        if self.logdir:
            with open(os.path.join(self.logdir, "scalars.txt"), "a") as fh:
                fh.write(f"{tag}\t{step}\t{float(value)}\n")

    def add_image(self, *a, **k):
        pass

    def add_images(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass
