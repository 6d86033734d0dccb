"""TEST INFRASTRUCTURE: medpy.metric stand-in (main_acdc.py:11, metrics_eval.py:3, utils_synapse.py:3, utils_skin.py)."""
from . import binary
from .binary import assd, dc, hd95, jc

__all__ = ["binary", "dc", "hd95", "jc", "assd"]
