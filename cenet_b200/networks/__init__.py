"""Drop-in replacement for the reference's `src/networks` package (src/networks/__init__.py:1-2).

Put the directory that contains this package (`cenet_b200/`) ahead of the reference's `src/` on `sys.path` and
`from networks import CENet, CENetOrg` in main_acdc.py / main_synapse.py / main_skin.py resolves here.
"""
from .cenet import CENet
from .cenet_org import CENetOrg

__all__ = ["CENet", "CENetOrg"]
