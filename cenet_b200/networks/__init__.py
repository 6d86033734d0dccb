"""Drop-in replacement for the reference's `src/networks` package (src/networks/__init__.py:1-2).

Two ways to import it:
  * `from cenet_b200.networks import CENet, CENetOrg`               (repo root on sys.path)
  * `from networks import CENet, CENetOrg`                          (what main_acdc.py:15 / main_synapse.py:15 /
    main_skin.py:13 say): put THIS package's parent directory (`<repo>/cenet_b200`) ahead of the reference's `src/` on
    `sys.path`.  A script's own directory is always sys.path[0], so the unchanged mains are started with
    `python -P` (or through `python -m cenet_b200.run_main main_acdc.py ...`), see INTEGRATION.md section 1.
When the package is imported under the bare name `networks`, it re-exports the one real package `cenet_b200.networks`
(same module object, same classes) instead of executing a second copy whose relative imports would leave the package.
"""
import os as _os
import sys as _sys

if __name__ != "cenet_b200.networks":
    _root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    if _root not in _sys.path:
        _sys.path.append(_root)
    import importlib as _importlib
    _real = _importlib.import_module("cenet_b200.networks")
    _sys.modules[__name__] = _real                     # `import networks` now IS cenet_b200.networks
    _sys.modules.setdefault(__name__ + ".cenet", _importlib.import_module("cenet_b200.networks.cenet"))
    _sys.modules.setdefault(__name__ + ".cenet_org", _importlib.import_module("cenet_b200.networks.cenet_org"))
    CENet, CENetOrg = _real.CENet, _real.CENetOrg
else:
    from .cenet import CENet
    from .cenet_org import CENetOrg

__all__ = ["CENet", "CENetOrg"]
