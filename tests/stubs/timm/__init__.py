"""TEST INFRASTRUCTURE: registers the timm / monai stand-ins of oracle/ref_shim.py under their real names."""
import importlib.util as _u
import os as _os

_p = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "..", "..", "oracle", "ref_shim.py")
_spec = _u.spec_from_file_location("_cenet_ref_shim", _p)
_m = _u.module_from_spec(_spec)
_spec.loader.exec_module(_m)
_m.install()          # replaces sys.modules['timm'/'monai'...] with the stand-in modules
