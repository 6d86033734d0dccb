"""TEST INFRASTRUCTURE: matplotlib stand-in (utils/utils.py:13, utils_synapse.py:24, utils_skin.py): every pyplot call is a
no-op; pandas' `.plot()` (utils/utils.py:22) is routed to a no-op plotting backend."""
from . import pyplot  # noqa: F401

try:
    import pandas as _pd
    _pd.set_option("plotting.backend", "cenet_stub_plot_backend")
except Exception:
    pass


def use(*a, **k):
    pass
