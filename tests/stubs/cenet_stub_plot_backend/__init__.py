"""TEST INFRASTRUCTURE: pandas plotting backend that draws nothing (see tests/stubs/matplotlib)."""


def plot(data, kind=None, **kw):
    return None
