def __getattr__(name):
    def _absent(*a, **k):
        raise RuntimeError("matplotlib.pyplot is a test shim: not available in this image")
    return _absent
