"""Test shim: the reference imports `easydict.EasyDict` (network.py:5, camera.py:4, tile.py:9, tools/*.py)
and this image does not have the package.  A functional stand-in written for these tests: a dict whose keys are also
attributes, nested dicts (also inside lists / tuples) converted on the way in."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        if d is None:
            d = {}
        if kwargs:
            d = dict(d, **kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setattr__(self, name, value):
        value = self._wrap(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            setattr(self, k, v)

    def pop(self, k, *args):
        if hasattr(self, k) and k in self.__dict__:
            delattr(self, k)
        return super().pop(k, *args)

    def __deepcopy__(self, memo):
        import copy
        return EasyDict({copy.deepcopy(k, memo): copy.deepcopy(v, memo) for k, v in self.items()})
