"""TEST INFRASTRUCTURE ONLY (oracle shim) -- never imported by the product path.

Minimal stand-in for the un-vendored `easydict` package that the reference imports
(`/root/reference/tools/registry.py:1`, `train.py:50-51`).  Attribute access on a dict,
recursive on nested dicts / lists, which is all the reference relies on.
"""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setattr__(self, k, v):
        v = self._wrap(v)
        super().__setattr__(k, v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            setattr(self, k, v)

    def pop(self, k, *a):
        if hasattr(self, k):
            delattr(self, k)
        return super().pop(k, *a)
