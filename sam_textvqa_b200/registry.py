"""Stand-alone equivalent of the reference's global `registry` (/root/reference/tools/registry.py:1-3).

When the package runs inside the reference tree, `sa_m4c.py` imports the reference's own registry
instead, so `train.py`'s `registry.update(...)` is seen by the model."""


class _Registry(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


registry = _Registry()
