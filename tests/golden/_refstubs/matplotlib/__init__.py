"""Import-only stand-in for matplotlib so the (read-only) reference modules import here.

Used ONLY by tests/golden/make_golden.py in the build container; never on the GPU box.
"""
import sys
import types


class _Sink:
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Sink()

    def __call__(self, *a, **k):
        return _Sink()

    def __getitem__(self, k):
        return _Sink()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())


def _sub_getattr(k):
    if k.startswith("__"):
        raise AttributeError(k)
    return _Sink()


for _n in ("pyplot", "animation", "font_manager"):
    _m = types.ModuleType("matplotlib." + _n)
    _m.__getattr__ = _sub_getattr
    _m.__file__ = __file__
    sys.modules["matplotlib." + _n] = _m
    globals()[_n] = _m
