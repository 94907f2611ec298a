"""Import-only stand-in for TensorFlow (not installable here): lets reference modules whose *rollout* code is plain
NumPy (baselines/a2c/runner.py imports a2c/utils.py, which imports tensorflow at the top) be imported and driven
with a NumPy model.  Any attribute resolves to an inert object; nothing numeric is provided.  TEST INFRASTRUCTURE ONLY."""
import sys as _sys


class _Inert(object):
    def __getattr__(self, k):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


class _Module(type(_sys)):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Inert()


_sys.modules[__name__].__class__ = _Module
