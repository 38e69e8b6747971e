"""Backend registry: which implementation of ``_rgb_to_husl`` / ``_husl_to_rgb``
/ ``_rgb_to_hue`` the public API is bound to.

The reference keeps one dictionary per backend and rebinds module globals
(``nphusl/__init__.py:45-88``, ``nphusl/nphusl.py:78-101``).  The dictionaries
and the rebinding survive here because callers rely on them
(``nphusl.SIMD``, ``nphusl.nphusl._rgb_to_husl`` ...); the switch functions are
generated from one table instead of being written out per backend, and CUDA
comes first in the preference order.
"""
from contextlib import contextmanager

from . import nphusl as _api

# every switchable backend, in the reference's preference order with CUDA in front
ORDER = ("cuda", "simd", "cython", "numexpr", "numpy")
# enable_best_optimized(): first backend that has the function wins.  NumPy is NOT in this list: the
# reference falls back to it silently (__init__.py:45-50); here a missing GPU must be loud, so the
# slow CPU path only ever runs when the caller asked for it by name.
BEST_ORDER = ("cuda", "simd", "cython", "numexpr")
TABLES = {"cuda": _api.CUDA, "simd": _api.SIMD, "cython": _api.CYTHON,
          "numexpr": _api.NUMEXPR, "numpy": _api.NUMPY}


def bind(name, fn, *modules):
    """Point ``<module>.<name>`` at ``fn`` in the API module and every mirror."""
    for mod in (_api,) + modules:
        setattr(mod, name, fn)


class Switchboard:
    def __init__(self, mirror):
        self.mirror = mirror            # the package module: reference code also reads nphusl._rgb_to_husl

    def best(self):
        for name in _api.BACKEND_NAMES:
            fn = next((TABLES[b][name] for b in BEST_ORDER if TABLES[b].get(name)), None)
            bind(name, fn or _api._unbound(name), self.mirror)

    def enable(self, backend):
        table = TABLES[backend]
        # the reference's message for a backend whose extension is not built (__init__.py:70)
        assert table, "implementation not available"
        if TABLES["numpy"]:             # reference order: NumPy everywhere, then overlay (__init__.py:63-66)
            for name, fn in TABLES["numpy"].items():
                bind(name, fn, self.mirror)
        for name, fn in table.items():
            bind(name, fn, self.mirror)

    def enabler(self, backend):
        def enable():
            self.enable(backend)
        enable.__name__ = "enable_" + backend
        enable.__doc__ = "Bind the public API to the {} backend.".format(backend)
        return enable

    def scope(self, enter):
        @contextmanager
        def enabled(back_to_std=False):
            enter()
            try:
                yield
            finally:
                # reference: revert to NumPy when asked (and present), else to the best available
                if back_to_std and TABLES["numpy"]:
                    self.enable("numpy")
                else:
                    self.best()
        return enabled
