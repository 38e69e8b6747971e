"""
HUSL (human-friendly Hue, Saturation, Lightness) colour conversion on NVIDIA
B200: the ``nphusl`` API of TadLeonard/husl-numpy with a CUDA backend.

HUSL <-> RGB conversion API (unchanged from the reference, ``nphusl/__init__.py:19-27``)
   * ``to_husl``: RGB array -> HUSL array
   * ``to_rgb``:  HUSL array -> RGB array
   * ``to_hue``:  RGB array -> array of HUSL hues
   * ``chunk``, ``chunk_img``: tiling helpers (README.md:116,164 import them from here)

Backend switches (reference ``nphusl/__init__.py:45-88``, plus the new one)
   * ``enable_cuda`` / ``cuda_enabled``: the sm_100a kernels of ``libnphusl_cuda.so``
   * ``enable_simd`` / ``enable_cython`` / ``enable_numexpr`` / ``enable_numpy``
     and the matching ``*_enabled`` context managers keep their names; their
     implementations are not part of this repository, so unless one has been
     registered (``nphusl.nphusl.register_backend``) they fail with the
     reference's own ``AssertionError("implementation not available")``.

The CUDA backend is the default and there is no CPU fallback: if the shared
library or a GPU is missing, conversions raise instead of silently running
somewhere else.
"""

__version__ = "1.5.0+b200.1"
__all__ = ["to_husl", "to_hue", "to_rgb", "edit", "chunk", "chunk_img"]

from contextlib import contextmanager
from functools import partial

from .nphusl import to_husl, to_hue, to_rgb, edit
from .nphusl import CUDA, SIMD, CYTHON, NUMEXPR, NUMPY, BACKEND_NAMES, register_backend
from .transform import chunk, chunk_img
from . import nphusl
from . import constants
from . import transform
from . import _cuda
from . import _shard

if _cuda.available():
    register_backend(CUDA, _cuda)

_ORDER = (CUDA, SIMD, CYTHON, NUMEXPR, NUMPY)


def enable_best_optimized():
    """Per function, bind the first implementation available in the order
    CUDA, SIMD, CYTHON, NUMEXPR, NUMPY (reference ``__init__.py:45-50`` with
    CUDA in front); names nobody implements go back to the raising stub."""
    for name in BACKEND_NAMES:
        chosen = next((d[name] for d in _ORDER if d.get(name)), None)
        _select_fn(chosen or nphusl._unbound(name), name)


@contextmanager
def _with_fns(enable_other_fns, back_to_std=False):
    revert = enable_numpy if back_to_std and NUMPY else enable_best_optimized
    enable_other_fns()
    try:
        yield
    finally:
        revert()


def _enable_fns(fn_dictionary=None):
    # reference __init__.py:63-66: "NumPy everywhere, then overlay"; NumPy is
    # only there when a module was registered for it
    assert fn_dictionary, "implementation not available"
    if NUMPY:
        _set_module_globals(NUMPY)
    _set_module_globals(fn_dictionary)


def _set_module_globals(fn_dictionary):
    assert fn_dictionary, "implementation not available"
    for name, fn in fn_dictionary.items():
        _select_fn(fn, name)


def _select_fn(fn, name):
    globals()[name] = fn
    setattr(nphusl, name, fn)


enable_cuda = partial(_enable_fns, CUDA)
enable_numpy = partial(_enable_fns, NUMPY)
enable_cython = partial(_enable_fns, CYTHON)
enable_numexpr = partial(_enable_fns, NUMEXPR)
enable_simd = partial(_enable_fns, SIMD)
best_enabled = partial(_with_fns, enable_best_optimized)
cuda_enabled = partial(_with_fns, enable_cuda)
numpy_enabled = partial(_with_fns, enable_numpy)
cython_enabled = partial(_with_fns, enable_cython)
numexpr_enabled = partial(_with_fns, enable_numexpr)
simd_enabled = partial(_with_fns, enable_simd)

enable_best_optimized()

del partial
del contextmanager
