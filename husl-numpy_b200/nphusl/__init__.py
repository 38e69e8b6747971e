"""
HUSL (human-friendly Hue, Saturation, Lightness) colour conversion on NVIDIA
B200: the ``nphusl`` API of TadLeonard/husl-numpy with a CUDA backend.

HUSL <-> RGB conversion API (unchanged from the reference, ``nphusl/__init__.py:19-27``)
   * ``to_husl``: RGB array -> HUSL array
   * ``to_rgb``:  HUSL array -> RGB array
   * ``to_hue``:  RGB array -> array of HUSL hues
   * ``chunk``, ``chunk_img``: tiling helpers (README.md:116,164 import them from here)
   * ``edit``: fused to_husl -> select -> edit -> to_rgb on the GPU (new)

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
import sys as _sys

__version__ = "1.5.0+b200.1"
__all__ = ["to_husl", "to_hue", "to_rgb", "edit", "chunk", "chunk_img"]

from .nphusl import to_husl, to_hue, to_rgb, edit
from .nphusl import CUDA, SIMD, CYTHON, NUMEXPR, NUMPY, BACKEND_NAMES, register_backend
from .transform import chunk, chunk_img
from . import nphusl, constants, transform, _cuda, _shard, _registry

if _cuda.available():
    register_backend(CUDA, _cuda)

_switch = _registry.Switchboard(_sys.modules[__name__])
enable_best_optimized = _switch.best
best_enabled = _switch.scope(_switch.best)
for _name in _registry.ORDER:
    _enable = _switch.enabler(_name)
    globals()["enable_" + _name] = _enable                  # enable_cuda, enable_simd, ...
    globals()[_name + "_enabled"] = _switch.scope(_enable)   # cuda_enabled, simd_enabled, ...
del _name, _enable

enable_best_optimized()
