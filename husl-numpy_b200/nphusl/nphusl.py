"""Public conversion API and the backend registry.

Host-side mirror of the reference's ``nphusl/nphusl.py:30-101`` for the hot
path: the three public functions keep their names, positional signature
``(arr, chunksize=None, out=None)``, decorator stack and return types, and the
backend is looked up in this module's globals at call time, under the private
names ``_rgb_to_husl`` / ``_husl_to_rgb`` / ``_rgb_to_hue`` -- so
``enable_*`` / ``*_enabled`` switch implementations exactly like the reference
(``nphusl/__init__.py:45-88``).

What differs: the registry has a fifth dictionary,
``CUDA``, filled from ``nphusl/_cuda.py`` and preferred over everything else.
``NUMPY`` is filled from ``nphusl/_numpy.py`` (a from-scratch float64 NumPy
backend, the one ``enable_numpy()`` / ``back_to_std=True`` select, and whose
step functions -- ``_to_light``, ``_max_lh_chroma``, ``_rgb_to_lch`` ... -- are
importable from this module like in the reference); the NumExpr / Cython /
SIMD dictionaries are only filled when such a module is registered from outside
(``register_backend``).  The NumPy backend is never chosen implicitly for the
public conversions: ``enable_best_optimized()`` binds CUDA or nothing, and with
nothing bound the public functions raise ``RuntimeError`` -- they never compute
on the CPU behind the caller's back.
"""
from . import transform
from . import _numpy
# step functions of the NumPy chain, under the reference's names (nphusl.py:104-392)
from ._numpy import (_lch_to_husl, _max_lh_chroma, _bounds, _ray_length, _luv_to_lch, _xyz_to_luv,   # noqa: F401
                     _to_light, _to_linear, _dot_product, _rgb_to_lch, _rgb_to_xyz, _lch_to_rgb,
                     _xyz_to_rgb, _lch_to_luv, _from_linear, _husl_to_lch, _luv_to_xyz, _from_light,
                     _channel, L_MAX, L_MIN)

__all__ = ["to_husl", "to_rgb", "to_hue", "edit", "CUDA", "SIMD", "CYTHON", "NUMEXPR", "NUMPY",
           "register_backend", "BACKEND_NAMES"]

# the three swappable callables of the hot path (reference nphusl.py:110,122,305)
BACKEND_NAMES = ("_rgb_to_husl", "_husl_to_rgb", "_rgb_to_hue")

CUDA = {}      # nphusl/_cuda.py -> libnphusl_cuda.so (sm_100a kernels)
SIMD = {}      # reference: nphusl/_simd_opt.pyx (not shipped here)
CYTHON = {}    # reference: nphusl/_cython_opt.pyx (not shipped here)
NUMEXPR = {}   # reference: nphusl/_numexpr_opt.py (not shipped here)
NUMPY = {name: getattr(_numpy, name) for name in _numpy.REGISTERED}   # nphusl/_numpy.py (reference: nphusl.py:104-392)


def _unbound(name):
    def missing(*args, **kwargs):
        raise RuntimeError(
            "nphusl: no backend is bound for {}: libnphusl_cuda.so is not built or no GPU is visible "
            "(see nphusl._cuda.load()).  The CUDA backend has no CPU fallback; the NumPy backend is only "
            "used when selected explicitly with nphusl.enable_numpy() / nphusl.numpy_enabled()".format(name))
    missing.__name__ = name
    missing.unbound = True
    return missing


_rgb_to_husl = _unbound("_rgb_to_husl")
_husl_to_rgb = _unbound("_husl_to_rgb")
_rgb_to_hue = _unbound("_rgb_to_hue")


def register_backend(registry: dict, module) -> dict:
    """Harvest the backend callables a module defines into ``registry`` (what
    ``optimized()`` does name by name in the reference, nphusl.py:84-101)."""
    for name in BACKEND_NAMES:
        fn = getattr(module, name, None)
        if fn is not None:
            registry[name] = fn
    return registry


### The API
### From RGB: to_husl, to_hue
### From HUSL: to_rgb

@transform.squeeze_output
@transform.reshape_image_input
@transform.reshape_rgba_input
def to_hue(rgb_img, chunksize: int = None, out=None, **backend_args):
    """RGB image (uint8, float in [0, 1], or grayscale float) -> array of HUSL
    hues in degrees, shape ``rgb_img.shape[:-1]``.  Reference: nphusl.py:30-36."""
    return transform.in_chunks(rgb_img, _rgb_to_hue, chunksize, out, **backend_args)


@transform.squeeze_output
@transform.reshape_husl_input
@transform.rgb_int_output
def to_rgb(husl_img, chunksize: int = None, out=None, **backend_args):
    """HUSL float array ``(..., 3)`` -> uint8 RGB array of the same shape.
    Reference: nphusl.py:39-45."""
    return transform.in_chunks(husl_img, _husl_to_rgb, chunksize, out, **backend_args)


@transform.squeeze_output
@transform.reshape_image_input
@transform.reshape_rgba_input
def to_husl(rgb_img, chunksize: int = None, out=None, **backend_args):
    """RGB image (uint8, float in [0, 1], or grayscale float) -> float64 HUSL
    array ``(..., 3)`` holding (H, S, L).  Reference: nphusl.py:48-54."""
    return transform.in_chunks(rgb_img, _rgb_to_husl, chunksize, out, **backend_args)


### Beyond the reference API: the callers' pipeline
### to_husl -> select -> edit -> to_rgb as one GPU pass.  CUDA backend only.

def edit(rgb_img, out=None, **edit_args):
    """``to_rgb(edit(to_husl(rgb_img)))`` fused into one kernel over uint8 RGB.

    Keywords (see ``nphusl._cuda.Edit``): ``hue=(lo, hi)``, ``saturation=(lo,
    hi)``, ``lightness=(lo, hi)`` select pixels (``invert=True`` for the
    complement); ``h=(mul, add)``, ``s=(mul, add)``, ``l=(mul, add)`` are applied
    to the selected ones.  Example (reference README.md:77-81)::

        out = nphusl.edit(img, hue=(250, 290), invert=True, l=(0.5, 0))
    """
    from . import _cuda
    return _cuda.edit_rgb(rgb_img, out=out, **edit_args)
