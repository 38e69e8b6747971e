"""Shape / dtype boundary of the nphusl API and chunked application.

Host-side mirror of the reference's ``nphusl/transform.py`` (public names,
argument meaning and error behaviour are kept so callers and tests written for
the reference keep working):

* dtype coercion -- ``ensure_{int,float,rgb_int,rgb_float}`` and the
  ``*_input`` / ``*_output`` decorators (reference ``transform.py:16-103``);
* shape normalisation -- ``reshape_image_input`` (triplets, flat lists,
  grayscale floats, RGBA), ``reshape_husl_input``, ``reshape_rgba_input``,
  ``squeeze_output`` (reference ``transform.py:137-246``);
* chunked application -- ``in_chunks``, ``chunk_img``, ``chunk``,
  ``chunk_apply``, ``chunk_apply_1d``, ``chunk_many`` (reference
  ``transform.py:251-308``).

Differences from the reference, all deliberate:

* objects exposing ``__cuda_array_interface__`` (device-resident frames) pass
  through every decorator untouched -- nothing here may call
  ``np.ascontiguousarray`` on them, that would drag the frame to the host;
* ``np.floating`` / ``np.integer`` replace the removed ``np.float`` alias, so
  float32 inputs are recognised as floats (App. D-7);
* ``in_chunks`` implements the documented ``chunksize`` / ``out`` behaviour
  (reference ``README.md:55-57``); in 1.5.0 the two arguments are swapped at the
  call sites (App. D-1).  Both argument orders are accepted.
"""

import math
import warnings
from collections import namedtuple
from enum import Enum
from functools import partial, wraps

import numpy as np
from numpy import ndarray


def is_device_array(obj) -> bool:
    """True for anything that lives in GPU memory (torch CUDA tensors, our own
    ``DeviceArray``, ...) -- recognised by ``__cuda_array_interface__``."""
    if isinstance(obj, ndarray):
        return False
    is_cuda = getattr(obj, "is_cuda", None)      # torch answers this without building the interface dict
    if isinstance(is_cuda, bool):
        return is_cuda
    return hasattr(obj, "__cuda_array_interface__")


_TORCH_DT = {"torch.uint8": np.dtype(np.uint8), "torch.float32": np.dtype(np.float32), "torch.float64": np.dtype(np.float64)}


def _dtype_of(arr):
    if is_device_array(arr):
        dt = getattr(arr, "dtype", None)
        if isinstance(dt, np.dtype):                 # DeviceArray
            return dt
        dt = _TORCH_DT.get(str(dt))                  # torch tensor: no interface dict (3.4 us to build)
        if dt is not None:
            return dt
        return np.dtype(arr.__cuda_array_interface__["typestr"])
    return arr.dtype


# --------------------------------------------------------------------------
# dtype coercion
# --------------------------------------------------------------------------

type_tuple = namedtuple("type_tuple", "base exact convert exact_required")


def scale_up_rgb(rgb: ndarray):
    """[0, 1] floats -> [0, 255], rounded half-to-even like ``np.round``."""
    return np.round(rgb * 255.0)


def scale_down_rgb(rgb: ndarray):
    """[0, 255] integers -> [0, 1] floats."""
    return rgb / 255.0


def to_dtype(dtype, arr: ndarray):
    return arr.astype(dtype.exact)


def to_rgb_int(dtype, arr: ndarray):
    """float RGB in [0, 1] -> uint8; integer input is only cast.  No clamp:
    negative values are mirrored by ``abs`` and values past 255 wrap, exactly
    like the reference (``transform.py:16-20``, App. D-6)."""
    if not np.issubdtype(arr.dtype, dtype.base):
        arr = scale_up_rgb(arr)
    return to_dtype(dtype, np.abs(arr))


def to_rgb_float(dtype, arr: ndarray):
    """integer RGB in [0, 255] -> float64 in [0, 1]; float input is only cast."""
    if not np.issubdtype(arr.dtype, dtype.base):
        arr = scale_down_rgb(arr)
    return to_dtype(dtype, arr)


class Dtype(type_tuple, Enum):
    int = type_tuple(np.integer, np.int64, to_dtype, False)
    float = type_tuple(np.floating, np.float64, to_dtype, False)
    rgb_int = type_tuple(np.integer, np.uint8, to_rgb_int, True)
    rgb_float = type_tuple(np.floating, np.float64, to_rgb_float, True)


def ensure_dtype(dtype: Dtype, arr):
    """Return ``arr`` unchanged when it already has an acceptable dtype, else a
    converted copy.  Device arrays are never touched: the CUDA backend accepts
    uint8 / float32 / float64 and converts inside its kernels."""
    if is_device_array(arr):
        return arr
    wanted = dtype.exact if dtype.exact_required else dtype.base
    if np.issubdtype(arr.dtype, wanted):
        return arr
    return dtype.convert(dtype, arr)


ensure_int = partial(ensure_dtype, Dtype.int)
ensure_float = partial(ensure_dtype, Dtype.float)
ensure_rgb_int = partial(ensure_dtype, Dtype.rgb_int)
ensure_rgb_float = partial(ensure_dtype, Dtype.rgb_float)


def ensure_input_dtype(dtype: Dtype):
    """Decorator factory: first positional argument is coerced to ``dtype``."""
    def decorate(fn):
        @wraps(fn)
        def wrapped(arr, *args, **kwargs):
            return fn(ensure_dtype(dtype, arr), *args, **kwargs)
        return wrapped
    return decorate


def ensure_output_dtype(dtype: Dtype):
    """Decorator factory: the return value is coerced to ``dtype``."""
    def decorate(fn):
        @wraps(fn)
        def wrapped(*args, **kwargs):
            return ensure_dtype(dtype, fn(*args, **kwargs))
        return wrapped
    return decorate


int_input = ensure_input_dtype(Dtype.int)
float_input = ensure_input_dtype(Dtype.float)
rgb_int_input = ensure_input_dtype(Dtype.rgb_int)
rgb_float_input = ensure_input_dtype(Dtype.rgb_float)
int_output = ensure_output_dtype(Dtype.int)
float_output = ensure_output_dtype(Dtype.float)
rgb_int_output = ensure_output_dtype(Dtype.rgb_int)
rgb_float_output = ensure_output_dtype(Dtype.rgb_float)


# --------------------------------------------------------------------------
# reshape alerts
# --------------------------------------------------------------------------

Alert = Enum("Alert", "error warn quiet")


class ReshapeAlert(str, Enum):
    grayscale = "Conversion from grayscale input"
    flat = "Conversion from a 1-dimensional input"
    rgba = "Conversion from RGBA"


# user-tunable: error / warn / quiet per kind of guess
alert_levels = {
    ReshapeAlert.grayscale: Alert.warn,
    ReshapeAlert.flat: Alert.warn,
    ReshapeAlert.rgba: Alert.warn,
}


def _alert(kind: ReshapeAlert, msg: str):
    level = alert_levels[kind]
    text = "{}: {}".format(kind, msg)
    if level == Alert.error:
        raise ValueError(text)
    if level == Alert.warn:
        warnings.warn(text)


_alert_flat = partial(_alert, ReshapeAlert.flat)
_alert_gray = partial(_alert, ReshapeAlert.grayscale)
_alert_rgba = partial(_alert, ReshapeAlert.rgba)


# --------------------------------------------------------------------------
# shape normalisation
# --------------------------------------------------------------------------

def _as_array(arr):
    if is_device_array(arr) or isinstance(arr, (np.ndarray, np.generic)):
        return arr
    return np.ascontiguousarray(arr)


def _is_float(arr) -> bool:
    return np.issubdtype(_dtype_of(arr), np.floating)


class GrayDevice:
    """(...,) device array presented as (..., 3): same pointer, channel stride
    0 in its ``__cuda_array_interface__`` (the CUDA backend reads one value per
    pixel and uses it for R, G and B)."""

    def __init__(self, base):
        cai = dict(base.__cuda_array_interface__)
        shape = tuple(cai["shape"])
        item = np.dtype(cai["typestr"]).itemsize
        strides = cai.get("strides")
        if strides is None:
            strides, acc = [], item
            for n in reversed(shape):
                strides.append(acc)
                acc *= n
            strides = tuple(reversed(strides))
        cai["shape"] = shape + (3,)
        cai["strides"] = tuple(strides) + (0,)
        self.__cuda_array_interface__ = cai
        self.base = base
        self.shape = cai["shape"]
        self.dtype = np.dtype(cai["typestr"])
        self.device = getattr(base, "device", None)


class RgbaDevice:
    """(..., 4) RGBA device array presented as the (..., 3) RGB image it stands
    for: the CUDA backend premultiplies by alpha while loading (no copy)."""
    rgba = True

    def __init__(self, base):
        self.__cuda_array_interface__ = base.__cuda_array_interface__
        self.base = base
        self.shape = tuple(base.shape[:-1]) + (3,)
        self.dtype = np.dtype(base.__cuda_array_interface__["typestr"])
        self.device = getattr(base, "device", None)


def from_grayscale(img):
    """(...,) single-channel floats -> (..., 3) with the value in R, G and B.
    Nothing is materialised: the CUDA backend has grayscale kernels that
    broadcast in registers instead of in memory."""
    if is_device_array(img):
        return GrayDevice(img)
    # zero-copy (..., 3) view with a 0 stride on the channel axis: any backend
    # reads it as RGB, the CUDA backend recognises the stride and ships ONE
    # channel to the GPU instead of three
    return np.broadcast_to(img[..., None], img.shape + (3,))


def reshape_image_input(fn):
    """RGB-side entry guard: lists become arrays, ``[r, g, b]`` becomes
    ``[[r, g, b]]``, flat multiples of 3 (or 4) become rows of triplets
    (quadruplets), 1-D / 2-D float arrays that cannot be pixels are taken as
    grayscale; anything else raises ``ValueError``."""
    @wraps(fn)
    def wrapped(arr, *args, **kwargs):
        arr = _as_array(arr)
        shape = arr.shape
        ndim = len(shape)
        if ndim >= 2 and shape[-1] == 3:                 # an RGB image or a batch of them: nothing to fix
            return fn(arr, *args, **kwargs)
        shape = tuple(shape)
        size = math.prod(shape) if shape else 1
        channels = shape[-1] if shape else 0
        bad = ValueError("Unrecognized image shape: {}".format(shape))
        if ndim == 1:
            if size in (3, 4):
                arr = arr.reshape((1, size))
            elif size % 3 == 0:
                _alert_flat("Assuming 1D RGB")
                arr = arr.reshape((size // 3, 3))
            elif size % 4 == 0:
                _alert_flat("Assuming 1D RGBA")
                arr = arr.reshape((size // 4, 4))
            elif _is_float(arr):
                _alert_flat("Assuming 1D grayscale")
                _alert_gray("Assuming 1D grayscale")
                arr = from_grayscale(arr)
            else:
                raise bad
        elif ndim == 2 and channels not in (3, 4):
            if not _is_float(arr):
                raise bad
            _alert_gray("Assuming 2D grayscale")
            arr = from_grayscale(arr)
        elif channels not in (3, 4):
            raise bad
        return fn(arr, *args, **kwargs)
    return wrapped


def reshape_husl_input(fn):
    """HUSL-side entry guard: last axis must hold (H, S, L)."""
    @wraps(fn)
    def wrapped(arr, *args, **kwargs):
        arr = _as_array(arr)
        shape = arr.shape
        if len(shape) >= 2 and shape[-1] == 3:
            return fn(arr, *args, **kwargs)
        shape = tuple(shape)
        size = math.prod(shape) if shape else 1
        bad = ValueError("Unrecognized HSL shape: {}".format(shape))
        if len(shape) == 1:
            if size == 3:
                arr = arr.reshape((1, 3))
            elif size % 3 == 0:
                _alert_flat("Assuming 1D HSL")
                arr = arr.reshape((size // 3, 3))
            else:
                raise bad
        elif not shape or shape[-1] != 3:
            raise bad
        return fn(arr, *args, **kwargs)
    return wrapped


def reshape_rgba_input(fn):
    """4-channel input is premultiplied by its alpha (i.e. composited over
    black, reference ``transform.py:221-230``) and handed on as float RGB in
    [0, 1].  Integer RGBA becomes ``c/255 * a/255`` -- the value the CUDA
    kernels compute for a uint8 RGBA device array -- NOT the reference's
    accidental *float64 values in 0..255* (App. D-5: its ``astype(arr.dtype)``
    runs after ``arr`` was rebound to the float product, so every backend then
    reads 0..255 as if it were [0, 1]; only float RGBA is tested there)."""
    @wraps(fn)
    def wrapped(arr, *args, **kwargs):
        if arr.shape[-1] == 4 and len(arr.shape) in (2, 3):
            _alert_rgba("Assumed RGBA with white background")
            if is_device_array(arr):
                return fn(RgbaDevice(arr), *args, **kwargs)
            if np.issubdtype(arr.dtype, np.integer):
                arr = (arr[..., :3] / 255.0) * (arr[..., 3] / 255.0)[..., None]
            else:
                arr = arr[..., :3] * arr[..., 3][..., None]
        return fn(arr, *args, **kwargs)
    return wrapped


def squeeze_output(fn):
    """``[[h, s, l]]`` -> ``[h, s, l]`` for single-triplet calls."""
    @wraps(fn)
    def wrapped(*args, **kwargs):
        out = fn(*args, **kwargs)
        if isinstance(out, np.ndarray) and out.size == 3 and out.ndim == 2:
            out = np.squeeze(out)
        return out
    return wrapped


# --------------------------------------------------------------------------
# chunked application
# --------------------------------------------------------------------------

def chunk(end: int, chunksize: int):
    """(start, stop) pairs that tile ``range(end)`` in steps of ``chunksize``;
    the last pair is short when ``chunksize`` does not divide ``end``."""
    start = 0
    if chunksize and end > chunksize:
        for stop in range(chunksize, end, chunksize):
            yield start, stop
            start = stop
    yield start, end


def chunk_img(img, chunksize: int = None):
    """Yield ``(tile, ((r0, r1), (c0, c1)))`` over ``chunksize`` x ``chunksize``
    square tiles of the first two axes (views, not copies).  Without a
    ``chunksize`` the whole image is the single tile."""
    rows, cols = img.shape[:2]
    if not chunksize:
        yield img, ((0, rows), (0, cols))
        return
    for r0, r1 in chunk(rows, chunksize):
        for c0, c1 in chunk(cols, chunksize):
            yield img[r0:r1, c0:c1], ((r0, r1), (c0, c1))


def _chunk_no_idx(img, chunksize: int = None):
    for tile, _ in chunk_img(img, chunksize):
        yield tile


def chunk_many(*arrays, chunksize: int = None):
    """Walk several same-shaped images tile by tile, in lock step."""
    return zip(*(_chunk_no_idx(a, chunksize) for a in arrays))


def chunk_apply(transform, chunks, out) -> None:
    """``out[tile] = transform(tile)`` for every (tile, bounds) of ``chunks``."""
    for tile, ((r0, r1), (c0, c1)) in chunks:
        out[r0:r1, c0:c1] = transform(tile)


def chunk_apply_1d(transform, chunks, out) -> None:
    """Row-only variant for 1-D outputs."""
    for tile, ((r0, r1), _) in chunks:
        out[r0:r1] = transform(tile)


def in_chunks(img, transform: callable, chunksize=None, out=None, **kwargs):
    """Apply ``transform`` to ``img``, optionally tile by tile.

    * no ``chunksize``: ``transform(img)``, copied into ``out`` when given;
    * ``chunksize``: the image is walked in square tiles (``chunk_img``) so the
      transform's temporaries -- for the CUDA backend: the device staging
      buffers -- stay bounded by one tile; results land in ``out``, which is
      allocated here from the first tile's result when the caller gave none.

    The reference declares ``(img, transform, out, chunksize)`` but is called
    with ``(img, transform, chunksize, out)``; an array in the ``chunksize``
    slot / an int in the ``out`` slot is therefore understood as the other order.
    Extra keyword arguments (``dtype=``, ``stream=``, ``mode=`` ... of the CUDA
    backend) are handed to ``transform`` unchanged.
    """
    if hasattr(chunksize, "shape") and (out is None or np.isscalar(out)):
        chunksize, out = out, chunksize
    if not chunksize:
        if out is None:
            return transform(img, **kwargs)
        if getattr(transform, "accepts_out", False):
            # the backend writes into `out` itself (host or device memory)
            return transform(img, out=out, **kwargs)
        if is_device_array(out):
            raise ValueError("device `out` arrays can only be filled by a backend "
                             "that accepts `out` (the CUDA backend)")
        out[...] = transform(img, **kwargs)
        return out
    if is_device_array(img):
        raise ValueError("`chunksize` applies to host arrays only: a device-"
                         "resident frame is already where the kernels read it")
    chunks = chunk_img(img, chunksize)
    if out is None:
        # learn dtype and trailing shape from the first tile
        first, ((r0, r1), (c0, c1)) = next(chunks)
        res = np.asarray(transform(first, **kwargs))
        lead = img.shape[:1] if res.ndim == 1 else img.shape[:2]
        out = np.empty(tuple(lead) + res.shape[2:] if res.ndim > 1 else lead,
                       dtype=res.dtype)
        if res.ndim == 1:
            out[r0:r1] = res
        else:
            out[r0:r1, c0:c1] = res
    if kwargs:
        transform = partial(transform, **kwargs)
    apply = chunk_apply_1d if len(out.shape) == 1 else chunk_apply
    apply(transform, chunks, out)
    return out
