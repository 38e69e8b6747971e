"""Sharding of frame batches / row bands over the GPUs of one box.

Every pixel is independent (no halo, no reduction), so multi-GPU here is pure
partitioning: contiguous blocks of the leading axis (frames of a batch, or rows
of one large image) go to each GPU, every GPU runs the same kernels through
the same C ABI on its own stream and staging buffers, and the results land in
disjoint slices of the caller's output.  There is NO collective on the data
path and no NCCL dependency.

Two ways to use it:

* one process, many GPUs -- ``to_husl`` / ``to_rgb`` / ``to_hue`` below fan a
  host array out with one host thread per GPU (ctypes releases the GIL while
  the C ABI streams the shard through pinned staging buffers);
* one process per GPU (``torchrun``; what ``bench.py --gpus N`` does) --
  ``shard_bounds(n, world)[rank]`` gives each rank its block.
"""
import threading

import numpy as np

from . import _cuda

__all__ = ["shard_bounds", "to_husl", "to_rgb", "to_hue"]


def shard_bounds(n_items: int, parts: int):
    """``parts`` contiguous ``(start, stop)`` blocks covering ``range(n_items)``;
    sizes differ by at most one, empty blocks only when ``parts > n_items``."""
    if parts < 1:
        raise ValueError("parts must be >= 1")
    base, extra = divmod(int(n_items), parts)
    bounds, start = [], 0
    for r in range(parts):
        stop = start + base + (1 if r < extra else 0)
        bounds.append((start, stop))
        start = stop
    return bounds


def _devices(devices):
    if devices is None:
        devices = list(range(_cuda.device_count()))
    devices = list(devices)
    if not devices:
        raise _cuda.CudaUnavailable("no CUDA device to shard over (there is no CPU fallback)")
    return devices


def _fan_out(fn, src, out, devices, **kw):
    """Run ``fn(src[a:b], out=out[a:b], device=d)`` for each device's block, one
    host thread per device; re-raise the first failure."""
    devices = _devices(devices)
    blocks = [(d, a, b) for d, (a, b) in zip(devices, shard_bounds(src.shape[0], len(devices))) if b > a]
    errors = []

    def work(dev, a, b):
        try:
            fn(src[a:b], out=out[a:b], device=dev, **kw)
        except BaseException as e:          # noqa: BLE001 - re-raised in the caller's thread
            errors.append(e)

    if len(blocks) == 1:
        work(*blocks[0])
    else:
        threads = [threading.Thread(target=work, args=blk) for blk in blocks]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return out


def _host(arr, what):
    if hasattr(arr, "__cuda_array_interface__"):
        raise ValueError("{}: sharding fans HOST arrays out over GPUs; a device array already "
                         "lives on one GPU -- convert it there".format(what))
    arr = np.ascontiguousarray(arr)
    if arr.ndim < 2 or arr.shape[-1] != 3:
        raise ValueError("{}: expected (n, ..., 3), got {}".format(what, arr.shape))
    return arr


def to_husl(rgb, out=None, devices=None, dtype=np.float64, mode="exact"):
    """RGB ``(n, ..., 3)`` host array -> HUSL, leading axis sharded over ``devices``."""
    rgb = _host(rgb, "rgb")
    if out is None:
        out = np.empty(rgb.shape, dtype)
    return _fan_out(_cuda._rgb_to_husl, rgb, out, devices, mode=mode)


def to_rgb(hsl, out=None, devices=None):
    """HUSL ``(n, ..., 3)`` host array -> uint8 RGB, leading axis sharded."""
    hsl = _host(hsl, "hsl")
    if out is None:
        out = np.empty(hsl.shape, np.uint8)
    return _fan_out(_cuda._husl_to_rgb, hsl, out, devices)


def to_hue(rgb, out=None, devices=None, dtype=np.float64):
    """RGB ``(n, ..., 3)`` host array -> hue ``(n, ...)``, leading axis sharded."""
    rgb = _host(rgb, "rgb")
    if out is None:
        out = np.empty(rgb.shape[:-1], dtype)
    return _fan_out(_cuda._rgb_to_hue, rgb, out, devices)
