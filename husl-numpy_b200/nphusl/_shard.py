"""Sharding of frame batches / row bands over the GPUs of one box.

Every pixel is independent (no halo, no reduction), so multi-GPU here is pure
partitioning: contiguous blocks of the leading axis (frames of a batch, or rows
of one large image) go to each GPU, every GPU runs the same kernels through
the same C ABI on its own stream and staging buffers, and the results land in
disjoint slices of the caller's output.  There is NO collective on the data
path and no NCCL dependency.

Three ways to use it:

* one process, many GPUs -- ``to_husl`` / ``to_rgb`` / ``to_hue`` below fan a
  host array out with one host thread per GPU (ctypes releases the GIL while
  the C ABI streams the shard through pinned staging buffers);
* one process per GPU (``torchrun``; what ``bench.py --gpus N`` does) --
  ``shard_bounds(n, world)[rank]`` gives each rank its block;
* a frame STREAM (a video: BASELINE.json configs[4]) -- ``stream_frames`` walks a cyclic pool
  of host frames in batches, each GPU converting its contiguous block of the stream on two
  alternating CUDA streams into a ring of page-locked result batches.
"""
import threading
import time

import numpy as np

from . import _cuda

__all__ = ["shard_bounds", "to_husl", "to_rgb", "to_hue", "frame_batches", "stream_frames"]


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


# ---------------------------------------------------------------------------
# frame streams (video): cyclic pool of host frames -> GPUs -> ring of host results
# ---------------------------------------------------------------------------

def frame_batches(start: int, stop: int, pool_len: int, batch: int):
    """Cut frames ``[start, stop)`` of a stream whose frame ``i`` is ``pool[i % pool_len]`` into
    batches that are contiguous IN THE POOL: yields ``(first_frame, pool_index, n)`` with
    ``1 <= n <= batch`` and ``pool_index + n <= pool_len``."""
    if pool_len < 1 or batch < 1:
        raise ValueError("pool_len and batch must be >= 1")
    i = int(start)
    while i < stop:
        p = i % pool_len
        n = min(batch, stop - i, pool_len - p)
        yield i, p, n
        i += n


def stream_frames(pool, n_frames, devices=None, first_frame=0, batch=8, dtype=np.float64, edit=None,
                  ready=None, on_batch=None, schedule="static"):
    """Run ``to_husl`` -> (``edit``) -> ``to_rgb`` over ``n_frames`` frames of a stream whose frame
    ``i`` is ``pool[i % len(pool)]`` (``pool``: ``(P, H, W, 3)`` uint8 host array, page-locked for
    direct DMA).  The stream is cut into contiguous blocks, one per device; each device is driven
    by its own host thread with two CUDA streams, two device HUSL buffers and a ring of two
    page-locked result batches, so the upload of batch k+1 overlaps the download of batch k.
    Every frame's uint8 result lands in host memory (the ring slot is then free for ``on_batch``
    -- a sink such as an encoder -- or simply overwritten two batches later).

    ``schedule``: ``"static"`` gives every device one contiguous block of the stream
    (``shard_bounds``); ``"dynamic"`` hands the batches out first come, first served, in stream order,
    each device keeping at most two in flight -- the GPUs of one box do not all see the same host link
    (PCIe switch / NUMA placement), and the faster ones then convert more of the stream.

    ``ready``: called once from the calling thread when every device has allocated its buffers and
    run one untimed batch, right before the timed region starts (a cross-process barrier, when
    several processes each drive their own GPU).  ``on_batch(device, first_frame, result_view)`` is
    called from the device's thread after the batch has arrived (this serialises that device's
    two streams; leave it ``None`` to measure the pipeline alone).

    -> ``{"seconds": wall time of the timed region, "frames": n_frames,
          "per_device": {dev: {"frames", "seconds", "batches"}},
          "last": {dev: (first_frame, pool_index, n, ndarray view of that batch's result)}}``"""
    pool = np.asarray(pool)
    if pool.ndim != 4 or pool.shape[-1] != 3 or pool.dtype != np.uint8:
        raise ValueError("pool: expected (P, H, W, 3) uint8 frames, got {} {}".format(pool.dtype, pool.shape))
    devices = _devices(devices)
    if schedule not in ("static", "dynamic"):
        raise ValueError("schedule must be 'static' or 'dynamic'")
    dynamic = schedule == "dynamic"
    if dynamic:
        blocks = [(d, first_frame, first_frame + n_frames) for d in devices] if n_frames > 0 else []
        shared = frame_batches(first_frame, first_frame + n_frames, len(pool), batch)
        shared_lock = threading.Lock()
    else:
        blocks = [(d, first_frame + a, first_frame + b)
                  for d, (a, b) in zip(devices, shard_bounds(n_frames, len(devices))) if b > a]
    frame_shape = tuple(pool.shape[1:])
    setup_done = threading.Barrier(len(blocks) + 1)
    go = threading.Barrier(len(blocks) + 1)
    stats, last, errors = {}, {}, []

    def work(dev, a, b):
        try:
            streams = [_cuda.Stream(dev), _cuda.Stream(dev)]
            hs = [_cuda.DeviceArray((batch,) + frame_shape, dtype, dev, streams[k]) for k in range(2)]
            ring = _cuda.pinned_empty((2, batch) + frame_shape, np.uint8)

            def submit(k, p, n):
                s = streams[k & 1]
                h = hs[k & 1][:n]
                _cuda._rgb_to_husl(pool[p:p + n], out=h, stream=s, device=dev, blocking=False)
                _cuda._husl_to_rgb(h, out=ring[k & 1][:n], stream=s, device=dev, blocking=False, edit=edit)
                return s
            for k in range(2):                               # first-use costs stay out of the timed region
                submit(k, 0, min(batch, len(pool))).synchronize()
        except BaseException as e:                           # noqa: BLE001 - re-raised in the caller's thread
            errors.append(e)
            setup_done.abort()
            return
        try:
            setup_done.wait()
            go.wait()
        except threading.BrokenBarrierError:
            return
        try:
            t0 = time.perf_counter()
            k, rec, frames = 0, None, 0
            own = None if dynamic else frame_batches(a, b, len(pool), batch)
            while True:
                if dynamic:
                    if k >= 2:
                        streams[k & 1].synchronize()         # at most two batches in flight: the ring slot is free again
                    with shared_lock:
                        item = next(shared, None)
                else:
                    item = next(own, None)
                if item is None:
                    break
                i, p, n = item
                s = submit(k, p, n)
                rec = (i, p, n, k & 1)
                frames += n
                if on_batch is not None:
                    s.synchronize()
                    on_batch(dev, i, ring[k & 1][:n])
                k += 1
            for s in streams:
                s.synchronize()
            stats[dev] = {"frames": frames, "seconds": time.perf_counter() - t0, "batches": k}
            if rec is not None:
                last[dev] = (rec[0], rec[1], rec[2], ring[rec[3]][:rec[2]])
        except BaseException as e:                           # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=blk) for blk in blocks]
    for t in threads:
        t.start()
    try:
        setup_done.wait()
        if ready is not None:
            try:
                ready()
            except BaseException as e:                       # noqa: BLE001 - workers must not wait for a start that never comes
                errors.append(e)
                go.abort()
        go.wait()
    except threading.BrokenBarrierError:
        pass
    t0 = time.perf_counter()
    for t in threads:
        t.join()
    seconds = time.perf_counter() - t0
    if errors:
        raise errors[0]
    return {"seconds": seconds, "frames": int(n_frames), "per_device": stats, "last": last}
