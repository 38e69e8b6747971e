"""CUDA backend of nphusl: a thin ctypes binding of ``libnphusl_cuda.so``
(C ABI in ``include/nphusl_cuda.h``, kernels in ``csrc/``).

A backend, in the reference, is a module that exposes callables named
``_rgb_to_husl``, ``_husl_to_rgb`` and ``_rgb_to_hue`` (harvested by
``nphusl/nphusl.py:84-101``); this module is the fifth one, next to NumPy,
NumExpr, Cython (``nphusl/_cython_opt.pyx:22,96,177``) and C/SIMD
(``nphusl/_simd_opt.pyx:20-33``).  Contract kept from those:

* ``_rgb_to_husl(rgb)``  ``(..., 3)`` uint8 / float RGB -> float64 HUSL, same shape
* ``_rgb_to_hue(rgb)``   ``(..., 3)`` RGB -> float64 hue, shape ``rgb.shape[:-1]``
* ``_husl_to_rgb(hsl)``  ``(..., 3)`` float HUSL -> **uint8** RGB (the
  ``to_rgb_int`` tail of ``transform.py:16-20`` is fused into the kernel; a
  uint8 return passes ``rgb_int_output`` untouched, ``transform.py:57-61``)

Added by this backend:

* any object exposing ``__cuda_array_interface__`` (torch CUDA tensors, the
  ``DeviceArray`` below) is converted where it lives and a ``DeviceArray`` is
  returned: device-resident frames never touch the host;
* ``out=`` (host or device), ``dtype=`` (float32 halves the HUSL bytes),
  ``mode=`` ("exact" | "lut" | "lut_interp"), ``stream=``, ``device=``,
  ``blocking=False`` (pinned host buffers: enqueue only, caller syncs ``stream``).

There is NO CPU implementation here: when the shared library or a GPU is
missing, importing works but ``available()`` is False and every conversion
raises ``CudaUnavailable``.
"""
import ctypes
import math
import os
import threading

import numpy as np

__all__ = ["_rgb_to_husl", "_husl_to_rgb", "_rgb_to_hue", "rgb_to_husl_planar", "husl_planar_to_rgb",
           "edit_rgb", "edit_husl", "Edit", "Bands", "Graph", "graph", "nv12_to_husl", "nv12_to_hue", "nv12_to_rgb", "nv12_edit_rgb", "rgb_to_nv12", "nv12_edit_nv12", "DeviceArray", "available", "CudaUnavailable", "CudaError"]

def _prod(shape):
    """Number of elements of a shape (np.prod on a tuple costs 3 us and runs several times per call)."""
    return math.prod(shape) if hasattr(shape, "__iter__") else int(shape)


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NPHUSL_CUDA_LIB", os.path.join(_HERE, "libnphusl_cuda.so"))

U8, F32, F64 = 0, 1, 2
HOST, DEVICE = 0, 1
MODES = {"exact": 0, "lut": 1, "lut_interp": 2}
_DT = {np.dtype(np.uint8): U8, np.dtype(np.float32): F32, np.dtype(np.float64): F64}


class CudaUnavailable(RuntimeError):
    """libnphusl_cuda.so could not be loaded or there is no usable GPU."""


class CudaError(RuntimeError):
    """A call into libnphusl_cuda.so returned a non-zero status."""


# every exported symbol of include/nphusl_cuda.h: (restype, argtypes)
_vp, _sz, _i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
_ip, _dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
_u16p = ctypes.POINTER(ctypes.c_uint16)
SIGNATURES = {
    "nphusl_cuda_abi_version": (_i, []),
    "nphusl_cuda_device_count": (_i, []),
    "nphusl_cuda_last_error": (ctypes.c_char_p, []),
    "nphusl_cuda_launch_count": (ctypes.c_ulonglong, []),
    "nphusl_cuda_set_chunk_pixels": (_i, [_sz]),
    "nphusl_cuda_set_host_async": (_i, [_i]),
    "nphusl_cuda_rgb_to_husl": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _sz, _i, _i, _vp]),
    "nphusl_cuda_rgb_to_hue": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _sz, _i, _vp]),
    "nphusl_cuda_husl_to_rgb": (_i, [_vp, _i, _i, _vp, _i, _i, _sz, _i, _vp]),
    "nphusl_cuda_rgb_to_husl_planar": (_i, [_vp, _vp, _vp, _vp, _i, _sz, _i, _vp]),
    "nphusl_cuda_husl_planar_to_rgb": (_i, [_vp, _vp, _vp, _i, _vp, _sz, _i, _vp]),
    "nphusl_cuda_rgb_planes_to_husl_planes": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _sz, _i, _vp]),
    "nphusl_cuda_husl_planes_to_rgb_planes": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "nphusl_cuda_rgb_edit_rgb": (_i, [_vp, _i, _vp, _i, _sz, _vp, _i, _vp]),
    "nphusl_cuda_husl_edit_to_rgb": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _sz, _i, _vp]),
    "nphusl_cuda_husl_bands_to_rgb": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _sz, _i, _vp]),
    "nphusl_cuda_rgb_bands_rgb": (_i, [_vp, _i, _vp, _i, _sz, _vp, _i, _vp]),
    "nphusl_cuda_husl_edit": (_i, [_vp, _vp, _i, _sz, _vp, _i, _vp]),
    "nphusl_cuda_rgb_planes_edit": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "nphusl_cuda_rgb_to_husl_strided": (_i, [_vp, _vp, _i, _i, _vp]),
    "nphusl_cuda_husl_to_rgb_strided": (_i, [_vp, _vp, _i, _vp]),
    "nphusl_cuda_rgb_edit_rgb_strided": (_i, [_vp, _vp, _vp, _i, _vp]),
    "nphusl_cuda_nv12_to_rgb": (_i, [_vp, _vp, _i, _vp]),
    "nphusl_cuda_nv12_to_husl": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "nphusl_cuda_nv12_edit_rgb": (_i, [_vp, _vp, _vp, _i, _vp]),
    "nphusl_cuda_rgb_to_nv12": (_i, [_vp, _vp, _i, _vp]),
    "nphusl_cuda_nv12_edit_nv12": (_i, [_vp, _vp, _vp, _i, _vp]),
    "nphusl_cuda_lut_generate": (_i, [_i, _i, _i, _i]),
    "nphusl_cuda_lut_upload": (_i, [_i, _u16p, _i, _dp, _dp, _u16p, _i, _i,
                                    ctypes.c_double, ctypes.c_double, _dp]),
    "nphusl_cuda_lut_download": (_i, [_i, _u16p, _u16p, _dp, _ip, _dp]),
    "nphusl_cuda_lut_generate_typed": (_i, [_i, _i, _i, _i, _i, _i]),
    "nphusl_cuda_lut_upload_typed": (_i, [_i, _i, ctypes.c_double, ctypes.c_double, _vp, _i, _dp, _dp, _vp, _i, _i,
                                          ctypes.c_double, ctypes.c_double, _dp]),
    "nphusl_cuda_lut_download_typed": (_i, [_i, _vp, _vp, _dp, _ip, _dp, _dp]),
    "nphusl_cuda_malloc": (_i, [ctypes.POINTER(_vp), _sz, _i]),
    "nphusl_cuda_free": (_i, [_vp, _i]),
    "nphusl_cuda_host_alloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "nphusl_cuda_host_free": (_i, [_vp]),
    "nphusl_cuda_memcpy": (_i, [_vp, _vp, _sz, _i, _i, _vp]),
    "nphusl_cuda_stream_create": (_i, [ctypes.POINTER(_vp), _i]),
    "nphusl_cuda_stream_destroy": (_i, [_vp, _i]),
    "nphusl_cuda_stream_synchronize": (_i, [_vp, _i]),
    "nphusl_cuda_device_synchronize": (_i, [_i]),
    "nphusl_cuda_event_record": (_i, [ctypes.POINTER(_vp), _vp, _i]),
    "nphusl_cuda_event_query": (_i, [_vp, _ip]),
    "nphusl_cuda_stream_wait_event": (_i, [_vp, _vp, _i]),
    "nphusl_cuda_event_destroy": (_i, [_vp, _i]),
    "nphusl_cuda_graph_begin": (_i, [_vp, _i]),
    "nphusl_cuda_graph_end": (_i, [_vp, _i, ctypes.POINTER(_vp)]),
    "nphusl_cuda_graph_launch": (_i, [_vp, _vp, _i]),
    "nphusl_cuda_graph_destroy": (_i, [_vp, _i]),
    "nphusl_cuda_pointer_info": (_i, [_vp, _ip, _ip, _ip]),
    "nphusl_cuda_device_info": (_i, [_i, ctypes.c_char_p, _sz, _ip, ctypes.POINTER(_sz)]),
}

_lib = None
_lib_error = None
_lock = threading.Lock()
_default_device = 0


def load():
    """dlopen the shared library once and type every entry point.  Raises
    ``CudaUnavailable`` (loudly, no fallback) when it cannot be loaded."""
    global _lib, _lib_error
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        try:
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
        except (OSError, AttributeError) as e:
            _lib_error = e
            raise CudaUnavailable("cannot load {}: {} -- build it with `make -C husl-numpy_b200/csrc` "
                                  "(there is no CPU fallback in the CUDA backend)".format(LIB_PATH, e))
        _lib = lib
    return _lib


def device_count() -> int:
    try:
        return load().nphusl_cuda_device_count()
    except CudaUnavailable:
        return 0


def available() -> bool:
    """True when the library loads AND at least one GPU is visible."""
    return device_count() > 0


def _check(rc, what):
    if rc:
        msg = load().nphusl_cuda_last_error().decode("utf-8", "replace")
        if rc == 3:
            raise CudaUnavailable("{}: {}".format(what, msg))
        raise CudaError("{} failed (status {}): {}".format(what, rc, msg))


def set_device(device: int):
    """Default GPU for host-array calls and new allocations of this process."""
    global _default_device
    _default_device = int(device)


def get_device() -> int:
    return _default_device


def launch_count() -> int:
    return int(load().nphusl_cuda_launch_count())


def set_chunk_pixels(pixels: int):
    """Pixels per staging slot of the host-array streaming pipeline (0 = default)."""
    _check(load().nphusl_cuda_set_chunk_pixels(int(pixels)), "set_chunk_pixels")


def synchronize(device=None, stream=None):
    dev = _default_device if device is None else device
    if stream:
        _check(load().nphusl_cuda_stream_synchronize(_stream_ptr(stream), dev), "stream_synchronize")
    else:
        _check(load().nphusl_cuda_device_synchronize(dev), "device_synchronize")


def device_info(device=None):
    dev = _default_device if device is None else device
    name = ctypes.create_string_buffer(256)
    sms, mem = ctypes.c_int(), ctypes.c_size_t()
    _check(load().nphusl_cuda_device_info(dev, name, 256, ctypes.byref(sms), ctypes.byref(mem)), "device_info")
    return {"name": name.value.decode(), "sm_count": sms.value, "total_mem": mem.value}


def _stream_ptr(stream):
    """None | int | object with .cuda_stream (torch.cuda.Stream) | Stream below."""
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):
        return ctypes.c_void_p(stream.cuda_stream)
    if hasattr(stream, "ptr"):
        return ctypes.c_void_p(stream.ptr)
    return ctypes.c_void_p(int(stream))


def _stream_key(stream) -> int:
    """The stream's handle as an int (0 = the NULL / legacy default stream)."""
    if stream is None:
        return 0
    if hasattr(stream, "cuda_stream"):
        return int(stream.cuda_stream)
    if hasattr(stream, "ptr"):
        return int(stream.ptr or 0)
    return int(stream)


class Stream:
    """A non-blocking CUDA stream owned through the C ABI."""

    def __init__(self, device=None):
        self.device = _default_device if device is None else device
        p = ctypes.c_void_p()
        _check(load().nphusl_cuda_stream_create(ctypes.byref(p), self.device), "stream_create")
        self.ptr = p.value or 0
        _device_pool.revive(self.ptr)

    def synchronize(self):
        _check(load().nphusl_cuda_stream_synchronize(ctypes.c_void_p(self.ptr), self.device), "stream_synchronize")

    def close(self):
        """Wait for the stream's work, then destroy it.  The wait is what lets the block pool forget the
        stream: nothing that ran on it can still be using a device block, and its handle is never
        handed to the driver again (a destroyed handle is a dangling pointer)."""
        if self.ptr:
            lib = load()
            lib.nphusl_cuda_stream_synchronize(ctypes.c_void_p(self.ptr), self.device)
            _device_pool.retire(self.ptr)
            lib.nphusl_cuda_stream_destroy(ctypes.c_void_p(self.ptr), self.device)
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Graph:
    """A recorded sequence of device-array calls, replayed with ONE call (a CUDA graph).

    Per-frame loops are bound by the host side of a call -- a 1080p frame is 10 us of kernel time and
    a call costs about as much -- so the loop body is recorded once and replayed::

        g = nphusl._cuda.Graph()
        with g.capture() as s:                       # s: the stream to pass to the recorded calls
            nphusl.to_husl(frame, out=hsl, stream=s)
            nphusl.to_rgb(hsl, out=back, stream=s, edit=e)
        for f in frames:
            frame.copy_(f)                           # the recorded calls read / write the SAME buffers
            g.replay()                               # on g.stream (or replay(stream=...))
        g.synchronize()

    Only calls whose arrays all live on the device can be recorded, every array they use must be passed
    in (``out=``: results of a recorded call cannot be allocated while recording) and must stay alive as
    long as the graph is replayed.  ``graph(fn)`` records ``fn(stream)`` in one line."""

    def __init__(self, device=None, stream=None):
        self.device = _default_device if device is None else device
        self.stream = Stream(self.device) if stream is None else stream
        self._exec = 0
        self._keep = []

    def capture(self):
        return _GraphCapture(self)

    def replay(self, stream=None):
        if not self._exec:
            raise CudaError("Graph.replay before capture")
        _check(load().nphusl_cuda_graph_launch(ctypes.c_void_p(self._exec), _stream_ptr(self.stream if stream is None else stream),
                                               self.device), "graph_launch")

    __call__ = replay

    def synchronize(self):
        synchronize(self.device, self.stream)

    def close(self):
        if self._exec and _lib is not None:
            _lib.nphusl_cuda_graph_destroy(ctypes.c_void_p(self._exec), self.device)
        self._exec = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _GraphCapture:
    def __init__(self, g):
        self.g = g

    def __enter__(self):
        g = self.g
        g.close()
        if not _stream_key(g.stream):
            raise ValueError("Graph: capture needs a stream of its own, not the NULL stream")
        _check(load().nphusl_cuda_graph_begin(_stream_ptr(g.stream), g.device), "graph_begin")
        return g.stream

    def __exit__(self, exc_type, exc, tb):
        g = self.g
        e = ctypes.c_void_p()
        rc = load().nphusl_cuda_graph_end(_stream_ptr(g.stream), g.device, ctypes.byref(e))
        if exc_type is None:
            _check(rc, "graph_end")
            g._exec = e.value or 0
        elif e.value:
            load().nphusl_cuda_graph_destroy(e, g.device)
        return False


def graph(fn, device=None, stream=None):
    """Record ``fn(stream)`` -- a function issuing device-array calls on the stream it is given -- and
    return the ``Graph``; ``g.replay()`` (or ``g()``) re-runs it."""
    g = Graph(device, stream)
    with g.capture() as s:
        fn(s)
    return g


class _DevicePool:
    """Recycles device blocks of dropped ``DeviceArray``s.  cudaMalloc + cudaFree cost 0.6 ms per
    result (cudaFree also synchronises the device), sixty times the conversion of a 1080p frame;
    with the pool a call that returns a new DeviceArray costs what a call with ``out=`` costs.

    Reuse is ordered by streams, not by luck.  Every ``DeviceArray`` remembers the streams the
    library used it on.  A block released after use on ONE stream goes to that stream's free list
    and is handed out again only to a call on the same stream (stream order makes that safe), or to
    a call on another stream after that stream has been made to wait on an event recorded behind
    the block's last use.  A block that was used on SEVERAL streams is parked with one event per
    stream and becomes reusable (by any stream) when all of them have completed.  Work the library
    cannot see -- a block exported through ``__cuda_array_interface__`` and used by other code on
    its own stream -- must be announced with ``DeviceArray.record_stream(stream)`` before the
    array is dropped.  Streams are named by their handle: a handle the pool still remembers must stay
    valid.  The library's own ``Stream.close()`` drains the stream and tells the pool to forget it;
    torch's streams live in a per-device pool that is never destroyed; whoever passes raw handles of
    their own keeps them alive (or calls ``empty_cache()`` before destroying them)."""

    ANY = -1                             # free-list key of blocks that are idle on every stream

    def __init__(self):
        self.max_cached = int(os.environ.get("NPHUSL_DEVICE_CACHE_MB", "16384")) << 20
        self.free = {}                   # (device, cap) -> {stream_key: [ptr, ...]}
        self.parked = []                 # [ptr, cap, device, [event, ...]]: in use on several streams when dropped
        self.events = {}                 # device -> [event, ...] spare events
        self.cached = 0
        self.retired = set()             # handles of streams that were drained and destroyed (Stream.close)
        self.lock = threading.RLock()    # re-entrant: a GC pass inside alloc() may finalise a DeviceArray -> release()
        self.hits = self.misses = self.waits = 0

    def retire(self, stream_key):
        """The stream was synchronised and is about to be destroyed: its blocks are idle everywhere."""
        with self.lock:
            self.retired.add(stream_key)
            for by in self.free.values():
                lst = by.pop(stream_key, None)
                if lst:
                    by.setdefault(self.ANY, []).extend(lst)

    def revive(self, stream_key):
        """A new stream got a handle value an old one had."""
        with self.lock:
            self.retired.discard(stream_key)

    @staticmethod
    def size_class(nbytes):
        if nbytes <= 512:
            return 512
        if nbytes < (1 << 20):
            return 1 << (nbytes - 1).bit_length()
        return (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)

    def _event(self, device, stream_key):
        """-> an event recorded now on `stream_key` (0 when that stream no longer takes work)"""
        spare = self.events.get(device)
        ev = ctypes.c_void_p(spare.pop() if spare else None)
        if load().nphusl_cuda_event_record(ctypes.byref(ev), ctypes.c_void_p(stream_key or None), device):
            if ev.value:
                self.events.setdefault(device, []).append(ev.value)
            return 0
        return ev.value

    def _reap(self):
        """parked blocks whose events have all completed become reusable by any stream"""
        if not self.parked:
            return
        lib, done, still = load(), ctypes.c_int(), []
        for ent in self.parked:
            ptr, cap, device, evs = ent
            while evs:
                if lib.nphusl_cuda_event_query(ctypes.c_void_p(evs[-1]), ctypes.byref(done)) or not done.value:
                    break
                self.events.setdefault(device, []).append(evs.pop())
            if evs:
                still.append(ent)
            else:
                self.free.setdefault((device, cap), {}).setdefault(self.ANY, []).append(ptr)
        self.parked = still

    def alloc(self, nbytes, device, stream_key=0):
        cap = self.size_class(nbytes)
        with self.lock:
            self._reap()
            by_stream = self.free.get((device, cap))
            if by_stream:
                for key in (stream_key, self.ANY):
                    lst = by_stream.get(key)
                    if lst:
                        self.cached -= cap
                        self.hits += 1
                        return lst.pop(), cap
                # a block last used on ANOTHER stream: this stream waits for that stream's tail
                for key, lst in by_stream.items():
                    if not lst:
                        continue
                    ev = self._event(device, key)
                    if not ev:
                        continue                 # that stream is gone; its blocks are reclaimed by trim()
                    ok = load().nphusl_cuda_stream_wait_event(ctypes.c_void_p(stream_key or None), ctypes.c_void_p(ev), device) == 0
                    self.events.setdefault(device, []).append(ev)
                    if ok:
                        self.cached -= cap
                        self.hits += 1
                        self.waits += 1
                        return lst.pop(), cap
            self.misses += 1
        p = ctypes.c_void_p()
        rc = load().nphusl_cuda_malloc(ctypes.byref(p), cap, device)
        if rc:                           # out of memory: give the cache back and try once more
            self.trim()
            _check(load().nphusl_cuda_malloc(ctypes.byref(p), cap, device), "malloc")
        return p.value or 0, cap

    def release(self, ptr, cap, device, streams=()):
        """`streams`: keys of the streams the block was used on"""
        with self.lock:
            if self.cached + cap <= self.max_cached and _lib is not None:
                keys = [k for k in streams if k not in self.retired]
                if len(keys) <= 1:
                    key = keys[0] if keys else self.ANY
                    self.free.setdefault((device, cap), {}).setdefault(key, []).append(ptr)
                    self.cached += cap
                    return
                evs = [self._event(device, k) for k in keys]
                if all(evs):
                    self.parked.append([ptr, cap, device, evs])
                    self.cached += cap
                    return
                self.events.setdefault(device, []).extend(e for e in evs if e)
        if _lib is not None:
            _lib.nphusl_cuda_free(ctypes.c_void_p(ptr), device)      # cudaFree waits for the device: always safe

    def trim(self):
        with self.lock:
            blocks = [(p, dev) for (dev, _), by in self.free.items() for lst in by.values() for p in lst]
            blocks += [(ent[0], ent[2]) for ent in self.parked]
            events = [(e, ent[2]) for ent in self.parked for e in ent[3]]
            events += [(e, dev) for dev, lst in self.events.items() for e in lst]
            self.free.clear()
            self.parked = []
            self.events.clear()
            self.cached = 0
        for p, dev in blocks:
            if _lib is not None:
                _lib.nphusl_cuda_free(ctypes.c_void_p(p), dev)
        for e, dev in events:
            if _lib is not None:
                _lib.nphusl_cuda_event_destroy(ctypes.c_void_p(e), dev)


_device_pool = _DevicePool()


def empty_cache():
    """Return the cached device and page-locked blocks to the driver."""
    _device_pool.trim()
    _pinned_pool.trim()


def device_pool_stats():
    return {"cached_bytes": _device_pool.cached, "hits": _device_pool.hits, "misses": _device_pool.misses,
            "cross_stream_waits": _device_pool.waits, "parked": len(_device_pool.parked)}


class DeviceArray:
    """C-contiguous array in GPU memory, owned through the C ABI (blocks come from a recycling
    pool).  Speaks ``__cuda_array_interface__`` v3, so ``torch.as_tensor(a, device="cuda")``
    wraps it without a copy."""

    def __init__(self, shape, dtype, device=None, stream=None):
        """``stream``: the stream the array will first be used on (its block may have been last
        used there, or is made to wait for its previous user, see ``_DevicePool``)."""
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.device = _default_device if device is None else device
        self.size = _prod(self.shape) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        key = _stream_key(stream)
        self.ptr, self._cap = _device_pool.alloc(self.nbytes, self.device, key)
        self._streams = {key}
        self._owner = True

    def record_stream(self, stream):
        """Tell the pool the array is (also) in use on ``stream`` -- for work enqueued by OTHER code
        (a torch kernel on ``torch.as_tensor(a, device="cuda")``, say); the library's own calls
        record their stream themselves.  The block is not reused before that work has finished."""
        root = self
        while getattr(root, "_base", None) is not None:
            root = root._base
        root._streams.add(_stream_key(stream))

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False),
                "version": 3, "strides": None, "stream": None}

    @classmethod
    def from_host(cls, arr, device=None, stream=None):
        arr = np.ascontiguousarray(arr)
        d = cls(arr.shape, arr.dtype, device, stream)
        _check(load().nphusl_cuda_memcpy(d.ptr, arr.ctypes.data, d.nbytes, 1, d.device, _stream_ptr(stream)), "memcpy h2d")
        return d

    def copy_to_host(self, out=None, stream=None):
        if out is None:
            out = np.empty(self.shape, self.dtype)
        assert out.nbytes == self.nbytes and out.flags.c_contiguous
        self.record_stream(stream)
        _check(load().nphusl_cuda_memcpy(out.ctypes.data, self.ptr, self.nbytes, 2, self.device, _stream_ptr(stream)), "memcpy d2h")
        return out

    def _view(self, shape, byte_offset=0):
        v = object.__new__(DeviceArray)
        v.shape = tuple(int(x) for x in shape)
        v.dtype, v.device = self.dtype, self.device
        v.size = _prod(v.shape) if v.shape else 1
        v.nbytes = v.size * v.dtype.itemsize
        v.ptr = self.ptr + byte_offset
        v._owner = False
        v._base = self                      # keeps the allocation alive
        v._streams = None                   # stream use is recorded on the owning array
        return v

    def reshape(self, *shape):
        if len(shape) == 1 and not np.isscalar(shape[0]):
            shape = tuple(shape[0])
        shape = [int(x) for x in shape]
        if shape.count(-1) > 1:
            raise ValueError("can only specify one unknown dimension")
        if -1 in shape:
            known = _prod([x for x in shape if x != -1]) or 1
            shape[shape.index(-1)] = self.size // known
        if _prod(shape) != self.size:
            raise ValueError("cannot reshape array of size {} into shape {}".format(self.size, tuple(shape)))
        return self._view(shape)

    def __getitem__(self, key):
        """Leading-axis indexing only (``a[i]``, ``a[i:j]``): such views stay C-contiguous."""
        if not self.shape:
            raise IndexError("too many indices")
        row = (self.size // self.shape[0]) * self.dtype.itemsize if self.shape[0] else 0
        if isinstance(key, (int, np.integer)):
            k = int(key) + (self.shape[0] if key < 0 else 0)
            if not 0 <= k < self.shape[0]:
                raise IndexError("index out of range")
            return self._view(self.shape[1:], k * row)
        if isinstance(key, slice):
            a, b, step = key.indices(self.shape[0])
            if step != 1:
                raise IndexError("DeviceArray slices must be contiguous")
            return self._view((max(b - a, 0),) + self.shape[1:], a * row)
        raise IndexError("DeviceArray supports integer and contiguous-slice indexing of the first axis")

    def __array__(self, dtype=None, copy=None):
        a = self.copy_to_host()
        return a if dtype is None else a.astype(dtype)

    def free(self):
        if getattr(self, "_owner", False) and self.ptr:
            _device_pool.release(self.ptr, self._cap, self.device, self._streams)
            self._owner = False
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _PinnedOwner:
    """Owner of one ``cudaHostAlloc`` block.  Arrays made by ``pinned_empty`` are plain
    ``numpy.ndarray``s whose ``.base`` chain ends here, so EVERY view derived from them --
    ``np.asarray(p)``, ``p.view(...)``, slices, ``torch.from_numpy(p)`` -- keeps the block
    page-locked and allocated (an ndarray subclass with ``__del__`` does not: numpy collapses
    ``.base`` past the subclass instance)."""
    __slots__ = ("ptr", "shape", "typestr")

    def __init__(self, ptr, shape, dtype):
        self.ptr, self.shape, self.typestr = ptr, tuple(shape), np.dtype(dtype).str

    @property
    def __array_interface__(self):
        return {"shape": self.shape, "typestr": self.typestr, "data": (self.ptr, False), "version": 3}

    def __del__(self):
        if self.ptr and _lib is not None:
            try:
                _lib.nphusl_cuda_host_free(ctypes.c_void_p(self.ptr))
            except Exception:
                pass
            self.ptr = 0


def pinned_empty(shape, dtype):
    """Page-locked host array (a plain ``numpy.ndarray``): host-pointer calls then DMA straight
    from / into it instead of staging through the library's own pinned slots.  The block is freed
    when the last array or view referring to it is gone."""
    dtype = np.dtype(dtype)
    shape = tuple(int(x) for x in shape) if hasattr(shape, "__iter__") else (int(shape),)
    n = _prod(shape) * dtype.itemsize
    p = ctypes.c_void_p()
    _check(load().nphusl_cuda_host_alloc(ctypes.byref(p), n), "host_alloc")
    return np.asarray(_PinnedOwner(p.value, shape, dtype))


# ---------------------------------------------------------------------------
# page-locked results for host-array calls
# ---------------------------------------------------------------------------
# `to_husl(numpy_image)` returns a NEW float64 array: 50 MB for a 1080p frame.  As
# a fresh pageable allocation that costs more than the conversion itself --
# 12 000 first-touch page faults plus a bounce copy out of the pinned staging
# slot (numpy's own `hsl.copy()` of that array runs at 213 Mpix/s on the B200
# hosts).  Results are therefore carved out of a small recycling pool of
# page-locked blocks: the D2H copy lands in the returned array directly, and when
# the caller drops the array the block goes back to the pool.  The returned object
# is a plain numpy.ndarray whose `.base` keeps the block alive; passing it back
# into `to_rgb` is a direct DMA as well.  `set_pinned_results(False)` restores
# plain `np.empty` results.

class _PinnedBlock:
    """Owner of one page-locked block, exposed through __array_interface__."""
    __slots__ = ("ptr", "cap", "shape", "typestr")

    def __init__(self, ptr, cap, shape, dtype):
        self.ptr, self.cap, self.shape, self.typestr = ptr, cap, tuple(shape), np.dtype(dtype).str

    @property
    def __array_interface__(self):
        return {"shape": self.shape, "typestr": self.typestr, "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            _pinned_pool.release(self.ptr, self.cap)
        except Exception:
            pass


class _PinnedPool:
    GRAIN = 2 << 20                      # size classes of 2 MiB

    def __init__(self):
        self.enabled = os.environ.get("NPHUSL_PINNED_RESULTS", "1") != "0"
        self.min_bytes = 1 << 20         # smaller results: np.empty
        self.max_cached = int(os.environ.get("NPHUSL_PINNED_CACHE_MB", "1024")) << 20
        self.max_outstanding = int(os.environ.get("NPHUSL_PINNED_MAX_MB", "8192")) << 20
        self.free = {}                   # cap -> [ptr, ...]
        self.cached = 0
        self.outstanding = 0
        self.lock = threading.RLock()    # re-entrant: see _DevicePool.lock
        self.hits = self.misses = 0

    def empty(self, shape, dtype):
        """-> ndarray backed by a pooled pinned block, or None (disabled / too small / over budget)"""
        dtype = np.dtype(dtype)
        n = _prod(shape) * dtype.itemsize
        if not self.enabled or n < self.min_bytes:
            return None
        cap = (n + self.GRAIN - 1) // self.GRAIN * self.GRAIN
        ptr = None
        with self.lock:
            if self.outstanding + cap > self.max_outstanding:
                return None
            lst = self.free.get(cap)
            if lst:
                ptr = lst.pop()
                self.cached -= cap
                self.hits += 1
            self.outstanding += cap
        if ptr is None:
            p = ctypes.c_void_p()
            if load().nphusl_cuda_host_alloc(ctypes.byref(p), cap) != 0 or not p.value:
                with self.lock:
                    self.outstanding -= cap
                return None              # the host refuses more pinned memory: pageable result
            ptr = p.value
            with self.lock:
                self.misses += 1
        return np.asarray(_PinnedBlock(ptr, cap, shape, dtype))

    def release(self, ptr, cap):
        keep = False
        with self.lock:
            self.outstanding -= cap
            if self.enabled and self.cached + cap <= self.max_cached:
                self.free.setdefault(cap, []).append(ptr)
                self.cached += cap
                keep = True
        if not keep and _lib is not None:
            _lib.nphusl_cuda_host_free(ctypes.c_void_p(ptr))

    def trim(self):
        with self.lock:
            blocks = [p for lst in self.free.values() for p in lst]
            self.free.clear()
            self.cached = 0
        for p in blocks:
            if _lib is not None:
                _lib.nphusl_cuda_host_free(ctypes.c_void_p(p))


_pinned_pool = _PinnedPool()


def set_pinned_results(enabled=True, max_cached_mb=None, max_outstanding_mb=None):
    """Host-array calls return page-locked (pooled) arrays when enabled (default)."""
    _pinned_pool.enabled = bool(enabled)
    if max_cached_mb is not None:
        _pinned_pool.max_cached = int(max_cached_mb) << 20
    if max_outstanding_mb is not None:
        _pinned_pool.max_outstanding = int(max_outstanding_mb) << 20
    if not enabled:
        _pinned_pool.trim()


def pinned_pool_stats():
    p = _pinned_pool
    return {"enabled": p.enabled, "cached_bytes": p.cached, "outstanding_bytes": p.outstanding,
            "hits": p.hits, "misses": p.misses}


# ---------------------------------------------------------------------------
# buffer description
# ---------------------------------------------------------------------------

class _Buf:
    ptr = 0
    shape = ()
    dtype = None
    loc = HOST
    device = None
    keep = None
    channels = 3
    strides = None          # byte strides of a NON-contiguous device array (else None)


def _is_device(obj):
    """Does `obj` live in GPU memory?  Cheap: torch builds its __cuda_array_interface__ dict on every
    attribute access (3.4 us), so plain hasattr() on the hot path cost more than the kernel launch."""
    if isinstance(obj, DeviceArray):
        return True
    if isinstance(obj, np.ndarray):
        return False
    is_cuda = getattr(obj, "is_cuda", None)          # torch tensors answer without building anything
    if isinstance(is_cuda, bool):
        return is_cuda
    return hasattr(obj, "__cuda_array_interface__")


def _device_of_ptr(ptr):
    loc, dev, pin = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    load().nphusl_cuda_pointer_info(ctypes.c_void_p(ptr), ctypes.byref(loc), ctypes.byref(dev), ctypes.byref(pin))
    return dev.value if loc.value == DEVICE and dev.value >= 0 else None


_TORCH_DT = {"torch.uint8": np.dtype(np.uint8), "torch.float32": np.dtype(np.float32), "torch.float64": np.dtype(np.float64)}


def _cai_fields(obj):
    """-> (shape, dtype, byte strides or None, data pointer).  torch tensors are read through
    data_ptr() / shape / stride() / dtype (0.2 us each): building their __cuda_array_interface__
    dict costs 3.4 us per access, a third of a whole call on a small frame."""
    dp = getattr(obj, "data_ptr", None)
    if dp is not None and isinstance(getattr(obj, "is_cuda", None), bool):
        dtype = _TORCH_DT.get(str(obj.dtype))
        if dtype is not None and not getattr(obj, "requires_grad", False):
            shape = tuple(obj.shape)
            strides = None if obj.is_contiguous() else tuple(s * dtype.itemsize for s in obj.stride())
            return shape, dtype, strides, dp()
    cai = obj.__cuda_array_interface__
    ptr = cai["data"][0]
    return tuple(cai["shape"]), np.dtype(cai["typestr"]), cai.get("strides"), int(ptr) if ptr else 0


def _describe_device(obj, what, allow_gray=False, strided_ok=False):
    shape, dtype, strides, ptr = _cai_fields(obj)
    channels = 3
    if allow_gray and strides is not None and shape and shape[-1] == 3 and strides[-1] == 0:
        # transform.GrayDevice: one stored value per pixel, broadcast over R, G, B
        channels = 1
        strides = tuple(strides[:-1]) + (dtype.itemsize,)
        full_shape, shape = shape, shape[:-1] + (1,)
    if strides is not None:
        expect, acc = [], dtype.itemsize
        for s in reversed(shape):
            expect.append(acc)
            acc *= s
        if any(s != e for s, e, n in zip(strides, reversed(expect), shape) if n > 1):
            if not strided_ok or channels == 1:
                raise ValueError("{}: device arrays must be C-contiguous here (got strides {})".format(what, strides))
            keep_strides = tuple(int(x) for x in strides)
        else:
            keep_strides = None
    else:
        keep_strides = None
    b = _Buf()
    b.strides = keep_strides
    b.ptr = ptr
    b.shape, b.dtype, b.loc, b.keep, b.channels = shape, dtype, DEVICE, obj, channels
    if channels == 1:
        b.shape = full_shape
    dev = getattr(obj, "device", None)
    if isinstance(dev, int):
        b.device = dev
    elif dev is not None and hasattr(dev, "index") and dev.index is not None:
        b.device = int(dev.index)
    else:
        b.device = _device_of_ptr(b.ptr) if b.ptr else None
    return b


def _describe_rgb_in(rgb):
    """RGB-side input: uint8 stays uint8, float32/float64 stay, every other
    integer type is cast to uint8 (what rgb_int_input does, transform.py:16-20);
    a zero-stride last axis (transform.from_grayscale's broadcast view) is sent
    as ONE channel and broadcast in registers."""
    if _is_device(rgb):
        b = _describe_device(rgb, "rgb", allow_gray=True, strided_ok=True)
        if b.dtype not in _DT:
            raise ValueError("rgb device array dtype {} unsupported (uint8/float32/float64)".format(b.dtype))
        if getattr(rgb, "rgba", False):          # transform.RgbaDevice: 4 stored channels, premultiplied on load
            if b.strides is not None:
                raise ValueError("rgb: RGBA device arrays must be C-contiguous")
            b.channels = 4
            b.shape = tuple(b.shape[:-1]) + (3,)
        return b
    a = np.asarray(rgb)
    b = _Buf()
    b.loc, b.device, b.channels = HOST, None, 3
    b.shape = a.shape
    if a.ndim >= 1 and a.shape[-1] == 3 and a.strides[-1] == 0 and a.size:
        a = a[..., 0]
        b.channels = 1
    if a.dtype not in _DT:
        if np.issubdtype(a.dtype, np.integer) or a.dtype == np.bool_:
            a = a.astype(np.uint8)
        else:
            a = a.astype(np.float64)
    a = np.ascontiguousarray(a)
    b.ptr, b.dtype, b.keep = a.ctypes.data, a.dtype, a
    return b


def _describe_out(out, shape, dtype, allowed, what, strided_ok=False):
    """Validate a caller-provided ``out`` (host ndarray or device array)."""
    if _is_device(out):
        b = _describe_device(out, what, strided_ok=strided_ok)
        if b.strides is not None and tuple(b.shape) != tuple(shape):
            raise ValueError("{}: a strided device array must have exactly the result's shape {}".format(what, tuple(shape)))
    else:
        if not isinstance(out, np.ndarray):
            raise TypeError("{} must be a numpy array or a device array".format(what))
        if not out.flags.c_contiguous or not out.flags.writeable:
            raise ValueError("{} must be C-contiguous and writeable".format(what))
        b = _Buf()
        b.ptr, b.shape, b.dtype, b.loc, b.device, b.keep, b.channels = out.ctypes.data, out.shape, out.dtype, HOST, None, out, 3
    if tuple(b.shape) != tuple(shape) and [n for n in b.shape if n != 1] != [n for n in shape if n != 1]:
        # only unit axes may differ (the squeezed single pixel, a batch of one frame)
        raise ValueError("{} has shape {}, expected {}".format(what, tuple(b.shape), tuple(shape)))
    if b.dtype not in allowed:
        raise ValueError("{} dtype {} unsupported here (allowed: {})".format(what, b.dtype, ", ".join(str(np.dtype(d)) for d in allowed)))
    return b


def _pick_device(device, *bufs):
    if device is not None:
        return int(device)
    for b in bufs:
        if b is not None and b.loc == DEVICE and b.device is not None:
            return b.device
    return _default_device


def _touch(stream, *objs):
    """Record that the library uses these arrays on ``stream`` (block reuse is ordered by it,
    see ``_DevicePool``).  Only our own ``DeviceArray``s -- bare, or wrapped by
    ``transform.GrayDevice`` / ``RgbaDevice`` -- are tracked; torch owns its tensors."""
    key = None
    for o in objs:
        if type(o) is not DeviceArray:
            o = getattr(o, "base", None)
            if type(o) is not DeviceArray:
                continue
        while o._streams is None:
            o = o._base
        if key is None:
            key = _stream_key(stream)
        o._streams.add(key)


def _make_out(src, out, shape, dtype, allowed, device, what, strided_ok=False, stream=None):
    """-> (_Buf, object to return)"""
    if out is not None:
        b = _describe_out(out, shape, dtype, allowed, what, strided_ok)
        _touch(stream, out, src.keep)
        return b, out
    _touch(stream, src.keep)
    if src.loc == DEVICE:
        d = DeviceArray(shape, dtype, device, stream)
        b = _Buf()
        b.ptr, b.shape, b.dtype, b.loc, b.device, b.keep, b.channels = d.ptr, d.shape, d.dtype, DEVICE, d.device, d, 3
        return b, d
    a = _pinned_pool.empty(shape, dtype)
    if a is None:
        a = np.empty(shape, dtype)
    b = _Buf()
    b.ptr, b.shape, b.dtype, b.loc, b.device, b.keep, b.channels = a.ctypes.data, a.shape, a.dtype, HOST, None, a, 3
    return b, a


_FLOATS = (np.dtype(np.float32), np.dtype(np.float64))
_ANY = (np.dtype(np.uint8),) + _FLOATS


# ---------------------------------------------------------------------------
# strided device images (cropped views, channel slices, pitched frames)
# ---------------------------------------------------------------------------

class View(ctypes.Structure):
    """``nphusl_view`` of include/nphusl_cuda.h"""
    _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int),
                ("frames", ctypes.c_longlong), ("rows", ctypes.c_longlong), ("cols", ctypes.c_longlong),
                ("frame_stride", ctypes.c_longlong), ("row_stride", ctypes.c_longlong),
                ("col_stride", ctypes.c_longlong), ("ch_stride", ctypes.c_longlong)]


def _c_strides(shape, itemsize):
    out, acc = [], itemsize
    for n in reversed(shape):
        out.append(acc)
        acc *= n
    return tuple(reversed(out))


def _views(bufs, pixel_ndim):
    """Collapse the pixel axes of several same-shaped buffers JOINTLY into
    (frames, rows, cols): neighbouring axes merge when they are mergeable in
    every buffer (stride[i] == stride[i+1] * shape[i+1]).  -> list of View.
    Raises ValueError when more than three axes remain (make the array contiguous)."""
    shape = tuple(bufs[0].shape[:pixel_ndim])
    strides = []
    for b in bufs:
        st = b.strides if b.strides is not None else _c_strides(b.shape, b.dtype.itemsize)
        strides.append(list(st))
    axes = [i for i, n in enumerate(shape) if n != 1]
    dims = [[shape[i]] + [st[i] for st in strides] for i in axes]          # [n, stride_buf0, stride_buf1, ...]
    merged = []
    for d in dims:
        if merged and all(merged[-1][k] == d[k] * d[0] for k in range(1, len(d))):
            merged[-1] = [merged[-1][0] * d[0]] + d[1:]
        else:
            merged.append(list(d))
    if len(merged) > 3:
        raise ValueError("strided device array with {} non-mergeable pixel axes: at most 3 are supported "
                         "(call .contiguous() first)".format(len(merged)))
    while len(merged) < 3:
        merged.insert(0, [1] + [0] * len(bufs))
    views = []
    for k, b in enumerate(bufs):
        st = strides[k]
        es = st[pixel_ndim] if len(b.shape) > pixel_ndim else 0
        views.append(View(b.ptr, _DT[b.dtype], merged[0][0], merged[1][0], merged[2][0],
                          merged[0][k + 1], merged[1][k + 1], merged[2][k + 1], es))
    return views


# ---------------------------------------------------------------------------
# the three backend callables
# ---------------------------------------------------------------------------

def _call(lib, blocking, fn, *args):
    """Invoke a C ABI conversion.  ``blocking=False``: for this one call, on this thread, pinned-host
    buffers are only enqueued (nphusl_cuda_set_host_async); pageable memory still completes before
    the call returns.  The default (blocking) path is the bare foreign call."""
    if blocking:
        return fn(*args)
    lib.nphusl_cuda_set_host_async(1)
    try:
        return fn(*args)
    finally:
        lib.nphusl_cuda_set_host_async(0)



def _strided_call(fn_name, src, dst, pixel_ndim, extra, dev, stream):
    """Either side is a non-contiguous device array: both go to the strided entry points."""
    if src.loc != DEVICE or dst.loc != DEVICE:
        raise ValueError("strided device arrays convert device -> device only (pass a device `out` or none)")
    vin, vout = _views([src, dst], pixel_ndim)
    args = [ctypes.byref(vin), ctypes.byref(vout)] + list(extra) + [dev, _stream_ptr(stream)]
    _check(getattr(load(), fn_name)(*args), fn_name[len("nphusl_cuda_"):])


def _rgb_to_husl(rgb, out=None, dtype=np.float64, mode="exact", stream=None, device=None, blocking=True):
    """RGB -> HUSL on the GPU (kernel (a) husl_from_rgb).  Replaces
    ``_simd_opt._rgb_to_husl`` (_simd_opt.pyx:20-33) / ``_cython_opt._rgb_to_husl``
    (_cython_opt.pyx:22-50)."""
    lib = load()
    src = _describe_rgb_in(rgb)
    if not src.shape or src.shape[-1] != 3:
        raise ValueError("expected (..., 3) RGB, got shape {}".format(src.shape))
    dev = _pick_device(device, src)
    dst, ret = _make_out(src, out, src.shape, dtype, _FLOATS, dev, "out", True, stream)
    dev = _pick_device(device, src, dst)                  # host input, `out` on another GPU: go where `out` lives
    n = _prod(src.shape[:-1])
    if src.strides is not None or dst.strides is not None:
        if mode != "exact":
            raise ValueError("LUT modes take C-contiguous arrays")
        _strided_call("nphusl_cuda_rgb_to_husl_strided", src, dst, len(src.shape) - 1, [0], dev, stream)
        return ret
    rc = _call(lib, blocking, lib.nphusl_cuda_rgb_to_husl, src.ptr, _DT[src.dtype], src.loc, src.channels,
               dst.ptr, _DT[dst.dtype], dst.loc, n, MODES[mode], dev, _stream_ptr(stream))
    if rc:
        _check(rc, "rgb_to_husl")
    return ret


def _rgb_to_hue(rgb, out=None, dtype=np.float64, stream=None, device=None, blocking=True):
    """RGB -> hue on the GPU (kernel (c) hue_from_rgb).  Replaces
    ``_cython_opt._rgb_to_hue`` (_cython_opt.pyx:96-111)."""
    lib = load()
    src = _describe_rgb_in(rgb)
    if not src.shape or src.shape[-1] != 3:
        raise ValueError("expected (..., 3) RGB, got shape {}".format(src.shape))
    dev = _pick_device(device, src)
    shape = src.shape[:-1]
    dst, ret = _make_out(src, out, shape, dtype, _FLOATS, dev, "out", True, stream)
    dev = _pick_device(device, src, dst)
    n = _prod(shape)
    if src.strides is not None or dst.strides is not None:
        _strided_call("nphusl_cuda_rgb_to_husl_strided", src, dst, len(shape), [1], dev, stream)
        return ret
    rc = _call(lib, blocking, lib.nphusl_cuda_rgb_to_hue, src.ptr, _DT[src.dtype], src.loc, src.channels,
               dst.ptr, _DT[dst.dtype], dst.loc, n, dev, _stream_ptr(stream))
    if rc:
        _check(rc, "rgb_to_hue")
    return ret


def _as_edit(edit):
    """``Edit`` | ``Bands`` | dict of ``Edit``'s keywords -> the ctypes structure"""
    if isinstance(edit, (Edit, Bands)):
        return edit
    if isinstance(edit, dict):
        return Edit(**edit)
    raise TypeError("edit must be an nphusl._cuda.Edit / Bands or a dict of Edit's keywords, got {!r}".format(type(edit)))


def _need_edit(edit, who):
    """entry points that take the select + affine ``Edit`` only"""
    edit = _as_edit(edit)
    if not isinstance(edit, Edit):
        raise TypeError("{} takes an Edit; a Bands map goes to to_rgb(hsl, edit=...) or edit_rgb(rgb, edit=...)".format(who))
    return edit


def _husl_to_rgb(hsl, out=None, dtype=np.uint8, stream=None, device=None, blocking=True, edit=None):
    """HUSL -> RGB on the GPU (kernel (b) rgb_from_husl).  Replaces
    ``_cython_opt._husl_to_rgb`` (_cython_opt.pyx:177-192); with the default
    ``dtype=uint8`` the round/abs/cast tail of ``to_rgb`` is fused in, float
    dtypes return RGB in [0, 1] like the reference callable.

    ``edit`` (an ``Edit`` or a dict of its keywords): the callers' HUSL-space edit of
    README.md:77-98 -- ``hsl[..., 2][mask] *= 0.5`` between ``to_husl`` and ``to_rgb`` --
    applied to every pixel as it is loaded, so the pipeline needs no pass of its own over
    the HUSL array: ``to_rgb(hsl, edit=dict(hue=(250, 290), invert=True, l=(0.5, 0)))``
    equals ``to_rgb(edit_husl(hsl, ...))``; ``hsl`` itself is left untouched.  A ``Bands`` instead
    (examples.py:56-103) paints bands of one channel with two alternating hues on load."""
    lib = load()
    if _is_device(hsl):
        src = _describe_device(hsl, "hsl", strided_ok=True)
        if src.dtype not in _FLOATS:
            raise ValueError("hsl device array must be float32/float64, got {}".format(src.dtype))
    else:
        a = np.asarray(hsl)
        if a.dtype not in _FLOATS:
            a = a.astype(np.float64)           # float_input, transform.py:57-61
        a = np.ascontiguousarray(a)
        src = _Buf()
        src.ptr, src.shape, src.dtype, src.loc, src.device, src.keep, src.channels = a.ctypes.data, a.shape, a.dtype, HOST, None, a, 3
    if not src.shape or src.shape[-1] != 3:
        raise ValueError("expected (..., 3) HUSL, got shape {}".format(src.shape))
    dev = _pick_device(device, src)
    dst, ret = _make_out(src, out, src.shape, dtype, _ANY, dev, "out", True, stream)
    dev = _pick_device(device, src, dst)
    n = _prod(src.shape[:-1])
    if src.strides is not None or dst.strides is not None:
        if edit is not None:
            raise ValueError("to_rgb(edit=...) takes C-contiguous arrays")
        _strided_call("nphusl_cuda_husl_to_rgb_strided", src, dst, len(src.shape) - 1, [], dev, stream)
        return ret
    if edit is not None:
        edit = _as_edit(edit)
        fn = lib.nphusl_cuda_husl_bands_to_rgb if isinstance(edit, Bands) else lib.nphusl_cuda_husl_edit_to_rgb
        rc = _call(lib, blocking, fn, src.ptr, _DT[src.dtype], src.loc, ctypes.byref(edit),
                   dst.ptr, _DT[dst.dtype], dst.loc, n, dev, _stream_ptr(stream))
    else:
        rc = _call(lib, blocking, lib.nphusl_cuda_husl_to_rgb, src.ptr, _DT[src.dtype], src.loc, dst.ptr, _DT[dst.dtype], dst.loc,
                   n, dev, _stream_ptr(stream))
    if rc:
        _check(rc, "husl_to_rgb")
    return ret


# backends that can write straight into a caller's `out` advertise it
# (transform.in_chunks then skips its copy)
for _f in (_rgb_to_husl, _rgb_to_hue, _husl_to_rgb):
    _f.accepts_out = True


# ---------------------------------------------------------------------------
# callers' side of the path on the GPU: planar HUSL, fused edit
# ---------------------------------------------------------------------------

def _to_device(arr, device, stream):
    """host ndarray -> DeviceArray (device arrays pass through)"""
    if _is_device(arr):
        return arr
    return DeviceArray.from_host(np.ascontiguousarray(arr), device, stream)


def rgb_to_husl_planar(rgb, out=None, dtype=np.float32, stream=None, device=None, channels_first=False):
    """uint8 RGB -> planar HUSL ``(3, ...)``: ``out[0]`` is the hue plane,
    ``out[1]`` saturation, ``out[2]`` lightness (what callers slice out of the
    interleaved array anyway, README.md:77-98).  ``rgb`` is ``(..., 3)``
    interleaved, or ``(3, ...)`` planes with ``channels_first=True`` (the torch /
    vision image layout).  Device arrays stay on the device; a host array is
    uploaded, converted and the planes downloaded."""
    lib = load()
    host_in = not _is_device(rgb)
    dev = device
    if host_in:
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        dev = _default_device if dev is None else dev
    src = _describe_device(_to_device(rgb, dev, stream), "rgb")
    ok = src.shape and (src.shape[0] == 3 if channels_first else src.shape[-1] == 3)
    if src.dtype != np.uint8 or not ok:
        raise ValueError("planar conversion takes uint8 RGB {}, got {} {}".format(
            "(3, ...)" if channels_first else "(..., 3)", src.dtype, src.shape))
    dev = _pick_device(dev, src)
    lead = tuple(src.shape[1:]) if channels_first else tuple(src.shape[:-1])
    shape = (3,) + lead
    n = _prod(lead)
    if out is not None and _is_device(out):
        dst = _describe_out(out, shape, dtype, _FLOATS, "out")
        ret = out
    else:
        d = DeviceArray(shape, np.dtype(dtype if out is None else out.dtype), dev, stream)
        dst = _describe_device(d, "out")
        ret = d
    _touch(stream, src.keep, ret)
    plane = n * dst.dtype.itemsize
    if channels_first:
        _check(lib.nphusl_cuda_rgb_planes_to_husl_planes(src.ptr, src.ptr + n, src.ptr + 2 * n,
                                                         dst.ptr, dst.ptr + plane, dst.ptr + 2 * plane, _DT[dst.dtype],
                                                         n, dev, _stream_ptr(stream)), "rgb_planes_to_husl_planes")
    else:
        _check(lib.nphusl_cuda_rgb_to_husl_planar(src.ptr, dst.ptr, dst.ptr + plane, dst.ptr + 2 * plane, _DT[dst.dtype],
                                                  n, dev, _stream_ptr(stream)), "rgb_to_husl_planar")
    if host_in or (out is not None and not _is_device(out)):
        synchronize(dev, stream)
        return ret.copy_to_host(out)
    return ret


def husl_planar_to_rgb(hsl, out=None, stream=None, device=None, channels_first=False):
    """planar HUSL ``(3, ...)`` float32/float64 -> uint8 RGB, ``(..., 3)``
    interleaved or ``(3, ...)`` planes with ``channels_first=True``."""
    lib = load()
    host_in = not _is_device(hsl)
    dev = device
    if host_in:
        hsl = np.ascontiguousarray(hsl)
        if hsl.dtype not in _FLOATS:
            hsl = hsl.astype(np.float64)
        dev = _default_device if dev is None else dev
    src = _describe_device(_to_device(hsl, dev, stream), "hsl")
    if src.dtype not in _FLOATS or not src.shape or src.shape[0] != 3:
        raise ValueError("expected planar HUSL (3, ...), float32/float64; got {} {}".format(src.dtype, src.shape))
    dev = _pick_device(dev, src)
    shape = ((3,) + tuple(src.shape[1:])) if channels_first else (tuple(src.shape[1:]) + (3,))
    n = _prod(src.shape[1:])
    if out is not None and _is_device(out):
        dst = _describe_out(out, shape, np.uint8, (np.dtype(np.uint8),), "out")
        ret = out
    else:
        d = DeviceArray(shape, np.uint8, dev, stream)
        dst = _describe_device(d, "out")
        ret = d
    _touch(stream, src.keep, ret)
    plane = n * src.dtype.itemsize
    if channels_first:
        _check(lib.nphusl_cuda_husl_planes_to_rgb_planes(src.ptr, src.ptr + plane, src.ptr + 2 * plane, _DT[src.dtype],
                                                         dst.ptr, dst.ptr + n, dst.ptr + 2 * n,
                                                         n, dev, _stream_ptr(stream)), "husl_planes_to_rgb_planes")
    else:
        _check(lib.nphusl_cuda_husl_planar_to_rgb(src.ptr, src.ptr + plane, src.ptr + 2 * plane, _DT[src.dtype],
                                                  dst.ptr, n, dev, _stream_ptr(stream)), "husl_planar_to_rgb")
    if host_in or (out is not None and not _is_device(out)):
        synchronize(dev, stream)
        return ret.copy_to_host(out)
    return ret


class Edit(ctypes.Structure):
    """``nphusl_edit`` of include/nphusl_cuda.h: a selection box in HUSL space
    and an affine edit applied to the selected pixels."""
    _fields_ = [("h_lo", ctypes.c_float), ("h_hi", ctypes.c_float), ("s_lo", ctypes.c_float), ("s_hi", ctypes.c_float),
                ("l_lo", ctypes.c_float), ("l_hi", ctypes.c_float), ("invert", ctypes.c_int),
                ("h_mul", ctypes.c_float), ("h_add", ctypes.c_float), ("s_mul", ctypes.c_float), ("s_add", ctypes.c_float),
                ("l_mul", ctypes.c_float), ("l_add", ctypes.c_float)]

    # unspecified ranges are unbounded (S can exceed 100 by an ulp on the gamut boundary)
    ALL = (-float("inf"), float("inf"))

    def __init__(self, hue=ALL, saturation=ALL, lightness=ALL, invert=False,
                 h=(1.0, 0.0), s=(1.0, 0.0), l=(1.0, 0.0)):     # noqa: E741
        super().__init__(hue[0], hue[1], saturation[0], saturation[1], lightness[0], lightness[1], int(bool(invert)),
                         h[0], h[1], s[0], s[1], l[0], l[1])


class Bands(ctypes.Structure):
    """``nphusl_bands`` of include/nphusl_cuda.h: the banded map of the reference's callers
    (examples.py:56-103).  The range ``[0, total)`` of channel ``src`` (``"h"``, ``"s"`` or ``"l"``) is cut
    like ``chunk(total, width)``; a pixel strictly inside an even band gets hue ``h_even``, inside an odd
    band ``h_odd``; with ``stripe=(half_width, saturation)`` pixels within ``half_width`` of a band centre
    also get that saturation.  ``hue_watermelon``: ``Bands("h", 45, 360, 130, 360)``; one frame of
    ``melonize``: ``Bands("l", chunksize, 100, 130, 360, stripe=(2, 100))``.  Pass it as ``edit=`` to
    ``to_rgb`` (applied on load) or ``edit_rgb`` / ``nphusl.edit`` (fused uint8 -> uint8)."""
    _fields_ = [("src", ctypes.c_int), ("width", ctypes.c_float), ("total", ctypes.c_float),
                ("h_even", ctypes.c_float), ("h_odd", ctypes.c_float), ("s_half", ctypes.c_float), ("s_value", ctypes.c_float)]
    SRC = {"h": 0, "hue": 0, "s": 1, "saturation": 1, "l": 2, "lightness": 2, 0: 0, 1: 1, 2: 2}

    def __init__(self, src="h", width=45.0, total=None, h_even=130.0, h_odd=360.0, stripe=None):
        if src not in self.SRC:
            raise ValueError("Bands: src must be 'h', 's' or 'l', got {!r}".format(src))
        code = self.SRC[src]
        if total is None:
            total = 360.0 if code == 0 else 100.0
        if not (width > 0 and total > 0):
            raise ValueError("Bands: width and total must be positive")
        half, value = stripe if stripe is not None else (0.0, 0.0)
        super().__init__(code, width, total, h_even, h_odd, half, value)


def edit_rgb(rgb, edit=None, out=None, stream=None, device=None, blocking=True, channels_first=False, **edit_args):
    """Fused ``to_rgb(edit(to_husl(rgb)))`` on uint8 pixels in one kernel pass.

    ``channels_first=True``: ``rgb`` (a device array) is ``(3, ...)`` planes, the
    layout nvJPEG / torchvision decoders produce; so is the result.

    ``edit`` is an ``Edit`` (or pass its fields as keywords), or a ``Bands`` map (interleaved pixels only):
    ``edit_rgb(img, hue=(250, 290), invert=True, l=(0.5, 0))`` darkens every
    non-bluish pixel (README.md:77-81); ``edit_rgb(img, lightness=(0, 62),
    l=(0, 0))`` blacks out dim pixels (README.md:94-98).  Host or device arrays;
    ``out`` may be the input itself."""
    lib = load()
    if edit is None:
        edit = Edit(**edit_args)
    elif edit_args:
        raise TypeError("pass either an Edit or its fields as keywords, not both")
    edit = _as_edit(edit)
    bands = isinstance(edit, Bands)
    if bands and channels_first:
        raise ValueError("a Bands map takes interleaved (..., 3) pixels")
    if channels_first:
        if not _is_device(rgb):
            raise ValueError("channels_first edit takes device arrays")
        src = _describe_device(rgb, "rgb")
        if src.dtype != np.uint8 or not src.shape or src.shape[0] != 3:
            raise ValueError("channels_first edit takes uint8 (3, ...) planes, got {} {}".format(src.dtype, src.shape))
        dev = _pick_device(device, src)
        dst, ret = _make_out(src, out, src.shape, np.uint8, (np.dtype(np.uint8),), dev, "out", False, stream)
        if dst.loc != DEVICE:
            raise ValueError("channels_first edit writes device arrays")
        n = _prod(src.shape[1:])
        _check(lib.nphusl_cuda_rgb_planes_edit(src.ptr, src.ptr + n, src.ptr + 2 * n, dst.ptr, dst.ptr + n, dst.ptr + 2 * n,
                                               n, ctypes.byref(edit), dev, _stream_ptr(stream)), "rgb_planes_edit")
        return ret
    src = _describe_rgb_in(rgb)
    if src.dtype != np.uint8 or src.channels != 3 or not src.shape or src.shape[-1] != 3:
        raise ValueError("edit_rgb takes uint8 RGB (..., 3), got {} {}".format(src.dtype, src.shape))
    dev = _pick_device(device, src)
    dst, ret = _make_out(src, out, src.shape, np.uint8, (np.dtype(np.uint8),), dev, "out", True, stream)
    dev = _pick_device(device, src, dst)
    n = _prod(src.shape[:-1])
    if src.strides is not None or dst.strides is not None:
        if bands:
            raise ValueError("a Bands map takes C-contiguous arrays")
        _strided_call("nphusl_cuda_rgb_edit_rgb_strided", src, dst, len(src.shape) - 1, [ctypes.byref(edit)], dev, stream)
        return ret
    fn = lib.nphusl_cuda_rgb_bands_rgb if bands else lib.nphusl_cuda_rgb_edit_rgb
    rc = _call(lib, blocking, fn, src.ptr, src.loc, dst.ptr, dst.loc, n, ctypes.byref(edit), dev, _stream_ptr(stream))
    if rc:
        _check(rc, "rgb_edit_rgb")
    return ret


def edit_husl(hsl, edit=None, out=None, stream=None, **edit_args):
    """Apply a select + affine edit to a HUSL DEVICE array (``(..., 3)`` float32 / float64) in one
    pass: what ``hue, lightness = hsl[..., 0], hsl[..., 2]; lightness[~bluish] *= 0.5`` does
    (README.md:77-81).  In place by default (``out=None`` returns ``hsl``); ``out`` may be another
    device array of the same shape and dtype.  ``to_rgb(hsl, edit=...)`` applies the same edit
    without touching the array at all."""
    if edit is None:
        edit = Edit(**edit_args)
    elif edit_args:
        raise TypeError("pass either an Edit or its fields as keywords, not both")
    if not _is_device(hsl):
        raise ValueError("edit_husl takes device arrays (edit a host array with numpy, or use to_rgb(hsl, edit=...))")
    src = _describe_device(hsl, "hsl")
    if src.dtype not in _FLOATS or not src.shape or src.shape[-1] != 3:
        raise ValueError("edit_husl takes float32/float64 (..., 3) HUSL, got {} {}".format(src.dtype, src.shape))
    dev = _pick_device(None, src)
    if out is None:
        dst, ret = src, hsl
    else:
        dst = _describe_out(out, src.shape, src.dtype, (src.dtype,), "out")
        if dst.loc != DEVICE:
            raise ValueError("out: edit_husl writes device arrays")
        ret = out
    _touch(stream, hsl, ret)
    _check(load().nphusl_cuda_husl_edit(src.ptr, dst.ptr, _DT[src.dtype], _prod(src.shape[:-1]), ctypes.byref(_need_edit(edit, "edit_husl")),
                                        dev, _stream_ptr(stream)), "husl_edit")
    return ret


# ---------------------------------------------------------------------------
# video-decoder surfaces (NV12) as the load stage
# ---------------------------------------------------------------------------

MATRICES = {"bt601": 0, "bt709": 1}


class Nv12(ctypes.Structure):
    """``nphusl_nv12`` of include/nphusl_cuda.h"""
    _fields_ = [("y", ctypes.c_void_p), ("uv", ctypes.c_void_p),
                ("y_pitch", ctypes.c_longlong), ("uv_pitch", ctypes.c_longlong),
                ("width", ctypes.c_int), ("height", ctypes.c_int),
                ("matrix", ctypes.c_int), ("full_range", ctypes.c_int)]


def _plane(obj, what):
    """-> (ptr, shape, row pitch in bytes, device, keep-alive) of a uint8 device plane whose rows are dense"""
    if not _is_device(obj):
        raise ValueError("{}: NV12 planes are device arrays (decoder surfaces); upload host frames with "
                         "DeviceArray.from_host".format(what))
    shape, dtype, strides, ptr = _cai_fields(obj)
    if dtype != np.uint8 or len(shape) not in (2, 3):
        raise ValueError("{}: expected a 2-D (or (rows, pairs, 2)) uint8 plane, got {} {}".format(what, dtype, shape))
    row = _prod(shape[1:])
    if strides is None:
        pitch = row
    else:
        if tuple(strides[1:]) != _c_strides(shape, 1)[1:]:
            raise ValueError("{}: rows must be dense (only the row pitch may be padded), got strides {}".format(what, strides))
        pitch = int(strides[0]) if shape[0] > 1 else row
    dev = getattr(obj, "device", None)
    if not isinstance(dev, int):
        dev = int(dev.index) if dev is not None and getattr(dev, "index", None) is not None else _device_of_ptr(ptr)
    return ptr, shape, pitch, dev, obj


def _nv12_surface(y, uv, matrix, full_range):
    """``y``: (H, W) uint8 device plane; ``uv``: (ceil(H/2), 2*ceil(W/2)) or (ceil(H/2), ceil(W/2), 2)
    interleaved chroma.  ``uv=None``: ``y`` is a whole (H*3/2, W) NVDEC surface, luma rows first."""
    yp, ys, ypitch, dev, keep = _plane(y, "y")
    if len(ys) != 2:
        raise ValueError("y: expected a 2-D luma plane, got shape {}".format(ys))
    if uv is None:
        if ys[0] % 3:
            raise ValueError("a whole NV12 surface has 3/2 * height rows, got {}".format(ys[0]))
        h, w = ys[0] * 2 // 3, ys[1]
        if h % 2 or w % 2:
            raise ValueError("a whole NV12 surface needs even width and height, got {}x{}".format(w, h))
        up, upitch, keep2 = yp + h * ypitch, ypitch, None
    else:
        h, w = ys
        up, us, upitch, udev, keep2 = _plane(uv, "uv")
        if us[0] != (h + 1) // 2 or _prod(us[1:]) != 2 * ((w + 1) // 2) or udev != dev:
            raise ValueError("uv: expected {} rows of {} bytes on device {}, got shape {} on device {}".format(
                (h + 1) // 2, 2 * ((w + 1) // 2), dev, us, udev))
    if matrix not in MATRICES:
        raise ValueError("matrix must be one of {}".format(sorted(MATRICES)))
    return Nv12(yp, up, ypitch, upitch, w, h, MATRICES[matrix], int(bool(full_range))), (h, w), dev, (keep, keep2)


def _nv12_out(out, shape, dtype, allowed, dev, stream, keep):
    _touch(stream, *keep)
    if out is None:
        d = DeviceArray(shape, dtype, dev, stream)
        return d.ptr, d.dtype, d
    b = _describe_out(out, shape, dtype, allowed, "out")
    if b.loc != DEVICE:
        raise ValueError("out: NV12 conversions write device arrays")
    _touch(stream, out)
    return b.ptr, b.dtype, out


def nv12_to_rgb(y, uv=None, out=None, matrix="bt709", full_range=False, stream=None):
    """NV12 decoder surface -> uint8 RGB ``(H, W, 3)`` on the device (integer BT.601 / BT.709
    conversion documented in include/nphusl_cuda.h)."""
    src, (h, w), dev, keep = _nv12_surface(y, uv, matrix, full_range)
    ptr, _, ret = _nv12_out(out, (h, w, 3), np.uint8, (np.dtype(np.uint8),), dev, stream, keep)
    _check(load().nphusl_cuda_nv12_to_rgb(ctypes.byref(src), ptr, dev, _stream_ptr(stream)), "nv12_to_rgb")
    return ret


def nv12_to_husl(y, uv=None, out=None, dtype=np.float32, matrix="bt709", full_range=False, stream=None, hue_only=False):
    """NV12 decoder surface -> HUSL ``(H, W, 3)`` (or hue ``(H, W)``) on the device in ONE pass:
    the colour conversion is the load stage of the to_husl kernel, the result equals
    ``to_husl(nv12_to_rgb(...))`` bit for bit."""
    src, (h, w), dev, keep = _nv12_surface(y, uv, matrix, full_range)
    shape = (h, w) if hue_only else (h, w, 3)
    ptr, dt, ret = _nv12_out(out, shape, dtype, _FLOATS, dev, stream, keep)
    _check(load().nphusl_cuda_nv12_to_husl(ctypes.byref(src), ptr, _DT[dt], int(bool(hue_only)), dev, _stream_ptr(stream)),
           "nv12_to_husl")
    return ret


def nv12_to_hue(y, uv=None, **kw):
    return nv12_to_husl(y, uv, hue_only=True, **kw)


def nv12_edit_rgb(y, uv=None, edit=None, out=None, matrix="bt709", full_range=False, stream=None, **edit_args):
    """Fused NV12 -> HUSL -> select + edit -> uint8 RGB ``(H, W, 3)`` (see ``edit_rgb``)."""
    if edit is None:
        edit = Edit(**edit_args)
    elif edit_args:
        raise TypeError("pass either an Edit or its fields as keywords, not both")
    edit = _need_edit(edit, "nv12_edit_rgb")
    src, (h, w), dev, keep = _nv12_surface(y, uv, matrix, full_range)
    ptr, _, ret = _nv12_out(out, (h, w, 3), np.uint8, (np.dtype(np.uint8),), dev, stream, keep)
    _check(load().nphusl_cuda_nv12_edit_rgb(ctypes.byref(src), ctypes.byref(edit), ptr, dev, _stream_ptr(stream)), "nv12_edit_rgb")
    return ret


def _nv12_dest(h, w, out_y, out_uv, matrix, full_range, dev, stream=None):
    """-> (Nv12 descriptor of the destination, (y, uv) arrays to return)"""
    if h % 2 or w % 2:
        raise ValueError("an NV12 destination needs even width and height, got {}x{}".format(w, h))
    if (out_y is None) != (out_uv is None):
        raise ValueError("pass both out_y and out_uv, or neither")
    if out_y is None:
        out_y, out_uv = DeviceArray((h, w), np.uint8, dev, stream), DeviceArray((h // 2, w), np.uint8, dev, stream)
    _touch(stream, out_y, out_uv)
    yp, ys, ypitch, ydev, _ = _plane(out_y, "out_y")
    up, us, upitch, udev, _ = _plane(out_uv, "out_uv")
    if tuple(ys) != (h, w) or us[0] != h // 2 or _prod(us[1:]) != w or ydev != dev or udev != dev:
        raise ValueError("out_y / out_uv: expected ({}, {}) and ({}, {}) uint8 planes on device {}".format(h, w, h // 2, w, dev))
    if matrix not in MATRICES:
        raise ValueError("matrix must be one of {}".format(sorted(MATRICES)))
    return Nv12(yp, up, ypitch, upitch, w, h, MATRICES[matrix], int(bool(full_range))), (out_y, out_uv)


def rgb_to_nv12(rgb, out_y=None, out_uv=None, matrix="bt709", full_range=False, stream=None):
    """uint8 RGB ``(H, W, 3)`` device array -> NV12 planes ``(y, uv)`` (what a hardware encoder takes);
    the integer conversion documented in include/nphusl_cuda.h, chroma = rounded mean of each 2 x 2 block."""
    if not _is_device(rgb):
        raise ValueError("rgb_to_nv12 takes a device array")
    src = _describe_device(rgb, "rgb")
    if src.dtype != np.uint8 or len(src.shape) != 3 or src.shape[-1] != 3:
        raise ValueError("rgb_to_nv12 takes uint8 (H, W, 3), got {} {}".format(src.dtype, src.shape))
    dev = _pick_device(None, src)
    dst, ret = _nv12_dest(src.shape[0], src.shape[1], out_y, out_uv, matrix, full_range, dev, stream)
    _touch(stream, src.keep)
    _check(load().nphusl_cuda_rgb_to_nv12(src.ptr, ctypes.byref(dst), dev, _stream_ptr(stream)), "rgb_to_nv12")
    return ret


def _nv12_edit_nv12_host(y, uv, edit, out_y, out_uv, matrix, full_range, out_matrix, out_full_range, stream, device, blocking):
    """HOST planes (frames as a capture card or a software decoder leaves them): 1.5 bytes per pixel
    over the link each way instead of the 3 of packed RGB.  Pinned arrays (``pinned_empty``) are
    DMA'd directly and with ``blocking=False`` the call only enqueues on ``stream``."""
    lib = load()
    y = np.ascontiguousarray(y, dtype=np.uint8)
    if uv is None or y.ndim != 2:
        raise ValueError("host NV12 input: pass the luma plane (H, W) and the interleaved chroma plane (H/2, W)")
    uv = np.ascontiguousarray(uv, dtype=np.uint8)
    h, w = y.shape
    if h % 2 or w % 2 or uv.shape[0] != h // 2 or uv.size != (h // 2) * w:
        raise ValueError("host NV12 input: even width and height, chroma plane of {} x {} bytes; got {} and {}".format(
            h // 2, w, y.shape, uv.shape))
    dev = _default_device if device is None else int(device)
    if out_y is None:
        out_y = _pinned_pool.empty((h, w), np.uint8)
        out_y = np.empty((h, w), np.uint8) if out_y is None else out_y
    if out_uv is None:
        out_uv = _pinned_pool.empty(uv.shape, np.uint8)
        out_uv = np.empty(uv.shape, np.uint8) if out_uv is None else out_uv
    for name, o, like in (("out_y", out_y, y), ("out_uv", out_uv, uv)):
        if not isinstance(o, np.ndarray) or o.dtype != np.uint8 or o.shape != like.shape or not o.flags.c_contiguous:
            raise ValueError("{}: expected a C-contiguous uint8 host array of shape {}".format(name, like.shape))
    sp = _stream_ptr(stream)
    dy, duv = DeviceArray((h, w), np.uint8, dev, stream), DeviceArray((h // 2, w), np.uint8, dev, stream)
    _check(lib.nphusl_cuda_memcpy(dy.ptr, y.ctypes.data, y.nbytes, 1, dev, sp), "memcpy h2d")
    _check(lib.nphusl_cuda_memcpy(duv.ptr, uv.ctypes.data, uv.nbytes, 1, dev, sp), "memcpy h2d")
    nv12_edit_nv12(dy, duv, edit, out_y=dy, out_uv=duv, matrix=matrix, full_range=full_range,      # in place on the device
                   out_matrix=out_matrix, out_full_range=out_full_range, stream=stream)
    _check(lib.nphusl_cuda_memcpy(out_y.ctypes.data, dy.ptr, y.nbytes, 2, dev, sp), "memcpy d2h")
    _check(lib.nphusl_cuda_memcpy(out_uv.ctypes.data, duv.ptr, uv.nbytes, 2, dev, sp), "memcpy d2h")
    if blocking:
        synchronize(dev, stream)
    return out_y, out_uv


def nv12_edit_nv12(y, uv=None, edit=None, out_y=None, out_uv=None, matrix="bt709", full_range=False,
                   out_matrix=None, out_full_range=None, stream=None, device=None, blocking=True, **edit_args):
    """The whole video loop in one pass: NV12 -> HUSL -> select + edit -> NV12 (3 bytes of memory
    traffic per pixel).  Equals ``rgb_to_nv12(nv12_edit_rgb(...))`` bit for bit; ``out_y`` / ``out_uv``
    may be the source planes themselves.  Returns ``(y, uv)``.  Host planes (numpy arrays) are uploaded,
    edited and downloaded (``device`` / ``blocking`` apply to that form only)."""
    if edit is None:
        edit = Edit(**edit_args)
    elif edit_args:
        raise TypeError("pass either an Edit or its fields as keywords, not both")
    edit = _need_edit(edit, "nv12_edit_nv12")
    if isinstance(y, np.ndarray):
        return _nv12_edit_nv12_host(y, uv, edit, out_y, out_uv, matrix, full_range, out_matrix, out_full_range,
                                    stream, device, blocking)
    src, (h, w), dev, keep = _nv12_surface(y, uv, matrix, full_range)
    dst, ret = _nv12_dest(h, w, out_y, out_uv, out_matrix or matrix, full_range if out_full_range is None else out_full_range, dev, stream)
    _touch(stream, *keep)
    _check(load().nphusl_cuda_nv12_edit_nv12(ctypes.byref(src), ctypes.byref(edit), ctypes.byref(dst), dev, _stream_ptr(stream)),
           "nv12_edit_nv12")
    return ret


# ---------------------------------------------------------------------------
# lookup tables (optional LUT mode)
# ---------------------------------------------------------------------------

LUT_TYPES = {"uint16": 0, "float": 1, "double": 2}
_LUT_DTYPES = {0: np.dtype(np.uint16), 1: np.dtype(np.float32), 2: np.dtype(np.float64)}


def lut_generate(light_seg=1024, nh=1024, nl=1024, device=None, table_type="uint16", int_scale=100):
    """Build the reference's light / chroma / linear tables ON THE DEVICE
    (make_lookup_tables:4-9, LUT/light.py, LUT/chroma.py, _linear_lookup.c).
    ``table_type``: "uint16" (the default build: round(value * int_scale)), "float"
    or "double" (the value to 4 / 3 decimals, scales 1), as `make_lookup_tables
    <table_type> <light size> <chroma size>` would bake them into _simd.c."""
    dev = _default_device if device is None else device
    _check(load().nphusl_cuda_lut_generate_typed(dev, light_seg, nh, nl, LUT_TYPES[table_type], int(int_scale)), "lut_generate")


def lut_upload(light, y_idx_step, y_thresh, chroma, h_idx_step, l_idx_step, linear, device=None,
               light_scale=None, chroma_scale=None):
    """Install caller-made tables (e.g. the reference's own).  uint16, float32 or float64
    tables; the scales default to 100 for integer tables and 1 for float tables."""
    dev = _default_device if device is None else device
    light, chroma = np.asarray(light), np.asarray(chroma)
    if np.issubdtype(chroma.dtype, np.integer):
        dt, kind, default_scale = np.uint16, 0, 100.0
    elif chroma.dtype == np.float32:
        dt, kind, default_scale = np.float32, 1, 1.0
    else:
        dt, kind, default_scale = np.float64, 2, 1.0
    light = np.ascontiguousarray(light, dt)
    chroma = np.ascontiguousarray(chroma, dt)
    st = np.ascontiguousarray(y_idx_step, np.float64)
    th = np.ascontiguousarray(y_thresh, np.float64)
    lin = np.ascontiguousarray(linear, np.float64)
    assert chroma.ndim == 2 and light.size % 3 == 0 and lin.size == 256
    _check(load().nphusl_cuda_lut_upload_typed(
        dev, kind, float(default_scale if light_scale is None else light_scale),
        float(default_scale if chroma_scale is None else chroma_scale),
        light.ctypes.data, light.size // 3, st.ctypes.data_as(_dp), th.ctypes.data_as(_dp),
        chroma.ctypes.data, chroma.shape[0], chroma.shape[1], float(h_idx_step), float(l_idx_step),
        lin.ctypes.data_as(_dp)), "lut_upload")


def lut_download(device=None):
    """-> dict(light, chroma, linear, y_idx_step, y_thresh, h_idx_step, l_idx_step, table_type,
    light_scale, chroma_scale); the tables come back in their element type"""
    dev = _default_device if device is None else device
    dims = (ctypes.c_int * 4)()
    steps = (ctypes.c_double * 7)()
    scales = (ctypes.c_double * 2)()
    _check(load().nphusl_cuda_lut_download_typed(dev, None, None, None, dims, steps, scales), "lut_download")
    seg, nh, nl, kind = dims[0], dims[1], dims[2], dims[3]
    dt = _LUT_DTYPES[kind]
    light = np.empty(3 * seg, dt)
    chroma = np.empty((nh, nl), dt)
    lin = np.empty(256, np.float64)
    _check(load().nphusl_cuda_lut_download_typed(dev, light.ctypes.data, chroma.ctypes.data,
                                                 lin.ctypes.data_as(_dp), dims, steps, scales), "lut_download")
    s = list(steps)
    return {"light": light, "chroma": chroma, "linear": lin, "y_idx_step": np.array(s[0:3]),
            "y_thresh": np.array(s[3:5]), "h_idx_step": s[5], "l_idx_step": s[6],
            "table_type": {v: k for k, v in LUT_TYPES.items()}[kind],
            "light_scale": scales[0], "chroma_scale": scales[1]}
