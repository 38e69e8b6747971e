"""Host-side logic of the ctypes binding that needs no GPU: buffer description,
the Edit struct layout against the C header, DeviceArray view arithmetic."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from nphusl import _cuda


def test_edit_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "nphusl_cuda.h")).read()
    body = re.search(r"typedef struct nphusl_edit \{(.*?)\} nphusl_edit;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, rest = decl.split(None, 1)
        for n in rest.split(","):
            names.append((n.strip(), ctype))
    assert [(n, {"float": ctypes.c_float, "int": ctypes.c_int}[t]) for n, t in names] == list(_cuda.Edit._fields_)
    assert ctypes.sizeof(_cuda.Edit) == 13 * 4
    e = _cuda.Edit(hue=(250, 290), invert=True, l=(0.5, 0))
    assert (e.h_lo, e.h_hi, e.invert, e.l_mul, e.l_add, e.s_mul) == (250.0, 290.0, 1, 0.5, 0.0, 1.0)
    assert e.s_lo == -np.inf and e.l_hi == np.inf            # unspecified ranges are unbounded


def test_bands_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "nphusl_cuda.h")).read()
    body = re.search(r"typedef struct nphusl_bands \{(.*?)\} nphusl_bands;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            ctype, rest = decl.split(None, 1)
            names += [(n.strip(), {"float": ctypes.c_float, "int": ctypes.c_int}[ctype]) for n in rest.split(",")]
    assert names == list(_cuda.Bands._fields_) and ctypes.sizeof(_cuda.Bands) == 7 * 4
    assert re.search(r"NPHUSL_SRC_H = 0, NPHUSL_SRC_S = 1, NPHUSL_SRC_L = 2", text)
    b = _cuda.Bands("l", 7, h_even=130, h_odd=360, stripe=(2, 100))       # one frame of melonize, examples.py:84-103
    assert (b.src, b.width, b.total, b.h_even, b.h_odd, b.s_half, b.s_value) == (2, 7.0, 100.0, 130.0, 360.0, 2.0, 100.0)
    w = _cuda.Bands()                                                     # hue_watermelon, examples.py:56-71
    assert (w.src, w.width, w.total, w.s_half) == (0, 45.0, 360.0, 0.0)
    for bad in (dict(src="v"), dict(width=0), dict(width=-1), dict(total=0)):
        with pytest.raises(ValueError):
            _cuda.Bands(**bad)
    with pytest.raises(TypeError):
        _cuda._need_edit(w, "edit_husl")                                  # Edit-only entry points refuse a Bands
    assert isinstance(_cuda._need_edit(dict(l=(0.5, 0)), "x"), _cuda.Edit)


def test_rgb_input_description():
    d = _cuda._describe_rgb_in(np.zeros((4, 5, 3), np.uint8))
    assert (d.loc, d.channels, d.dtype, d.shape) == (_cuda.HOST, 3, np.uint8, (4, 5, 3))
    assert _cuda._describe_rgb_in(np.zeros((4, 3), np.int32)).dtype == np.uint8        # other ints: cast like rgb_int_input
    assert _cuda._describe_rgb_in(np.zeros((4, 3), np.float16)).dtype == np.float64
    assert _cuda._describe_rgb_in(np.zeros((4, 3), np.float32)).dtype == np.float32
    grey = np.random.default_rng(0).random((6, 7))
    g3 = np.broadcast_to(grey[..., None], (6, 7, 3))
    d = _cuda._describe_rgb_in(g3)
    assert d.channels == 1 and d.shape == (6, 7, 3) and d.keep.shape == (6, 7)          # one channel travels
    nc = np.zeros((8, 8, 3), np.uint8)[:, ::2]
    d = _cuda._describe_rgb_in(nc)
    assert d.keep.flags.c_contiguous and d.keep.shape == (8, 4, 3)                      # compacted copy


class FakeDev:
    def __init__(self, shape, typestr, strides=None, ptr=4096, device=None):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": strides}
        if device is not None:
            self.device = device


def test_device_description_and_contiguity():
    d = _cuda._describe_device(FakeDev((4, 5, 3), "|u1", device=2), "x")
    assert (d.loc, d.device, d.ptr, d.shape) == (_cuda.DEVICE, 2, 4096, (4, 5, 3))
    _cuda._describe_device(FakeDev((4, 5, 3), "<f4", strides=(60, 12, 4), device=0), "x")      # explicit C strides: fine
    _cuda._describe_device(FakeDev((1, 5, 3), "<f4", strides=(999, 12, 4), device=0), "x")     # size-1 axis: stride ignored
    with pytest.raises(ValueError):
        _cuda._describe_device(FakeDev((4, 5, 3), "<f4", strides=(120, 12, 4), device=0), "x")
    g = _cuda._describe_device(FakeDev((4, 5, 3), "<f8", strides=(40, 8, 0), device=0), "x", allow_gray=True)
    assert g.channels == 1 and g.shape == (4, 5, 3)
    with pytest.raises(ValueError):
        _cuda._describe_device(FakeDev((4, 5, 3), "<f8", strides=(40, 8, 0), device=0), "x")


def test_out_validation():
    with pytest.raises(ValueError):
        _cuda._describe_out(np.zeros((4, 3)), (5, 3), np.float64, _cuda._FLOATS, "out")           # wrong size
    with pytest.raises(ValueError):
        _cuda._describe_out(np.zeros((4, 3), np.int16), (4, 3), np.float64, _cuda._FLOATS, "out")  # wrong dtype
    with pytest.raises(ValueError):
        _cuda._describe_out(np.zeros((4, 6))[:, ::2], (4, 3), np.float64, _cuda._FLOATS, "out")    # not contiguous
    with pytest.raises(TypeError):
        _cuda._describe_out([1, 2, 3], (1, 3), np.float64, _cuda._FLOATS, "out")


def test_device_array_views_without_gpu():
    a = object.__new__(_cuda.DeviceArray)
    a.shape, a.dtype, a.device, a.size, a.nbytes, a.ptr, a._owner = (3, 4, 5), np.dtype(np.float32), 0, 60, 240, 1 << 20, False
    assert a.reshape(-1, 5).shape == (12, 5) and a.reshape((60,)).shape == (60,)
    with pytest.raises(ValueError):
        a.reshape(7, -1)
    v = a[2]
    assert v.shape == (4, 5) and v.ptr == a.ptr + 2 * 80 and v.nbytes == 80
    assert a[-1].ptr == v.ptr
    s = a[1:3]
    assert s.shape == (2, 4, 5) and s.ptr == a.ptr + 80
    assert s.__cuda_array_interface__["data"][0] == s.ptr and s.__cuda_array_interface__["strides"] is None
    with pytest.raises(IndexError):
        a[::2]
    with pytest.raises(IndexError):
        a[3]


def test_stream_pointer_forms():
    class TorchLike:
        cuda_stream = 1234
    assert _cuda._stream_ptr(None) is None
    assert _cuda._stream_ptr(TorchLike()).value == 1234
    assert _cuda._stream_ptr(77).value == 77


def test_strided_views_collapse_to_three_axes():
    """_views: joint collapse of the pixel axes of (src, dst) into frames x rows x
    cols with byte strides (include/nphusl_cuda.h nphusl_view); pure host logic."""
    from nphusl import _cuda

    def buf(shape, dtype, strides=None, ptr=4096):
        b = _cuda._Buf()
        b.ptr, b.shape, b.dtype, b.strides = ptr, tuple(shape), np.dtype(dtype), strides
        return b

    def dims(v):
        return (v.frames, v.rows, v.cols, v.frame_stride, v.row_stride, v.col_stride, v.ch_stride)

    # batch of cropped frames: (B, h, w, 3) view of (B, H, W, 4) uint8 -> float32 contiguous result
    B, H, W, h, w = 5, 40, 60, 30, 20
    src = buf((B, h, w, 3), np.uint8, (H * W * 4, W * 4, 4, 1))
    dst = buf((B, h, w, 3), np.float32)
    vi, vo = _cuda._views([src, dst], 3)
    assert dims(vi) == (B, h, w, H * W * 4, W * 4, 4, 1)
    assert dims(vo) == (B, h, w, h * w * 12, w * 12, 12, 4)
    # single cropped frame: frames = 1
    vi, vo = _cuda._views([buf((h, w, 3), np.uint8, (W * 4, 4, 1)), buf((h, w, 3), np.float64)], 2)
    assert dims(vi)[:3] == (1, h, w) and dims(vi)[4:] == (W * 4, 4, 1)
    assert dims(vo)[:3] == (1, h, w) and dims(vo)[4:] == (w * 24, 24, 8)
    # channel slice of full frames: every pixel axis merges -> one row; hue image has no channel axis
    vi, vo = _cuda._views([buf((B, H, W, 3), np.uint8, (H * W * 4, W * 4, 4, 1)), buf((B, H, W), np.float32)], 3)
    assert dims(vi)[:3] == (1, 1, B * H * W) and vi.col_stride == 4
    assert dims(vo)[:3] == (1, 1, B * H * W) and vo.col_stride == 4 and vo.ch_stride == 0
    # size-1 axes never get in the way
    vi, _ = _cuda._views([buf((1, h, 1, w, 3), np.uint8, (999, W * 4, 777, 4, 1)), buf((1, h, 1, w, 3), np.float32)], 4)
    assert dims(vi)[:3] == (1, h, w) and (vi.row_stride, vi.col_stride) == (W * 4, 4)
    # four independent pixel axes cannot be expressed
    with pytest.raises(ValueError):
        _cuda._views([buf((4, 3, 4, 5, 3), np.uint8, (10000, 2000, 200, 8, 1)), buf((4, 3, 4, 5, 3), np.float32)], 4)


def test_pinned_result_pool_recycles_blocks(monkeypatch):
    """_PinnedPool (page-locked results of host-array calls): block reuse, budgets and the
    plain-ndarray contract, with malloc/free standing in for cudaHostAlloc on this CPU-only box."""
    import ctypes
    import gc
    from nphusl import _cuda
    libc = ctypes.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = ctypes.c_void_p, [ctypes.c_size_t]
    libc.free.argtypes = [ctypes.c_void_p]
    live = set()

    class FakeLib:
        @staticmethod
        def nphusl_cuda_host_alloc(pp, n):
            p = libc.malloc(n)
            pp._obj.value = p
            live.add(p)
            return 0

        @staticmethod
        def nphusl_cuda_host_free(p):
            live.discard(p.value)
            libc.free(p)
            return 0

    monkeypatch.setattr(_cuda, "load", lambda: FakeLib)
    monkeypatch.setattr(_cuda, "_lib", FakeLib)
    pool = _cuda._PinnedPool()
    monkeypatch.setattr(_cuda, "_pinned_pool", pool)
    pool.enabled, pool.max_cached, pool.max_outstanding = True, 64 << 20, 96 << 20
    shape = (1080, 1920, 3)
    a = pool.empty(shape, np.float64)
    assert type(a) is np.ndarray and a.shape == shape and a.dtype == np.float64 and a.flags.writeable and a.flags.c_contiguous
    a[...] = 1.5
    view = a[..., 0]                                           # views keep the block alive
    ptr = a.ctypes.data
    del a
    gc.collect()
    assert pool.cached == 0 and pool.outstanding > 0 and view[5, 5] == 1.5
    del view
    gc.collect()
    assert pool.outstanding == 0 and pool.cached == (1080 * 1920 * 24 + pool.GRAIN - 1) // pool.GRAIN * pool.GRAIN
    b = pool.empty(shape, np.float64)                          # same size class: the block comes back
    assert b.ctypes.data == ptr and pool.hits == 1 and pool.misses == 1 and pool.cached == 0
    assert pool.empty((100, 3), np.float64) is None            # small results stay pageable
    c = pool.empty(shape, np.float64)
    assert pool.empty(shape, np.float64) is None               # outstanding budget (96 MiB) exhausted
    del b, c
    gc.collect()
    assert pool.cached <= pool.max_cached and len(live) == 1   # only one 48 MiB block fits the 64 MiB cache
    pool.trim()
    assert not live and pool.cached == 0
    pool.enabled = False
    assert pool.empty(shape, np.float64) is None


def test_device_pool_size_classes_and_cheap_device_check():
    from nphusl import _cuda
    sc = _cuda._DevicePool.size_class
    assert sc(1) == 512 and sc(512) == 512 and sc(513) == 1024 and sc(100000) == 131072
    assert sc((1 << 20) - 1) == 1 << 20 and sc(1 << 20) == 2 << 20 and sc((2 << 20) + 1) == 4 << 20
    assert sc(1080 * 1920 * 24) % (2 << 20) == 0 and sc(1080 * 1920 * 24) >= 1080 * 1920 * 24

    class FakeTorchCpu:
        is_cuda = False

        @property
        def __cuda_array_interface__(self):
            raise AssertionError("must not be evaluated for a host tensor")

    class FakeTorchCuda(FakeTorchCpu):
        is_cuda = True

    class Other:
        __cuda_array_interface__ = {"shape": (1,), "typestr": "|u1", "data": (0, False), "version": 3}

    assert not _cuda._is_device(np.zeros(3)) and not _cuda._is_device(FakeTorchCpu()) and not _cuda._is_device([1, 2, 3])
    assert _cuda._is_device(FakeTorchCuda()) and _cuda._is_device(Other())
    assert _cuda._prod((2, 3, 4)) == 24 and _cuda._prod(()) == 1 and _cuda._prod(7) == 7


def test_nv12_struct_matches_header_and_surface_description():
    """nphusl_nv12 layout against the header, and the surface descriptor built from device planes
    (pitched planes, (rows, pairs, 2) chroma, one whole NVDEC surface) -- no GPU needed."""
    text = open(os.path.join(ROOT, "include", "nphusl_cuda.h")).read()
    body = re.search(r"typedef struct nphusl_nv12 \{(.*?)\} nphusl_nv12;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = re.sub(r"\b(const|void|long|int)\b|\*", " ", decl)
        names += [n.strip() for n in decl.split(",") if n.strip()]
    assert names == [f[0] for f in _cuda.Nv12._fields_]
    y = FakeDev((1080, 1920), "|u1", strides=(2048, 1), ptr=1 << 20, device=0)            # pitched luma
    uv = FakeDev((540, 960, 2), "|u1", strides=(2048, 2, 1), ptr=1 << 23, device=0)       # pairs as a last axis
    s, (h, w), dev, _ = _cuda._nv12_surface(y, uv, "bt709", False)
    assert (s.y, s.uv, s.y_pitch, s.uv_pitch, s.width, s.height, s.matrix, s.full_range) == (1 << 20, 1 << 23, 2048, 2048, 1920, 1080, 1, 0)
    assert (h, w, dev) == (1080, 1920, 0)
    whole = FakeDev((1620, 1920), "|u1", strides=(2048, 1), ptr=1 << 20, device=1)       # luma rows, then chroma rows
    s, (h, w), dev, _ = _cuda._nv12_surface(whole, None, "bt601", True)
    assert (s.uv - s.y, s.uv_pitch, h, w, dev, s.matrix, s.full_range) == (1080 * 2048, 2048, 1080, 1920, 1, 0, 1)
    odd = FakeDev((5, 7), "|u1", device=0)
    s, (h, w), _, _ = _cuda._nv12_surface(odd, FakeDev((3, 8), "|u1", device=0), "bt709", False)   # odd sizes: ceil(h/2) x 2*ceil(w/2)
    assert (h, w) == (5, 7)
    for bad_uv in (FakeDev((540, 1918), "|u1", device=0), FakeDev((539, 1920), "|u1", device=0), FakeDev((540, 1920), "|u1", device=1),
                   FakeDev((540, 1920), "<f4", device=0), FakeDev((540, 960, 2), "|u1", strides=(4096, 4, 1), device=0)):
        with pytest.raises(ValueError):
            _cuda._nv12_surface(FakeDev((1080, 1920), "|u1", device=0), bad_uv, "bt709", False)
    with pytest.raises(ValueError):
        _cuda._nv12_surface(y, uv, "bt2020", False)
    with pytest.raises(ValueError):
        _cuda._nv12_surface(np.zeros((4, 4), np.uint8), None, "bt709", False)             # host planes are refused
    with pytest.raises(ValueError):
        _cuda._nv12_surface(FakeDev((1621, 1920), "|u1", device=0), None, "bt709", False)  # not 3/2 * height rows


def _fake_runtime(monkeypatch):
    """malloc/free standing in for the CUDA allocation calls of the C ABI, plus scriptable events."""
    import ctypes
    from nphusl import _cuda
    libc = ctypes.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = ctypes.c_void_p, [ctypes.c_size_t]
    libc.free.argtypes = [ctypes.c_void_p]

    class FakeLib:
        live, host_live = set(), set()
        events, done, waits, next_event, destroyed = {}, set(), [], [1000], []

        @classmethod
        def nphusl_cuda_malloc(cls, pp, n, dev):
            p = libc.malloc(max(n, 64))
            pp._obj.value = p
            cls.live.add(p)
            return 0

        @classmethod
        def nphusl_cuda_free(cls, p, dev):
            cls.live.discard(p.value)
            libc.free(p)
            return 0

        @classmethod
        def nphusl_cuda_host_alloc(cls, pp, n):
            p = libc.malloc(max(n, 64))
            pp._obj.value = p
            cls.host_live.add(p)
            return 0

        @classmethod
        def nphusl_cuda_host_free(cls, p):
            cls.host_live.discard(p.value)
            libc.free(p)
            return 0

        @classmethod
        def nphusl_cuda_event_record(cls, pev, stream, dev):
            if not pev._obj.value:
                cls.next_event[0] += 1
                pev._obj.value = cls.next_event[0]
            cls.events[pev._obj.value] = stream.value or 0
            cls.done.discard(pev._obj.value)
            return 0

        @classmethod
        def nphusl_cuda_event_query(cls, ev, pdone):
            pdone._obj.value = 1 if ev.value in cls.done else 0
            return 0

        @classmethod
        def nphusl_cuda_stream_wait_event(cls, stream, ev, dev):
            cls.waits.append((stream.value or 0, cls.events[ev.value]))
            return 0

        @classmethod
        def nphusl_cuda_event_destroy(cls, ev, dev):
            cls.destroyed.append(ev.value)
            return 0

    monkeypatch.setattr(_cuda, "load", lambda: FakeLib)
    monkeypatch.setattr(_cuda, "_lib", FakeLib)
    return FakeLib


def test_device_pool_orders_block_reuse_by_stream(monkeypatch):
    """_DevicePool: a dropped block is reused on its own stream by stream order, on another stream only
    behind an event wait, and a block dropped while in use on several streams only after all of their
    events completed (free lists keyed by (device, size) alone let stream s2 overwrite a block
    a kernel on s1 was still reading)."""
    from nphusl import _cuda
    fake = _fake_runtime(monkeypatch)
    pool = _cuda._DevicePool()
    n = 3 << 20
    p1, cap = pool.alloc(n, 0, 11)
    assert pool.misses == 1 and cap == 4 << 20
    pool.release(p1, cap, 0, {11})
    p2, _ = pool.alloc(n, 0, 11)                         # same stream: plain reuse, no event traffic
    assert p2 == p1 and pool.hits == 1 and not fake.waits and not fake.events
    pool.release(p2, cap, 0, {11})
    p3, _ = pool.alloc(n, 0, 22)                         # other stream: 22 waits for the tail of 11
    assert p3 == p1 and fake.waits == [(22, 11)] and pool.waits == 1
    pool.release(p3, cap, 0, {22, 33})                   # in use on two streams when dropped: parked
    assert len(pool.parked) == 1 and sorted(fake.events.values())[-2:] == [22, 33]
    p4, _ = pool.alloc(n, 0, 22)                         # events pending: NOT handed out, not even to 22
    assert p4 != p1 and pool.misses == 2
    fake.done.update(fake.events)                        # both streams passed their events
    p5, _ = pool.alloc(n, 0, 44)
    assert p5 == p1 and not pool.parked and len(fake.waits) == 1      # idle on every stream: no wait needed
    pool.release(p4, cap, 0, ())                         # never used by a kernel
    pool.release(p5, cap, 0, {44})
    assert pool.cached == 2 * cap
    # a stream that was drained and destroyed (Stream.close) is forgotten: its blocks are idle everywhere and
    # its handle is never given to the driver again
    p8, _ = pool.alloc(n, 0, 55)
    assert p8 in (p1, p4)
    pool.release(p8, cap, 0, {55})
    n_events = len(fake.events)
    pool.retire(55)
    p9, _ = pool.alloc(n, 0, 66)
    assert p9 == p8 and len(fake.events) == n_events and len(fake.waits) == 1
    pool.release(p9, cap, 0, {55, 66})                   # 55 is retired: one live stream left -> plain free list
    assert not pool.parked
    pool.revive(55)
    # a second device never sees device 0's blocks
    p6, _ = pool.alloc(n, 1, 44)
    assert p6 not in (p1, p4) and pool.misses == 3
    pool.release(p6, cap, 1, {44})
    pool.trim()
    assert not fake.live and pool.cached == 0 and sorted(fake.destroyed) == sorted(set(fake.events))
    # over budget: released straight to the driver
    pool.max_cached = 0
    p7, _ = pool.alloc(n, 0, 0)
    pool.release(p7, cap, 0, {0})
    assert not fake.live


def test_device_array_records_the_streams_it_is_used_on(monkeypatch):
    from nphusl import _cuda
    _fake_runtime(monkeypatch)
    monkeypatch.setattr(_cuda, "_device_pool", _cuda._DevicePool())

    class S:
        def __init__(self, h):
            self.cuda_stream = h
    a = _cuda.DeviceArray((4, 6, 3), np.float32, 0, S(7))
    assert a._streams == {7}
    v = a[1:3].reshape(-1, 3)
    _cuda._touch(S(9), v, np.zeros(3), None)             # views record on the owner; host arrays are ignored
    assert a._streams == {7, 9}
    a.record_stream(None)
    assert a._streams == {0, 7, 9}

    class Wrapper:                                        # transform.GrayDevice / RgbaDevice keep the array in .base
        base = a
    _cuda._touch(5, Wrapper())
    assert a._streams == {0, 5, 7, 9}
    ptr = a.ptr
    del v
    a.free()
    assert len(_cuda._device_pool.parked) == 1 and _cuda._device_pool.parked[0][0] == ptr


def test_pinned_empty_block_outlives_every_view(monkeypatch):
    """pinned_empty returns a plain ndarray whose .base chain owns the page-locked block: base-class views
    (np.asarray, .view(np.ndarray), slices) keep it alive (an ndarray subclass with __del__ freed
    the block while such views -- and in-flight DMA -- still used it)."""
    import gc
    from nphusl import _cuda
    fake = _fake_runtime(monkeypatch)
    a = _cuda.pinned_empty((64, 5), np.float64)
    assert type(a) is np.ndarray and a.shape == (64, 5) and a.flags.writeable and a.flags.c_contiguous
    assert len(fake.host_live) == 1
    a[...] = 2.5
    views = [np.asarray(a), a.view(np.ndarray), np.ascontiguousarray(a), a[3:9, 1], a.reshape(-1).view(np.uint8)]
    del a
    gc.collect()
    assert len(fake.host_live) == 1 and all(v.flat[0] in (2.5, 0) for v in views[:4]) and views[0][5, 2] == 2.5
    while views:
        views.pop()
        gc.collect()
        assert len(fake.host_live) == (1 if views else 0)
    assert _cuda.pinned_empty(7, np.uint8).shape == (7,)


def test_out_must_have_the_result_shape():
    from nphusl import _cuda
    ok = _cuda._describe_out(np.empty((1, 4, 5, 3)), (4, 5, 3), np.float64, _cuda._FLOATS, "out")     # unit axes may differ
    assert ok.shape == (1, 4, 5, 3)
    _cuda._describe_out(np.empty(3), (1, 3), np.float64, _cuda._FLOATS, "out")                          # the squeezed single pixel
    for bad in ((3, 20), (60,), (5, 4, 3)):
        with pytest.raises(ValueError, match="shape"):
            _cuda._describe_out(np.empty(bad), (4, 5, 3), np.float64, _cuda._FLOATS, "out")
