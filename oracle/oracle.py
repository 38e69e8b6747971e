"""ctypes front-end of the CPU oracle (``husl_oracle.c``) and of the reference
builds under ``oracle/_ref/``.

TEST INFRASTRUCTURE, NOT PRODUCT: imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs.  The product
package never imports this module and has no CPU fallback.

* ``rgb_to_husl / rgb_to_hue / husl_to_rgb[_float]`` -- the float64
  restatement of the reference NumPy path (nphusl/nphusl.py:110-392);
* ``simd_rgb_to_husl(variant=...)`` -- restatement of nphusl/_simd.c;
* ``light_table / chroma_table / linear_table6`` -- restatement of
  LUT/light.py, LUT/chroma.py and nphusl/_linear_lookup.c:45-75;
* ``nv12_to_rgb`` -- numpy statement of the integer BT.601 / BT.709 conversion that
  include/nphusl_cuda.h specifies for decoder surfaces (not reference code: the
  reference's callers decode with imageio / moviepy, examples.py:117-134), plus
  ``nv12_to_rgb_float`` (the textbook float matrix, to bound the integer one), ``rgb_to_nv12`` (the encode side);
* ``ref_simd(...)``, ``ref_tables()``, ``ref_cython()``, ``ref_package()`` -- the reference's own
  compiled code (``oracle/_ref``, built by ``oracle/build_ref.sh`` from the
  sources under /root/reference; present on the GPU box as prebuilt files).
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_LIB = None

_c_u8p = ctypes.POINTER(ctypes.c_uint8)
_c_u16p = ctypes.POINTER(ctypes.c_uint16)
_c_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile husl_oracle.c (gcc, seconds)."""
    so = os.path.join(HERE, "libhusl_oracle.so")
    src = os.path.join(HERE, "husl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_max_chroma.restype = ctypes.c_double
        _LIB.oracle_max_chroma.argtypes = [ctypes.c_double, ctypes.c_double]
        _LIB.oracle_omp_threads.restype = ctypes.c_int
    return _LIB


def _p(arr, typ):
    return arr.ctypes.data_as(typ)


def _rgb_in(rgb):
    rgb = np.asarray(rgb)
    if rgb.dtype == np.uint8:
        return np.ascontiguousarray(rgb), True
    return np.ascontiguousarray(rgb, dtype=np.float64), False


def rgb_to_husl(rgb):
    """uint8 RGB (scaled by /255, transform.py:35-37) or float RGB in [0,1]
    -> float64 HUSL, same shape."""
    a, is_u8 = _rgb_in(rgb)
    out = np.empty(a.shape, dtype=np.float64)
    n = a.size // 3
    if is_u8:
        lib().oracle_rgb_u8_to_husl(_p(a, _c_u8p), _p(out, _c_f64p), ctypes.c_size_t(n))
    else:
        lib().oracle_rgb_to_husl_f64(_p(a, _c_f64p), _p(out, _c_f64p), ctypes.c_size_t(n))
    return out


def rgb_to_hue(rgb):
    a, is_u8 = _rgb_in(rgb)
    out = np.empty(a.shape[:-1], dtype=np.float64)
    n = a.size // 3
    if is_u8:
        lib().oracle_rgb_u8_to_hue(_p(a, _c_u8p), _p(out, _c_f64p), ctypes.c_size_t(n))
    else:
        lib().oracle_rgb_to_hue_f64(_p(a, _c_f64p), _p(out, _c_f64p), ctypes.c_size_t(n))
    return out


def husl_to_rgb_float(hsl):
    """HUSL -> float64 RGB in [0,1] (what the reference's _husl_to_rgb returns)."""
    a = np.ascontiguousarray(hsl, dtype=np.float64)
    out = np.empty(a.shape, dtype=np.float64)
    lib().oracle_husl_to_rgb_f64(_p(a, _c_f64p), _p(out, _c_f64p), ctypes.c_size_t(a.size // 3))
    return out


def husl_to_rgb(hsl):
    """HUSL -> uint8 RGB (to_rgb's output: round-half-even, abs, uint8 cast)."""
    a = np.ascontiguousarray(hsl, dtype=np.float64)
    out = np.empty(a.shape, dtype=np.uint8)
    lib().oracle_husl_to_rgb_u8(_p(a, _c_f64p), _p(out, _c_u8p), ctypes.c_size_t(a.size // 3))
    return out


def _nv12_k(matrix, full_range):
    kr, kb = {"bt601": (0.299, 0.114), "bt709": (0.2126, 0.0722)}[matrix]
    kg = 1.0 - kr - kb
    ys, cs, yoff = (1.0, 1.0, 0) if full_range else (255.0 / 219.0, 255.0 / 224.0, 16)
    return ys, cs * 2 * (1 - kr), cs * 2 * kb * (1 - kb) / kg, cs * 2 * kr * (1 - kr) / kg, cs * 2 * (1 - kb), yoff


def _nv12_planes(y, uv):
    y = np.asarray(y).astype(np.int64)
    h, w = y.shape
    uv = np.asarray(uv).reshape((h + 1) // 2, (w + 1) // 2, 2).astype(np.int64)
    up = np.repeat(np.repeat(uv, 2, axis=0), 2, axis=1)[:h, :w]      # chroma replicated over its 2 x 2 block
    return y, up[..., 0] - 128, up[..., 1] - 128


def nv12_to_rgb(y, uv, matrix="bt709", full_range=False):
    """NV12 planes -> uint8 RGB (H, W, 3), the integer conversion of include/nphusl_cuda.h:
    16.16 fixed-point coefficients, round half up, clamp."""
    ys, rv, gu, gv, bu, yoff = _nv12_k(matrix, full_range)
    ky, rv, gu, gv, bu = (int(np.floor(65536.0 * v + 0.5)) for v in (ys, rv, gu, gv, bu))
    yv, d, e = _nv12_planes(y, uv)
    yv = ky * (yv - yoff) + 32768
    rgb = np.stack([(yv + rv * e) >> 16, (yv - (gu * d + gv * e)) >> 16, (yv + bu * d) >> 16], axis=-1)
    return np.clip(rgb, 0, 255).astype(np.uint8)


def nv12_to_rgb_float(y, uv, matrix="bt709", full_range=False):
    """The same conversion with the real-valued matrix, unrounded and unclamped (0..255 scale)."""
    ys, rv, gu, gv, bu, yoff = _nv12_k(matrix, full_range)
    yv, d, e = _nv12_planes(y, uv)
    yv = ys * (yv - yoff)
    return np.stack([yv + rv * e, yv - gu * d - gv * e, yv + bu * d], axis=-1)


def rgb_to_nv12(rgb, matrix="bt709", full_range=False):
    """uint8 RGB (H, W, 3), H and W even -> (y (H, W), uv (H/2, W)) uint8: the integer conversion of
    include/nphusl_cuda.h (16.16 coefficients, round half up, clamp; chroma from the rounded 2 x 2 mean)."""
    kr, kb = {"bt601": (0.299, 0.114), "bt709": (0.2126, 0.0722)}[matrix]
    kg = 1.0 - kr - kb
    ys, cs, yoff = (1.0, 1.0, 0) if full_range else (219.0 / 255.0, 224.0 / 255.0, 16)
    q = lambda v: int(np.floor(65536.0 * abs(v) + 0.5)) * (1 if v >= 0 else -1)     # lround: half away from zero
    yr, yg, yb = q(ys * kr), q(ys * kg), q(ys * kb)
    ur, ug, ub = q(-cs * 0.5 * kr / (1 - kb)), q(-cs * 0.5 * kg / (1 - kb)), q(cs * 0.5)
    vr, vg, vb = q(cs * 0.5), q(-cs * 0.5 * kg / (1 - kr)), q(-cs * 0.5 * kb / (1 - kr))
    a = np.asarray(rgb).astype(np.int64)
    h, w, _ = a.shape
    y = np.clip((yr * a[..., 0] + yg * a[..., 1] + yb * a[..., 2] + 32768 + (yoff << 16)) >> 16, 0, 255)
    m = (a.reshape(h // 2, 2, w // 2, 2, 3).sum(axis=(1, 3)) + 2) >> 2
    u = np.clip((ur * m[..., 0] + ug * m[..., 1] + ub * m[..., 2] + 32768 + (128 << 16)) >> 16, 0, 255)
    v = np.clip((vr * m[..., 0] + vg * m[..., 1] + vb * m[..., 2] + 32768 + (128 << 16)) >> 16, 0, 255)
    return y.astype(np.uint8), np.stack([u, v], axis=-1).reshape(h // 2, w).astype(np.uint8)


def max_chroma(l, h):
    return lib().oracle_max_chroma(float(l), float(h))


# ---- tables -------------------------------------------------------------

def linear_table6():
    t = np.empty(256, dtype=np.float64)
    lib().oracle_linear_table6(_p(t, _c_f64p))
    return t


_KINDS = {"uint16": (0, np.uint16, 100.0), "float": (1, np.float32, 1.0), "double": (2, np.float64, 1.0)}


def light_table(seg=1024, table_type="uint16"):
    """-> (table[3*seg], Y_IDX_STEP[3], Y_THRESH[2]); table dtype uint16 (L*100), float32 or float64
    (L to 4 decimals) as `LUT/light.py -t <table_type>` writes it"""
    kind, dt, scale = _KINDS[table_type]
    t = np.empty(3 * seg, dtype=np.float64)
    steps = np.empty(3, dtype=np.float64)
    thresh = np.empty(2, dtype=np.float64)
    lib().oracle_light_table_typed(_p(t, _c_f64p), ctypes.c_int(seg), _p(steps, _c_f64p), _p(thresh, _c_f64p),
                                   ctypes.c_int(kind), ctypes.c_double(scale))
    return t.astype(dt), steps, thresh


def chroma_table(nh=1024, nl=1024, table_type="uint16"):
    """-> (table[nh, nl], H_IDX_STEP, L_IDX_STEP); dtype as light_table (chroma to 3 decimals)"""
    kind, dt, scale = _KINDS[table_type]
    t = np.empty((nh, nl), dtype=np.float64)
    hs, ls = ctypes.c_double(), ctypes.c_double()
    lib().oracle_chroma_table_typed(_p(t, _c_f64p), ctypes.c_int(nh), ctypes.c_int(nl),
                                    ctypes.byref(hs), ctypes.byref(ls), ctypes.c_int(kind), ctypes.c_double(scale))
    return t.astype(dt), hs.value, ls.value


_TABLE_CACHE = {}


def table_scale(table):
    """LIGHT_SCALE / CHROMA_SCALE of a table: 100 for the integer types, 1 for float / double
    (LUT/light.py:31-32 with the default --int-scale)"""
    return 100.0 if np.issubdtype(np.asarray(table).dtype, np.integer) else 1.0


def simd_rgb_to_husl(rgb, variant="lut", tables=None):
    """Restatement of nphusl/_simd.c::rgb_to_husl_nd.  variant: "exact" (no -D
    flags), "lut" (setup.py default), "lut_interp" (+INTERPOLATE_CHROMA).
    ``tables`` = (light, ystep3, ythresh2, chroma, hstep, lstep) with uint16,
    float32 or float64 tables; made by the oracle's own uint16 generators when
    omitted."""
    v = {"exact": 0, "lut": 1, "lut_interp": 2}[variant]
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    out = np.empty(a.shape, dtype=np.float64)
    lin = _TABLE_CACHE.setdefault("lin", linear_table6())
    if tables is None:
        if "tab" not in _TABLE_CACHE:
            lt, st, th = light_table()
            ct, hs, ls = chroma_table()
            _TABLE_CACHE["tab"] = (lt, st, th, ct, hs, ls)
        tables = _TABLE_CACHE["tab"]
    lt, st, th, ct, hs, ls = tables
    lsc, csc = table_scale(lt), table_scale(ct)
    ct = np.asarray(ct)
    nh, nl = ct.shape
    lt = np.ascontiguousarray(lt, dtype=np.float64)          # every table type is exact in a double
    ct = np.ascontiguousarray(ct, dtype=np.float64)
    st = np.ascontiguousarray(st, dtype=np.float64)
    th = np.ascontiguousarray(th, dtype=np.float64)
    lib().oracle_simd_rgb_to_husl(
        _p(a, _c_u8p), _p(out, _c_f64p), ctypes.c_size_t(a.size // 3), ctypes.c_int(v),
        _p(lin, _c_f64p), _p(lt, _c_f64p), ctypes.c_int(lt.size // 3), ctypes.c_double(lsc),
        _p(st, _c_f64p), _p(th, _c_f64p), _p(ct, _c_f64p),
        ctypes.c_int(nh), ctypes.c_int(nl), ctypes.c_double(csc),
        ctypes.c_double(hs), ctypes.c_double(ls))
    return out


# ---- the reference's own compiled code (oracle/_ref) ---------------------

_REF_LIBS = {}
_LIBC = None


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libsimd_lut.so"))


def _ref_lib(name):
    if name not in _REF_LIBS:
        path = os.path.join(REF_DIR, "libsimd_%s.so" % name)
        l = ctypes.CDLL(path)
        l.rgb_to_husl_nd.restype = ctypes.c_void_p       # _simd.h:4-5
        l.rgb_to_husl_nd.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        _REF_LIBS[name] = l
    return _REF_LIBS[name]


def ref_simd(rgb, build="lut", copy=True):
    """Call the reference's rgb_to_husl_nd (nphusl/_simd.c:87) the way its
    Cython wrapper does (_simd_opt.pyx:20-39): uint8 C-contiguous in, malloc'd
    doubles out, copied and freed.  build: lut | lut_strict | lut_interp |
    lut_interp_strict | exact."""
    global _LIBC
    if _LIBC is None:
        _LIBC = ctypes.CDLL(None)
        _LIBC.free.argtypes = [ctypes.c_void_p]
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    ptr = _ref_lib(build).rgb_to_husl_nd(a.ctypes.data, a.size)
    try:
        buf = (ctypes.c_double * a.size).from_address(ptr)
        out = np.frombuffer(buf, dtype=np.float64).reshape(a.shape)
        return out.copy() if copy else None
    finally:
        _LIBC.free(ptr)


def ref_tables(build="lut_strict"):
    """Tables baked into a reference build (generated by its own LUT/light.py and
    LUT/chroma.py) -> same tuple as simd_rgb_to_husl(tables=).  build "lutf_strict":
    the float-table build (`make_lookup_tables float`)."""
    l = _ref_lib(build)
    ctype, dt = (ctypes.c_float, np.float32) if build.startswith("lutf") else (ctypes.c_uint16, np.uint16)
    seg = ctypes.c_ushort.in_dll(l, "L_SEGMENT_SIZE").value
    nh = ctypes.c_ushort.in_dll(l, "CH_TABLE_SIZE").value
    nl = ctypes.c_ushort.in_dll(l, "CL_TABLE_SIZE").value
    lt = np.frombuffer((ctype * (3 * seg)).in_dll(l, "light_table_big"), dtype=dt).copy()
    ct = np.frombuffer((ctype * (nh * nl)).in_dll(l, "chroma_table"), dtype=dt).reshape(nh, nl).copy()
    st = np.array([ctypes.c_double.in_dll(l, "Y_IDX_STEP_%d" % i).value for i in range(3)])
    th = np.array([ctypes.c_double.in_dll(l, "Y_THRESH_%d" % i).value for i in range(2)])
    hs = ctypes.c_double.in_dll(l, "H_IDX_STEP").value
    ls = ctypes.c_double.in_dll(l, "L_IDX_STEP").value
    return lt, st, th, ct, hs, ls


def ref_linear_table():
    l = _ref_lib("lut_strict")
    return np.frombuffer((ctypes.c_double * 256).in_dll(l, "linear_table"), dtype=np.float64).copy()


def ref_package():
    """The reference package itself (nphusl 1.5.0), as compiled by build_ref.sh into
    ``oracle/_ref/refpkg/nphusl_ref``: byte-compiled ``__init__`` / ``nphusl`` / ``transform`` /
    ``constants`` (marshalled code objects behind generated loaders) plus its ``_cython_opt`` and
    ``_simd_opt`` extension modules.  Imported under the
    name ``nphusl_ref`` (all of its imports are relative), so it can sit beside this repository's
    ``nphusl`` in one process; nothing of this repository is on its code path.  The removed
    ``np.float`` / ``np.int`` aliases it still uses are restored first (SURVEY App. B fix 1)."""
    import warnings
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    p = os.path.join(REF_DIR, "refpkg")
    if p not in sys.path:
        sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # "No numexpr" etc.
        import nphusl_ref
    return nphusl_ref


def have_ref_package():
    return os.path.exists(os.path.join(REF_DIR, "refpkg", "nphusl_ref", "nphusl.code"))


def ref_cython():
    """The reference's Cython+OpenMP module (nphusl/_cython_opt.pyx): the only compiled
    HUSL->RGB / hue-only code the reference has."""
    ref_package()
    import nphusl_ref._cython_opt as m
    return m
