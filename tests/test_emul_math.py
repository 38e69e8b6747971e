"""float32 FORMULATION check on a CPU-only box: csrc/husl_math.cuh compiled for
the host (tests/emul, MUFU ops replaced by float32 libm + a random +-3 ulp
perturbation) is swept over all 2^24 colours against the float64 oracle.

This does not test the CUDA kernels (tests/test_cuda_parity.py does, on the
GPU); it guards the algebra -- cancellation-free hue / saturation, the
one-reciprocal inverse, the exact uint8 round trip -- so a change to the math
can be rejected before GPU time is spent."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle as orc
from conftest import ROOT, hue_diff

EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not installed")
    subprocess.run(["make", "-C", EMUL, "-s"], check=True)
    return ctypes.CDLL(os.path.join(EMUL, "libemul.so"))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("noise", [0, 3])
def test_cube_forward_and_round_trip(emul, cube, noise):
    emul.emul_set_noise(noise)
    flat = cube.reshape(-1, 3)
    n = flat.shape[0]
    got = np.empty((n, 3), np.float32)
    emul.emul_fwd_u8(_p(flat), _p(got), ctypes.c_longlong(n))
    ref = orc.rgb_to_husl(flat)
    grey = (flat[:, 0] == flat[:, 1]) & (flat[:, 1] == flat[:, 2])
    assert hue_diff(got[:, 0], ref[:, 0])[~grey].max() < 2e-4        # bar: 1e-3
    assert np.abs(got[:, 1] - ref[:, 1]).max() < 2e-4
    assert np.abs(got[:, 2] - ref[:, 2]).max() < 2e-4
    u8 = np.empty((n, 3), np.uint8)
    f = np.empty((n, 3), np.float32)
    emul.emul_inv(_p(got), _p(f), _p(u8), ctypes.c_longlong(n))
    assert np.array_equal(u8, flat)
    assert np.abs(f - flat).max() < 0.1                               # margin to the .5 rounding boundary
    emul.emul_set_noise(0)


def test_float_inputs_and_clamps(emul, golden):
    frgb = np.ascontiguousarray(golden["frgb"])
    got = np.empty(frgb.shape, np.float32)
    emul.emul_fwd_f64(_p(frgb), _p(got), ctypes.c_longlong(frgb.shape[0]))
    ref = golden["frgb_hsl"]
    sat = ref[:, 1] > 0.1
    assert hue_diff(got[sat, 0], ref[sat, 0]).max() < 1e-3
    assert np.abs(got[:, 1:] - ref[:, 1:]).max() < 1e-3


def test_inverse_on_arbitrary_husl(emul, golden):
    hsl = golden["hsl_in"].astype(np.float32)
    n = hsl.shape[0]
    f = np.empty((n, 3), np.float32)
    u8 = np.empty((n, 3), np.uint8)
    emul.emul_inv(_p(hsl), _p(f), _p(u8), ctypes.c_longlong(n))
    assert np.abs(f / 255.0 - golden["rgb_float"]).max() < 2e-4
    assert np.abs(u8.astype(int) - golden["rgb_u8"].astype(int)).max() <= 1


def _ulps(x, k):
    return (x.view(np.int32) + k.astype(np.int32)).view(np.float32)


@pytest.mark.parametrize("noise", [0, 3])
def test_float32_channels_without_float64(emul, noise):
    """diff_from_f32 (husl_math.cuh): float32 RGB channels -> the four quantities the forward core is built on,
    formed without cancellation and without float64.  Swept where a float32-only formulation could fail --
    channels a few ulps apart, both sides of the 0.04045 segment switch, near 0 and near 1 -- against the
    float64 oracle (bar 1e-3; hue where it is defined, S > 0.1, as for every float-input check) and, for the
    intermediate differences themselves, against float64 linearisation (relative error)."""
    rng = np.random.default_rng(1)
    n = 200000
    g = rng.random(n, dtype=np.float32) * np.float32(0.98) + np.float32(0.01)
    sets = {"uniform": rng.random((n, 3), dtype=np.float32)}
    for name, kmax in (("2ulp", 2), ("100ulp", 100), ("1e4ulp", 10 ** 4), ("1e6ulp", 10 ** 6)):
        a = np.stack([_ulps(g, rng.integers(-kmax, kmax + 1, n)), g, _ulps(g, rng.integers(-kmax, kmax + 1, n))], 1)
        sets[name] = np.clip(a, 0, 1).astype(np.float32)
    t = np.float32(0.04045)
    sets["switch"] = (t + (rng.random((n, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(0.02)).astype(np.float32)
    sets["switch_tight"] = _ulps(np.full((n, 3), t, np.float32), rng.integers(-50, 51, (n, 3)))
    sets["dark"] = rng.random((n, 3), dtype=np.float32) * np.float32(0.05)
    sets["bright"] = np.float32(1) - rng.random((n, 3), dtype=np.float32) * np.float32(0.02)
    sets["quant8"] = (rng.integers(0, 256, (n, 3)) / 255).astype(np.float32)
    emul.emul_set_noise(noise)
    try:
        for name, a in sets.items():
            a = np.ascontiguousarray(a)
            got = np.empty(a.shape, np.float32)
            emul.emul_fwd_f32(_p(a), _p(got), ctypes.c_longlong(len(a)))
            ref = orc.rgb_to_husl(a.astype(np.float64))
            # away from the L > 99.99 / L < 0.01 clamps, where a float32 Y and a float64 Y may fall on different sides
            mid = (ref[:, 2] > 0.02) & (ref[:, 2] < 99.98)
            sat = mid & (ref[:, 1] > 0.1)
            if sat.any():
                assert hue_diff(got[sat, 0], ref[sat, 0]).max() < 5e-4, name
            assert np.abs(got[mid, 1:] - ref[mid, 1:]).max() < 5e-4, name
            w = np.zeros(4)
            emul.emul_diff_f32_err(_p(a), ctypes.c_longlong(len(a)), _p(w))
            assert w.max() < 1e-5, (name, w)          # 86 degrees per unit relative error of (dr, db)
    finally:
        emul.emul_set_noise(0)
    # NaN channels stay NaN, equal channels give exactly zero differences
    a = np.array([[np.nan, 0.5, 0.5], [0.5, np.nan, 0.5], [0.3, 0.3, 0.3], [0.01, 0.01, 0.01]], np.float32)
    got = np.empty(a.shape, np.float32)
    emul.emul_fwd_f32(_p(a), _p(got), ctypes.c_longlong(len(a)))
    assert np.isnan(got[0]).any() and np.isnan(got[1]).any()
    assert np.all(np.abs(got[2:, 0] - 19.916405993809086) < 1e-3) and np.all(got[2:, 1] < 1e-3)


# ---- the HUSL-space stages (csrc/husl_edit.cuh) on the CPU --------------------------------------------

class _Bands(ctypes.Structure):
    _fields_ = [("src", ctypes.c_int), ("width", ctypes.c_float), ("total", ctypes.c_float), ("h_even", ctypes.c_float),
                ("h_odd", ctypes.c_float), ("s_half", ctypes.c_float), ("s_value", ctypes.c_float)]


class _Edit(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("h_lo", "h_hi", "s_lo", "s_hi", "l_lo", "l_hi")] + [("invert", ctypes.c_int)] + \
               [(n, ctypes.c_float) for n in ("h_mul", "h_add", "s_mul", "s_add", "l_mul", "l_add")]


def test_banded_map_closed_form_equals_the_callers_loop(emul):
    """apply_bands evaluates the callers' `for low, high in chunk(total, width)` loops (examples.py:56-103) without a loop:
    x * (1 / width) proposes a band, the float32 edges decide, the stripe test looks at three band centres.  Swept against
    the loop itself (tests/specs.py) over every band edge +- a few ulps, values below 0 and beyond `total`, NaN and
    infinities, for integer and fractional widths, widths that do and do not divide the range, stripes wider than a band."""
    from specs import bands_spec
    from nphusl import chunk
    rng = np.random.default_rng(3)
    cases = [(0, 45, 360, None), (0, 45, 360, (3, 90)), (2, 7, 100, (2, 100)), (2, 1, 100, (2, 100)), (2, 3, 100, (2, 100)),
             (2, 100, 100, (2, 100)), (2, 250, 100, (2, 100)), (1, 12.5, 100, (1.5, 40)), (0, 0.1, 360, (0.01, 5)),
             (2, 33, 100, (20, 50)), (0, 359.99, 360, (1, 1)), (1, 1e-3, 1, (1e-4, 2)), (0, 7.3, 355.5, (0.7, 9))]
    for src, width, total, stripe in cases:
        nb = len(list(chunk(int(total), int(width)))) if float(width).is_integer() and float(total).is_integer() else None
        b = _Bands(src, width, total, 130, 360, stripe[0] if stripe else 0, stripe[1] if stripe else 0)
        last = emul.emul_bands_last(ctypes.byref(b))
        if nb is not None:
            assert last == nb - 1, (width, total, last, nb)            # the band count of the callers' chunk()
        edges = np.arange(0, last + 3, dtype=np.float32) * np.float32(width)
        x = np.concatenate([_ulps(np.repeat(edges, 9), np.tile(np.arange(-4, 5), edges.size)),
                            _ulps(np.full(9, total, np.float32), np.arange(-4, 5)),
                            rng.uniform(-0.2 * total, 1.3 * total, 20000).astype(np.float32),
                            np.array([0.0, -0.0, np.nan, np.inf, -np.inf, -1e30, 1e30], np.float32)])
        hsl = np.stack([rng.uniform(0, 360, x.size), rng.uniform(0, 100, x.size), rng.uniform(0, 100, x.size)], 1).astype(np.float32)
        hsl[:, src] = x
        want = bands_spec(hsl, src, width, total, 130, 360, stripe)
        got = hsl.copy()
        emul.emul_apply_bands(_p(got), ctypes.c_longlong(len(got)), ctypes.byref(b), 1 if stripe else 0)
        bad = np.flatnonzero(~np.all((got == want) | (np.isnan(got) & np.isnan(want)), axis=1))
        assert bad.size == 0, (src, width, total, stripe, x[bad][:6], got[bad][:3], want[bad][:3])
        if not stripe:                                                  # the <true> instantiation is valid for s_half <= 0 too
            got2 = hsl.copy()
            emul.emul_apply_bands(_p(got2), ctypes.c_longlong(len(got2)), ctypes.byref(b), 1)
            assert np.array_equal(got2, got, equal_nan=True)


def test_edit_stage_equals_the_numpy_statement(emul):
    """apply_edit<true, true, true> against tests/specs.py: hue arcs through 0, bounded and unbounded ranges, inversion, every
    combination of edited channels; untouched channels and unselected pixels bit for bit."""
    from specs import edit_spec
    rng = np.random.default_rng(4)
    n = 50000
    hsl = np.stack([rng.uniform(0, 360, n), rng.uniform(0, 100, n), rng.uniform(0, 100, n)], 1).astype(np.float32)
    hsl[:64, 0] = np.repeat(np.array([0, 250, 290, 360], np.float32), 16)
    inf = float("inf")
    for kw in (dict(hue=(250, 290), invert=True, l=(0.5, 0)), dict(lightness=(0, 62), l=(0, 0)), dict(saturation=(-inf, 80), l=(0, 0)),
               dict(hue=(300, 40), saturation=(10, 95), lightness=(5, 90), h=(1.0, 77.5), s=(1.5, -5), l=(0.8, 3)),
               dict(hue=(10, 200), h=(-1.0, 0.0)), dict(h=(2.0, 100.0), s=(0.0, 100.0)), dict(invert=True, l=(2.0, 0))):
        e = _Edit(*(kw.get("hue", (-inf, inf)) + kw.get("saturation", (-inf, inf)) + kw.get("lightness", (-inf, inf))),
                  int(kw.get("invert", False)), *(kw.get("h", (1.0, 0.0)) + kw.get("s", (1.0, 0.0)) + kw.get("l", (1.0, 0.0))))
        got = hsl.copy()
        emul.emul_apply_edit(_p(got), ctypes.c_longlong(n), ctypes.byref(e))
        assert np.array_equal(got, edit_spec(hsl, **kw)), kw
