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
