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
