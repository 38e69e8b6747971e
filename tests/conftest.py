"""Test configuration: puts the host package (``husl-numpy_b200/nphusl``) and the
CPU oracle (``oracle/``, test infrastructure only) on ``sys.path`` and registers
the ``gpu`` marker.  ``-m "not gpu"`` runs on a CPU-only box; ``-m gpu`` needs a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(ROOT, "husl-numpy_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


# options of tests/performance_test.py, the reference's own (conftest.py:5-13) with "cuda" in front
DEFAULT_IMPLS = "cuda numpy".split()


def pytest_addoption(parser):
    parser.addoption("--impls", action="store", nargs="*", default=DEFAULT_IMPLS,
                     help="backends to time: cuda numpy reference (reference = the reference package's own best backends on the CPU)")
    parser.addoption("--iters", action="store", type=int, default=3, help="best of this many timeit runs")
    parser.addoption("--img", action="store", default="small",
                     help="small (270x480) | 1080p | 4m (2000x2000) | 4k | cube (4096x4096, all 2^24 colours) | path to a .npy uint8 image")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "husl_golden.npz")))


@pytest.fixture(scope="session")
def simd_golden():
    return dict(np.load(os.path.join(GOLDEN, "simd_golden.npz")))


@pytest.fixture(scope="session")
def simd_float_golden():
    """the reference built with FLOAT lookup tables (`make_lookup_tables float 1024 1024`)"""
    return dict(np.load(os.path.join(GOLDEN, "simd_float_golden.npz")))


@pytest.fixture(scope="session")
def cube_slices():
    return dict(np.load(os.path.join(GOLDEN, "cube_numpy_slices.npz")))


@pytest.fixture(scope="session")
def golden_meta():
    import json
    with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cube():
    """all 2^24 colours, r slowest / b fastest (reference tests/make_rgb16mil.py:8-17)"""
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    return np.ascontiguousarray(np.stack([r, g, b], axis=-1).reshape(4096, 4096, 3))


def hue_diff(a, b):
    """circular distance in degrees"""
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) % 360.0
    return np.minimum(d, 360.0 - d)
