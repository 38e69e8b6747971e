"""The C-ABI shared library on a CPU-only box: it loads, exports every symbol
include/nphusl_cuda.h declares, the ctypes signature table covers the header,
and -- with no GPU -- entry points fail with a status instead of computing."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from nphusl import _cuda

HEADER = os.path.join(ROOT, "include", "nphusl_cuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nphusl_cuda_\w+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _cuda.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert lib.nphusl_cuda_abi_version() == 1


def test_ctypes_table_matches_header():
    assert sorted(_cuda.SIGNATURES) == declared_symbols()


def test_header_cites_the_reference_interfaces():
    text = open(HEADER).read()
    for cite in ("_simd.h:4-5", "_simd.c:87", "_cython_opt.pyx:22-93", "_cython_opt.pyx:96-156",
                 "_cython_opt.pyx:177-246", "transform.py:16-20"):
        assert cite in text


def test_no_torch_or_numpy_types_in_the_abi():
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)       # declarations only
    for banned in ("torch", "Tensor", "PyObject", "at::", "numpy", "#include <cuda"):
        assert banned not in code


def test_status_not_compute_without_gpu():
    lib = _cuda.load()
    if lib.nphusl_cuda_device_count() > 0:
        pytest.skip("GPU present")
    rgb = np.zeros((4, 3), np.uint8)
    out = np.full((4, 3), -1.0)
    rc = lib.nphusl_cuda_rgb_to_husl(rgb.ctypes.data, 0, 0, 3, out.ctypes.data, 2, 0, 4, 0, 0, None)
    assert rc == 3                                            # NPHUSL_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.nphusl_cuda_last_error()
    assert np.all(out == -1.0)                                # nothing was computed on the CPU
    assert lib.nphusl_cuda_rgb_to_husl(None, 0, 0, 3, None, 2, 0, 4, 0, 0, None) == 1    # NPHUSL_ERR_ARG
    assert lib.nphusl_cuda_rgb_to_husl(rgb.ctypes.data, 7, 0, 3, out.ctypes.data, 2, 0, 4, 0, 0, None) == 1
    assert lib.nphusl_cuda_husl_to_rgb(out.ctypes.data, 0, 0, rgb.ctypes.data, 0, 0, 4, 0, None) == 1
    assert lib.nphusl_cuda_rgb_to_husl(rgb.ctypes.data, 0, 0, 3, out.ctypes.data, 2, 0, 0, 0, 0, None) == 0  # n = 0
    p = ctypes.c_void_p()
    assert lib.nphusl_cuda_malloc(ctypes.byref(p), 16, 0) == 3
    assert lib.nphusl_cuda_lut_generate(0, 1024, 1024, 1024) == 3


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "husl-numpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "gen_coeffs.py":
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "husl_oracle" not in src and "libemul" not in src, f


def _client():
    import subprocess
    d = os.path.join(ROOT, "tests", "cabi")
    subprocess.run(["make", "-C", d, "-s"], check=True)
    return os.path.join(d, "cabi_client")


def test_plain_c_client_links_against_the_header_and_library():
    """A C99 program that includes nothing but include/nphusl_cuda.h links and runs (no CUDA
    headers, no Python): the boundary really is a C ABI."""
    import subprocess
    out = subprocess.run([_client()], check=True, capture_output=True, text=True).stdout
    assert out.startswith("abi 1 devices ")


@pytest.mark.gpu
def test_plain_c_client_round_trip_on_the_gpu():
    import subprocess
    r = subprocess.run([_client(), "roundtrip"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "roundtrip ok" in r.stdout
