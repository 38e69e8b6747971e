"""The NumPy backend (``nphusl/_numpy.py``: what ``enable_numpy`` / ``numpy_enabled`` /
``back_to_std=True`` select, reference ``nphusl/nphusl.py:104-392`` and
``nphusl/__init__.py:53-66,80-88``) against vectors produced by the reference's
own NumPy path (tests/golden/, made by tests/golden/make_golden.py), plus the
switch semantics.  CPU only; the oracle is not involved."""
import warnings

import numpy as np
import pytest

import nphusl
from nphusl import _numpy, transform
from conftest import hue_diff

TOL = 1e-9          # float64 vs float64: only the order of a few roundings differs


def _grey(rgb):
    return (rgb[..., 0] == rgb[..., 1]) & (rgb[..., 1] == rgb[..., 2])


def test_to_husl_matches_reference_numpy(golden):
    rgb, ref = golden["rgb"], golden["hsl"]
    got = _numpy._rgb_to_husl(rgb)
    assert got.dtype == np.float64 and got.shape == rgb.shape
    ng = ~_grey(rgb)
    assert hue_diff(got[ng, 0], ref[ng, 0]).max() < TOL
    assert np.abs(got[..., 1:] - ref[..., 1:]).max() < TOL
    assert hue_diff(_numpy._rgb_to_hue(rgb)[ng], golden["hue"][ng]).max() < TOL


def test_to_rgb_matches_reference_numpy(golden):
    f = _numpy._husl_to_rgb(golden["hsl_in"])
    assert np.abs(f - golden["rgb_float"]).max() < 1e-11
    assert np.array_equal(transform.ensure_rgb_int(f), golden["rgb_u8"])
    # exact round trip, reference tests/test.py:373-377
    back = transform.ensure_rgb_int(_numpy._husl_to_rgb(golden["hsl"]))
    assert np.array_equal(back, golden["rgb"])


def test_float_and_grayscale_inputs(golden):
    got, ref = _numpy._rgb_to_husl(golden["frgb"]), golden["frgb_hsl"]
    sat = ref[:, 1] > 1e-6
    assert hue_diff(got[sat, 0], ref[sat, 0]).max() < 1e-7
    assert np.abs(got[:, 1:] - ref[:, 1:]).max() < TOL
    with nphusl.numpy_enabled(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hsl = nphusl.to_husl(golden["grey"])
        hue = nphusl.to_hue(golden["grey"])
    assert hsl.shape == golden["grey"].shape + (3,) and hue.shape == golden["grey"].shape
    assert np.abs(hsl[..., 1:] - golden["grey_hsl"][..., 1:]).max() < TOL


def test_cube_slices_match_reference_numpy(cube, cube_slices):
    """4094 sampled cube colours and the white / black / near-extreme conventions"""
    flat = cube.reshape(-1, 3)
    idx = cube_slices["sample_idx"]
    got = _numpy._rgb_to_husl(flat[idx])
    ref = cube_slices["sample_hsl"]
    ng = ~_grey(flat[idx])
    assert hue_diff(got[ng, 0], ref[ng, 0]).max() < TOL
    assert np.abs(got[:, 1:] - ref[:, 1:]).max() < TOL
    ends = _numpy._rgb_to_husl(np.array([[255, 255, 255], [0, 0, 0]], np.uint8))
    assert abs(ends[0, 0] - 19.916405993809086) < 1e-6 and ends[0, 1] == 0 and ends[0, 2] == 100
    assert np.all(ends[1] == 0)


def test_step_functions_keep_the_reference_call_shapes():
    """the names tests/test.py and LUT/*.py call (nphusl.nphusl._to_light, _max_lh_chroma, ...)"""
    api = nphusl.nphusl
    rgb = np.random.default_rng(3).random((7, 5, 3))
    xyz = api._rgb_to_xyz(rgb)
    luv = api._xyz_to_luv(xyz)
    lch = api._luv_to_lch(luv)
    assert xyz.shape == luv.shape == lch.shape == rgb.shape
    assert np.allclose(api._rgb_to_lch(rgb), lch)
    assert np.allclose(api._luv_to_xyz(luv), xyz, atol=1e-12)
    assert np.allclose(api._lch_to_luv(lch), luv, atol=1e-9)
    assert np.allclose(api._xyz_to_rgb(xyz), rgb, atol=1e-12)
    assert np.allclose(api._lch_to_rgb(lch), rgb, atol=1e-9)
    assert np.allclose(api._husl_to_lch(api._lch_to_husl(lch)), lch, atol=1e-7)
    assert np.allclose(api._from_light(api._to_light(xyz[..., 1])), xyz[..., 1], atol=1e-14)
    assert np.allclose(api._dot_product(nphusl.constants.M_INV, api._to_linear(rgb)), xyz)
    # six boundary lines per lightness; the shortest non-negative ray is the max chroma (LUT/chroma.py:73-84)
    flat = lch.reshape(-1, 3)
    lines = list(api._bounds(flat[:, 0]))
    assert len(lines) == 6 and all(m.shape == b.shape == (35,) for m, b in lines)
    theta = np.radians(flat[:, 2])
    lens = np.stack([api._ray_length(theta, ln) for ln in lines])
    lens[np.isnan(lens) | (lens < 0)] = np.inf
    assert np.allclose(lens.min(0), api._max_lh_chroma(flat))
    # tests/test.py:559-563 literal: max chroma at L = 0.25, H = 40 degrees
    assert abs(api._max_lh_chroma(np.array([[0.25, 0.0, 40.0]]))[0] - 0.37678582700482516) < 1e-9
    assert api._channel(lch, 2).shape == (7, 5)


def test_enable_numpy_and_context_managers():
    assert set(_numpy.REGISTERED) == set(nphusl.NUMPY)
    nphusl.enable_numpy()
    try:
        for name in nphusl.BACKEND_NAMES:
            assert getattr(nphusl.nphusl, name) is nphusl.NUMPY[name]
            assert getattr(nphusl, name) is nphusl.NUMPY[name]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.allclose(nphusl.to_husl([0.5, 0.2, 0.0, 0.5]), [28.42153418, 100.0, 14.39285828], atol=1e-6)
        assert nphusl.to_rgb([360.0, 100.0, 100.0]).tolist() == [255, 255, 255]
        img = np.random.default_rng(1).integers(0, 256, (24, 31, 3), dtype=np.uint8)
        hsl = nphusl.to_husl(img)
        assert hsl.dtype == np.float64 and np.array_equal(nphusl.to_rgb(hsl), img)
        assert np.array_equal(nphusl.to_hue(img), hsl[..., 0])
        out = np.empty(img.shape)
        assert nphusl.to_husl(img, chunksize=10, out=out) is out and np.allclose(out, hsl)
    finally:
        nphusl.enable_best_optimized()
    best = nphusl.nphusl._rgb_to_husl
    with nphusl.numpy_enabled():
        assert nphusl.nphusl._rgb_to_husl is nphusl.NUMPY["_rgb_to_husl"]
    after = nphusl.nphusl._rgb_to_husl          # without a GPU both are (distinct) "no backend bound" stubs
    assert after is best or (getattr(after, "unbound", False) and getattr(best, "unbound", False))


def test_best_never_selects_numpy_implicitly():
    """enable_best_optimized binds CUDA or nothing: the CPU path only runs when asked for by name"""
    nphusl.enable_best_optimized()
    bound = nphusl.nphusl._rgb_to_husl
    assert bound is not nphusl.NUMPY["_rgb_to_husl"]
    if not nphusl.CUDA:
        assert getattr(bound, "unbound", False)
        with pytest.raises(RuntimeError, match="enable_numpy"):
            nphusl.to_husl(np.zeros((2, 2, 3), np.uint8))


def test_back_to_std_returns_to_numpy(monkeypatch):
    """reference __init__.py:53-60: `with X_enabled(back_to_std=True)` leaves NumPy bound on exit"""
    class Fake:
        @staticmethod
        def _rgb_to_husl(rgb, **kw):
            return np.zeros(np.shape(rgb))
    saved = dict(nphusl.SIMD)
    nphusl.register_backend(nphusl.SIMD, Fake)
    try:
        with nphusl.simd_enabled(back_to_std=True):
            assert nphusl.nphusl._rgb_to_husl is Fake._rgb_to_husl
            # reference __init__.py:63-66: NumPy first, then the overlay -> the other two are NumPy's
            assert nphusl.nphusl._husl_to_rgb is nphusl.NUMPY["_husl_to_rgb"]
            assert nphusl.nphusl._rgb_to_hue is nphusl.NUMPY["_rgb_to_hue"]
        assert nphusl.nphusl._rgb_to_husl is nphusl.NUMPY["_rgb_to_husl"]
    finally:
        nphusl.SIMD.clear()
        nphusl.SIMD.update(saved)
        nphusl.enable_best_optimized()


def test_numpy_backend_is_independent_of_the_oracle_and_cuda():
    import inspect
    src = inspect.getsource(_numpy)
    assert "oracle" not in src.replace("``oracle/``", "") and "_cuda" not in src.replace("``nphusl/_cuda.py``", "")
