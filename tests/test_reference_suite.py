"""The reference's API-level tests (tests/test.py) replayed against the CUDA
backend (``[cuda]``, needs the GPU) and the NumPy backend (``[numpy]``, runs on CPU) --
the counterpart of the reference's ``__with_cuda`` / ``__with_numpy`` test clones
(tests/test.py:33-67) -- with the reference's OWN acceptance criteria (`_diff`, `_diff_husl`:
both HUSL arrays converted back with to_rgb must agree within 2 levels, else
|dH| <= 0.3, |dS| <= 2.0, |dL| <= 0.1; tests/test.py:697-727) and the float64
oracle standing in for the third-party `husl` package those tests import.
Each test names the reference test it replays."""
import warnings

import numpy as np
import pytest

import oracle as orc



@pytest.fixture(scope="module", params=[pytest.param("cuda", marks=pytest.mark.gpu), "numpy"])
def nph(request):
    import nphusl
    if request.param == "cuda":
        from nphusl import _cuda
        assert _cuda.available()
        nphusl.enable_cuda()
    else:
        nphusl.enable_numpy()
    yield nphusl
    nphusl.enable_best_optimized()


def _img(seed=0):
    """stand-in for tests/test.py:733-743: a 39 x 29 photo-like image, row 0
    black, row 1 white, two blocks of uniform noise"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:39, 0:29]
    img = np.stack([(yy * 6) % 256, (xx * 9) % 256, ((yy + xx) * 4) % 256], -1).astype(np.uint8)
    img[12:25, 12:25] = rng.integers(0, 256, (13, 13, 3))
    img[0] = 0
    img[1] = 255
    img[10:20, 10:20] = rng.integers(0, 2, (10, 10, 3))         # reference quirk: near-black block
    return np.ascontiguousarray(img)


def _diff(a, b, diff=0.20):
    assert np.all(np.abs(np.asarray(a, np.float64) - b) <= diff)


def _diff_husl(nph, a, b, max_rgb_diff=2):
    rgb_a = nph.to_rgb(a).astype(int)
    rgb_b = nph.to_rgb(b).astype(int)
    if np.all(np.abs(rgb_a - rgb_b) <= max_rgb_diff):
        return
    a, b = np.asarray(a), np.asarray(b)
    _diff(a[..., 0], b[..., 0], 0.3)
    _diff(a[..., 1], b[..., 1], 2.0)
    _diff(a[..., 2], b[..., 2], 0.1)


def test_to_husl_2d_and_3d(nph):                         # test_to_husl_2d / test_to_husl_3d (:70-101)
    img = _img()
    for arr in (img[:, 4], img):
        _diff_husl(nph, nph.to_husl(arr), orc.rgb_to_husl(arr))


def test_to_husl_gray_3d(nph):                           # test_to_husl_gray_3d (:104-121)
    img = _img()
    img[..., 1] = img[..., 0]
    img[..., 2] = img[..., 0]
    _diff_husl(nph, nph.to_husl(img), orc.rgb_to_husl(img))


def test_to_hue_vs_old_and_gray(nph):                    # test_to_hue_vs_old / test_to_hue_gray (:124-148)
    img = _img()
    ref = orc.rgb_to_husl(img)
    hue = nph.to_hue(img)
    tol = np.where(ref[..., 1] < 1, 5.0, 0.1)
    d = np.abs(hue - ref[..., 0]) % 360
    assert np.all(np.minimum(d, 360 - d) <= tol)
    grey = img[..., 0] / 255.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hue_g = nph.to_hue(grey)
    assert hue_g.shape == grey.shape
    assert np.all(orc.rgb_to_husl(np.repeat(grey[..., None], 3, -1))[..., 1] < 1)


def test_to_rgb_2d_3d_round_trip(nph):                   # test_to_rgb_2d / test_to_rgb_3d (:369-390): exact for 3d
    img = _img()
    assert np.all(nph.to_rgb(nph.to_husl(img)) == img)
    assert np.all(nph.to_rgb(nph.to_husl(img[:, 14])) == img[:, 14])


def test_husl_to_rgb_from_reference_husl(nph):           # test_husl_to_rgb (:393-402): within 1
    img = _img()
    back = nph.to_rgb(orc.rgb_to_husl(img))
    assert np.all(np.abs(back.astype(int) - img) <= 1)


def test_to_hue_2d_3d_equals_husl_hue(nph):              # test_to_hue_2d / test_to_hue_3d (:517-530)
    img = _img()
    _diff(nph.to_husl(img[:, 14])[..., 0], nph.to_hue(img[:, 14]))
    _diff(nph.to_husl(img)[..., 0], nph.to_hue(img))


def test_transform_rgb_in_chunks(nph):                   # test_transform_rgb (:509-514)
    from nphusl import transform
    img = _img()
    _diff(nph.to_husl(img), transform.in_chunks(img, nph.to_husl, 10))


def test_to_husl_rgba(nph):                              # test_to_husl_rgba (:547-556)
    from nphusl import transform
    rgb = transform.ensure_rgb_float(_img())
    rgba = np.zeros(rgb.shape[:-1] + (4,), rgb.dtype)
    rgba[..., :3] = rgb
    rgba[..., 3] = 0.5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _diff_husl(nph, nph.to_husl(rgba), nph.to_husl(rgb * 0.5))


def test_triplets_and_quadruplets(nph):                  # test_to_husl_triplet / _quadruplet / _1d_grayscale (:655-684)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert isinstance(nph.to_husl([10, 30, 40]), np.ndarray)
        _diff(nph.to_husl([0x80, 0x80, 0x80])[1:], nph.to_husl([0.5, 0.5, 0.5])[1:], diff=0.3)
        assert nph.to_husl([30, 30, 30])[1] < 0.01
        assert nph.to_husl([255, 0, 0])[0] < 13
        assert nph.to_rgb([360.0, 100.0, 100.0]).tolist() == [255, 255, 255]
        assert np.allclose(nph.to_husl([0.5, 0.2, 0.0, 0.5]), [28.42153418, 100.0, 14.39285828], atol=1e-3)
        a = nph.to_husl(np.array([0.1, 0.1]))
        b = nph.to_husl(np.array([[.1, .1, .1]] * 2))
        assert a.shape == b.shape == (2, 3)
        _diff_husl(nph, a, b)
        # the literal in tests/test.py:681-684 is the uint8-quantised answer (0.1 -> 26/255)
        _diff_husl(nph, a, np.array([[288.467242, 1.02865661e-05, 9.30666400]] * 2))


def test_max_chroma_literal():                           # test_cython_max_chroma (:559-563), via the oracle
    assert abs(orc.max_chroma(0.25, 40.0) - 0.37678582700482516) < 1e-3
