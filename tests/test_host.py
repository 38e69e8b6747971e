"""Host side of the drop-in boundary on a CPU-only box: dtype / shape guards of
``nphusl.transform`` (semantics of reference transform.py:16-308, exercised the
way reference tests/test.py:484-652 does), the backend registry and switches
(reference nphusl/__init__.py:45-88) driven with a recording stand-in backend,
and the no-CPU-fallback behaviour."""
import warnings

import numpy as np
import pytest

import nphusl
from nphusl import transform


# ---- dtype coercion ---------------------------------------------------------------

def test_float_input_keeps_floats_converts_ints():
    seen = []

    @transform.float_input
    def go(a):
        seen.append(a.dtype)
    go(np.arange(6, dtype=np.float32))
    go(np.arange(6, dtype=np.float64))
    go(np.arange(6, dtype=np.uint8))
    assert seen == [np.float32, np.float64, np.float64]


def test_int_input():
    seen = []

    @transform.int_input
    def go(a):
        seen.append(a.dtype)
    go(np.arange(6, dtype=np.int64))
    go(np.arange(6, dtype=np.uint8))
    go(np.arange(6, dtype=np.float32))
    assert seen == [np.int64, np.uint8, np.int64]


def test_rgb_float_input_scales_integers_down():
    @transform.rgb_float_input
    def go(a):
        return a
    assert np.array_equal(go(np.arange(10, dtype=np.int64)), np.arange(10) / 255)
    f32 = go(np.arange(10, dtype=np.float32))
    assert f32.dtype == np.float64 and np.array_equal(f32, np.arange(10))
    same = np.arange(10, dtype=np.float64)
    assert go(same) is same


def test_rgb_int_input_scales_floats_up_half_even():
    @transform.rgb_int_input
    def go(a):
        return a
    got = go(np.array([0.1, 0.15, 0.66, 0.33]))
    assert got.dtype == np.uint8 and got.tolist() == [26, 38, 168, 84]
    assert go(np.array([0.5 / 255, 1.5 / 255, 2.5 / 255])).tolist() == [0, 2, 2]     # np.round: half to even
    assert go(np.arange(10, dtype=np.int32)).dtype == np.uint8
    same = np.arange(0, 255, dtype=np.uint8)
    assert go(same) is same
    # no clamp: negative mirrored, >255 wraps (reference transform.py:16-20)
    assert go(np.array([-0.1, 1.01])).tolist() == [26, 258 % 256]


# ---- shape guards --------------------------------------------------------------------

def test_reshape_image_input_cases():
    @transform.reshape_image_input
    def go(a):
        return a
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert isinstance(go([1, 2, 3]), np.ndarray) and go([1, 2, 3]).shape == (1, 3)
        with pytest.raises(ValueError):
            go(list(range(10)))
        with pytest.raises(ValueError):
            go(np.arange(10))
        assert go(np.arange(9)).shape == (3, 3)
        ok = np.arange(9).reshape(3, 3)
        assert go(ok) is ok
        assert go([1, 2, 3, 4]).shape == (1, 4)
        assert go(list(range(8))).shape == (2, 4)
        g = go(np.linspace(0, 1, 10))                 # 1-D float, not a pixel count: grayscale
        assert g.shape == (10, 3) and np.array_equal(g[:, 0], g[:, 2])
        g2 = go(np.random.default_rng(0).random((5, 7)))
        assert g2.shape == (5, 7, 3) and g2.strides[-1] == 0      # zero-copy broadcast view
        with pytest.raises(ValueError):
            go(np.zeros((5, 7), np.uint8))
        with pytest.raises(ValueError):
            go(np.zeros((5, 7, 2)))


def test_reshape_alert_levels():
    @transform.reshape_image_input
    def go(a):
        return a
    old = dict(transform.alert_levels)
    try:
        transform.alert_levels[transform.ReshapeAlert.flat] = transform.Alert.error
        with pytest.raises(ValueError):
            go(np.arange(9))
        transform.alert_levels[transform.ReshapeAlert.flat] = transform.Alert.warn
        with pytest.warns(UserWarning):
            go(np.arange(9))
        transform.alert_levels[transform.ReshapeAlert.flat] = transform.Alert.quiet
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            go(np.arange(9))
    finally:
        transform.alert_levels.update(old)


def test_reshape_husl_input():
    @transform.reshape_husl_input
    def go(a):
        return a
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert go([1.0, 2.0, 3.0]).shape == (1, 3)
        assert go(np.zeros(12)).shape == (4, 3)
        with pytest.raises(ValueError):
            go(np.zeros(10))
        with pytest.raises(ValueError):
            go(np.zeros((4, 4)))


def test_rgba_premultiply():
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (6, 5, 3), dtype=np.uint8)
    rgba = np.concatenate([rgb, np.full((6, 5, 1), 0x80, np.uint8)], -1)
    go = transform.reshape_rgba_input(lambda a: a)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = go(rgba)
        assert out.shape == rgb.shape
        # integer RGBA -> float RGB in [0, 1], c/255 * a/255: what the CUDA kernels compute for uint8 RGBA device
        # arrays.  (The reference's tests/test.py:533-544 pins its accidental float-0..255 result, App. D-5.)
        assert out.dtype == np.float64 and np.allclose(out, rgb / 255.0 * (0x80 / 255.0), atol=1e-15)
        f = go(np.array([[0.5, 0.2, 0.0, 0.5]]))
        assert np.allclose(f, [[0.25, 0.1, 0.0]])


def test_squeeze_output():
    @transform.squeeze_output
    def go():
        return np.asarray([[1, 2, 3]])
    assert go().shape == (3,) and list(go()) == [1, 2, 3]


# ---- chunking ---------------------------------------------------------------------------

def test_chunk_bounds():
    assert list(transform.chunk(10, 3)) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert list(transform.chunk(9, 3)) == [(0, 3), (3, 6), (6, 9)]
    assert list(transform.chunk(2, 3)) == [(0, 2)]
    assert list(transform.chunk(5, None)) == [(0, 5)]


def test_chunk_img_tiles_are_views_covering_the_image():
    img = np.ones((25, 33, 3), np.int64)
    for i, (tile, _) in enumerate(transform.chunk_img(img, 10)):
        tile[:] = i
    seen = np.zeros(img.shape[:2], bool)
    for i, (tile, ((r0, r1), (c0, c1))) in enumerate(transform.chunk_img(img, 10)):
        assert np.all(tile == i) and tile.shape[0] <= 10 and tile.shape[1] <= 10
        assert not seen[r0:r1, c0:c1].any()
        seen[r0:r1, c0:c1] = True
    assert seen.all() and np.all(img[:10, :10] == 0)
    assert nphusl.chunk_img is transform.chunk_img and nphusl.chunk is transform.chunk   # README.md:116,164


def test_in_chunks_matches_whole_and_accepts_both_argument_orders():
    img = np.random.default_rng(2).random((23, 17, 3))
    fn = lambda a: a * 2.0 + 1.0           # noqa: E731
    whole = transform.in_chunks(img, fn)
    assert np.array_equal(transform.in_chunks(img, fn, 10), whole)
    out = np.empty_like(img)
    assert transform.in_chunks(img, fn, 10, out) is out and np.array_equal(out, whole)
    out2 = np.empty_like(img)
    transform.in_chunks(img, fn, out2, 10)          # the reference's declared order
    assert np.array_equal(out2, whole)
    out3 = np.empty_like(img)
    transform.in_chunks(img, fn, None, out3)
    assert np.array_equal(out3, whole)
    hue = transform.in_chunks(img, lambda a: a[..., 0], 8)
    assert hue.shape == img.shape[:2] and np.array_equal(hue, img[..., 0])


def test_chunk_apply_in_place():
    img = np.random.default_rng(3).integers(0, 3, (20, 20, 3))
    zero_at = img == 0

    def bump(tile):
        tile = tile.copy()
        tile[tile == 0] = 100
        return tile
    transform.chunk_apply(bump, transform.chunk_img(img, 7), out=img)
    assert not (img == 0).any() and np.all(img[zero_at] == 100)


# ---- registry / switches --------------------------------------------------------------------

class Recorder:
    """stand-in backend module: records what reaches it"""

    def __init__(self, tag):
        self.tag, self.calls = tag, []

    def _rgb_to_husl(self, rgb, **kw):
        self.calls.append(("husl", rgb.shape, rgb.dtype, kw))
        return np.full(rgb.shape, 1.0)

    def _rgb_to_hue(self, rgb, **kw):
        self.calls.append(("hue", rgb.shape, rgb.dtype, kw))
        return np.full(rgb.shape[:-1], 2.0)

    def _husl_to_rgb(self, hsl, **kw):
        self.calls.append(("rgb", hsl.shape, hsl.dtype, kw))
        return np.full(hsl.shape, 0.5)


@pytest.fixture
def fake_simd():
    rec = Recorder("simd")
    saved = dict(nphusl.SIMD)
    nphusl.register_backend(nphusl.SIMD, rec)
    try:
        yield rec
    finally:
        nphusl.SIMD.clear()
        nphusl.SIMD.update(saved)
        nphusl.enable_best_optimized()


def test_switch_names_exist():
    for name in ("cuda", "simd", "cython", "numexpr", "numpy"):
        assert callable(getattr(nphusl, "enable_" + name))
        assert callable(getattr(nphusl, name + "_enabled"))
    assert callable(nphusl.best_enabled) and callable(nphusl.enable_best_optimized)
    for d in ("CUDA", "SIMD", "CYTHON", "NUMEXPR", "NUMPY"):
        assert isinstance(getattr(nphusl, d), dict)


def test_absent_backends_raise_like_the_reference():
    assert nphusl.NUMPY, "the NumPy backend ships with the package"
    for name in ("simd", "cython", "numexpr"):
        if getattr(nphusl, name.upper()):
            continue
        with pytest.raises(AssertionError, match="implementation not available"):
            getattr(nphusl, "enable_" + name)()
        with pytest.raises(AssertionError):
            with getattr(nphusl, name + "_enabled")():
                pass


def test_dispatch_through_public_api(fake_simd):
    with nphusl.simd_enabled():
        assert nphusl.nphusl._rgb_to_husl == fake_simd._rgb_to_husl
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h = nphusl.to_husl([10, 30, 40])
            assert h.shape == (3,) and fake_simd.calls[-1][:2] == ("husl", (1, 3))
            nphusl.to_hue(np.zeros((4, 5, 3), np.uint8))
            assert fake_simd.calls[-1][:2] == ("hue", (4, 5, 3))
            rgb = nphusl.to_rgb(np.zeros((4, 5, 3)))
            assert rgb.dtype == np.uint8 and np.all(rgb == 128)          # rgb_int_output: round(0.5*255)
            nphusl.to_husl(np.zeros((2, 2, 4)))                          # RGBA -> premultiplied RGB
            assert fake_simd.calls[-1][1] == (2, 2, 3)
            nphusl.to_husl(np.zeros((6, 7)))                             # grayscale floats
            assert fake_simd.calls[-1][1] == (6, 7, 3)
            nphusl.to_husl(np.zeros((8, 8, 3), np.uint8), dtype=np.float32)
            assert fake_simd.calls[-1][3] == {"dtype": np.float32}       # backend extras pass through
    # leaving the context reverts to the best available: CUDA if present, else the next in line
    assert nphusl.nphusl._rgb_to_husl == (nphusl.CUDA or nphusl.SIMD)["_rgb_to_husl"]


def test_cuda_is_preferred_over_other_backends(fake_simd):
    rec = Recorder("cuda")
    saved = dict(nphusl.CUDA)
    nphusl.register_backend(nphusl.CUDA, rec)
    try:
        nphusl.enable_best_optimized()
        assert nphusl.nphusl._rgb_to_husl == rec._rgb_to_husl
        nphusl.enable_simd()
        assert nphusl.nphusl._rgb_to_husl == fake_simd._rgb_to_husl
    finally:
        nphusl.CUDA.clear()
        nphusl.CUDA.update(saved)
        nphusl.enable_best_optimized()


def test_no_cpu_fallback_without_gpu():
    from nphusl import _cuda
    if _cuda.available():
        pytest.skip("a GPU is visible: the CUDA backend is bound")
    assert nphusl.CUDA == {}
    with pytest.raises(AssertionError, match="implementation not available"):
        nphusl.enable_cuda()
    with pytest.raises(RuntimeError, match="no backend"):
        nphusl.to_husl(np.zeros((2, 2, 3), np.uint8))
    with pytest.raises(_cuda.CudaUnavailable):
        _cuda._rgb_to_husl(np.zeros((2, 3), np.uint8))
    with pytest.raises(_cuda.CudaUnavailable):
        _cuda._husl_to_rgb(np.zeros((2, 3)))
    with pytest.raises(_cuda.CudaUnavailable):
        _cuda._rgb_to_hue(np.zeros((2, 3), np.uint8))


def test_device_arrays_pass_the_guards_untouched():
    class FakeDev:
        def __init__(self, shape, typestr):
            self.shape = shape
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (256, False),
                                             "version": 3, "strides": None}
    d = FakeDev((4, 5, 3), "|u1")
    assert transform.is_device_array(d)
    assert transform.ensure_rgb_float(d) is d and transform.ensure_rgb_int(d) is d
    assert transform.reshape_image_input(lambda a: a)(d) is d
    assert transform.reshape_husl_input(lambda a: a)(FakeDev((7, 3), "<f4")).shape == (7, 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = transform.reshape_image_input(lambda a: a)(FakeDev((4, 5), "<f8"))
    assert g.shape == (4, 5, 3) and g.__cuda_array_interface__["strides"] == (40, 8, 0)
    with pytest.raises(ValueError):
        transform.in_chunks(d, lambda a: a, 10)
