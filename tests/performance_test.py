"""Best-of-N timing table per backend, through the PUBLIC API -- the method of the reference's
tests/performance_test.py:45-103 (``timeit.repeat("fn(img)", repeat=iters, number=1)`` inside each
``*_enabled()`` context, minimum taken, pixels/s = ``img.size / 3 / best``) with ``cuda`` among the
implementations (reference conftest.py:5, performance_test.py:14 list ``simd cython numexpr numpy``).

    pytest tests/performance_test.py -s --impls cuda numpy reference --iters 5 --img 1080p
    pytest tests/performance_test.py -s --impls cuda --iters 10 --img 4k          # also: 4m, cube, small, file.npy

Rows: ``cuda`` with the image in host memory (what a drop-in caller of the reference passes) and already
on the device (``cuda/device``: a torch CUDA tensor in, a device array out); ``numpy`` (this package's
NumPy backend); ``reference`` = the reference package itself (oracle/_ref/refpkg, built from
/root/reference) with its own best backends, C + OpenMP for to_husl and Cython + OpenMP for to_rgb /
to_hue, on the host cores -- printed beside the others, never part of the product.

The synthetic images are the reference's generators, seeded (tests/make_rand1080p.py:4-5,
make_rand4million.py:4-5, make_rgb16mil.py:5-17).  Like the reference's file this prints tables and
asserts only that every backend returns the documented dtype and shape.
"""
import timeit
import warnings

import numpy as np
import pytest

import nphusl

SIZES = {"small": (270, 480), "1080p": (1080, 1920), "4m": (2000, 2000), "4k": (2160, 3840)}
KNOWN = ("cuda", "numpy", "reference")


def _impl_params(config):
    impls = config.getoption("--impls")
    assert all(m in KNOWN for m in impls), "--impls takes: " + " ".join(KNOWN)
    return impls


@pytest.fixture(scope="module")
def impls(request):
    impls = list(_impl_params(request.config))
    if "cuda" in impls and not nphusl._cuda.available():
        impls.remove("cuda")                                 # CPU-only box: time what can run here (there is no CPU fallback)
    if "reference" in impls:
        import oracle as orc
        if not orc.have_ref_package():
            impls.remove("reference")
    if not impls:
        pytest.skip("none of the requested --impls can run on this box")
    return impls


@pytest.fixture(scope="module")
def iters(request):
    return request.config.getoption("--iters")


class CachedImg:
    rgb = None
    hsl = None
    name = None


@pytest.fixture(scope="module")
def img(request):
    if CachedImg.rgb is None:
        name = request.config.getoption("--img")
        CachedImg.name = name
        if name in SIZES:
            CachedImg.rgb = np.random.default_rng(0).integers(0, 256, SIZES[name] + (3,), dtype=np.uint8)
        elif name == "cube":
            v = np.arange(256, dtype=np.uint8)
            r, g, b = np.meshgrid(v, v, v, indexing="ij")
            CachedImg.rgb = np.ascontiguousarray(np.stack([r, g, b], -1).reshape(4096, 4096, 3))
        else:
            CachedImg.rgb = np.ascontiguousarray(np.load(name), dtype=np.uint8)
        with nphusl.numpy_enabled():                         # float64 HUSL input for to_rgb, same for every backend
            small = CachedImg.rgb.size <= 3 * 600 * 600
            CachedImg.hsl = nphusl.to_husl(CachedImg.rgb) if small else None
        if CachedImg.hsl is None:
            import oracle as orc
            CachedImg.hsl = orc.rgb_to_husl(CachedImg.rgb)
    return CachedImg


def _runners(impls):
    """-> [(row name, context manager factory, module to call, array adapter)]"""
    rows = []
    for impl in impls:
        if impl == "cuda":
            import torch
            rows.append(("cuda", nphusl.cuda_enabled, nphusl, lambda a: a))
            rows.append(("cuda/device", nphusl.cuda_enabled, nphusl,
                         lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda() if isinstance(a, np.ndarray) else a))
        elif impl == "numpy":
            rows.append(("numpy", nphusl.numpy_enabled, nphusl, lambda a: a))
        elif impl == "reference":
            import oracle as orc
            ref = orc.ref_package()
            rows.append(("reference", ref.best_enabled if hasattr(ref, "best_enabled") else _null, ref, lambda a: a))
    return rows


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _sync():
    if nphusl._cuda.available():
        nphusl._cuda.synchronize()


def _test_all(fn, data, impls, iters, check):
    import tabulate
    dims = "x".join(str(i) for i in np.shape(data)) if np.size(data) > 3 else "3x1"
    print("\n\nnphusl.{}(img)\nbest of {} with dims {} (img: {})\n".format(fn, iters, dims, CachedImg.name))
    px = np.size(data) / 3
    times = {}
    for name, enabled, mod, adapt in _runners(impls):
        arr = adapt(data)
        call = getattr(mod, fn)
        with enabled(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = call(arr)                                   # untimed first call: page-ins, pools, OpenMP teams
            _sync()
            check(name, out)

            def run():
                r = call(arr)
                _sync()                                       # device results are asynchronous: wait for them
                return r
            times[name] = min(timeit.repeat(run, repeat=iters, number=1))
    best = min(times.values())
    rows = [[name, t, px / t, "1.0x (fastest)" if t == best else "{:0.1f}x slower".format(t / best)]
            for name, t in sorted(times.items(), key=lambda kv: kv[1])]
    print(tabulate.tabulate(rows, headers=("Impl", "Duration (s)", "Pixels/s", "Relative speed"), tablefmt="pipe", floatfmt="6.2e"))
    print()
    return times


def _shape_of(out):
    return tuple(out.shape)


def test_perf_rgb_to_husl(impls, iters, img):
    def check(name, out):
        assert _shape_of(out) == img.rgb.shape and np.dtype(out.dtype) == np.float64, name
    assert _test_all("to_husl", img.rgb, impls, iters, check)


def test_perf_husl_to_rgb(impls, iters, img):
    def check(name, out):
        assert _shape_of(out) == img.rgb.shape and np.dtype(out.dtype) == np.uint8, name
    assert _test_all("to_rgb", img.hsl, impls, iters, check)


def test_perf_rgb_to_hue(impls, iters, img):
    def check(name, out):
        assert _shape_of(out) == img.rgb.shape[:2] and np.dtype(out.dtype) == np.float64, name
    assert _test_all("to_hue", img.rgb, impls, iters, check)


def test_perf_rgb_to_husl_one_triplet(impls, iters):
    triplet = [int(v) for v in np.random.default_rng(1).integers(0, 256, 3)]

    def check(name, out):
        assert np.shape(out) == (3,), name
    host_only = [m for m in impls]
    assert _test_all("to_husl", triplet, host_only, iters, check)
