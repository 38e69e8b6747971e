"""GPU parity tests: the CUDA path (through the C ABI of libnphusl_cuda.so,
bound by nphusl/_cuda.py) against the CPU oracle on the same inputs, against
the committed reference-made fixtures, and -- at BASELINE.json's full sizes --
through size-independent properties (exact uint8 round trip, chunk invariance).

Tolerances (BASELINE.json north_star): float HUSL within 1e-3 absolute of the
reference NumPy float64 path (hue compared circularly, masked where S < 0.1 as
reference tests/accuracy_test.py:80 does: only the 256 greys of the cube);
to_rgb round trip within +-1 LSB (we require 0); LUT mode bit-exact."""
import hashlib

import numpy as np
import pytest

import oracle as orc
from conftest import hue_diff
from specs import bands_spec

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def cu():
    from nphusl import _cuda
    _cuda.load()                      # loud failure if the .so is missing
    assert _cuda.available(), "no CUDA device visible"
    return _cuda


def _grey(rgb):
    return (rgb[..., 0] == rgb[..., 1]) & (rgb[..., 1] == rgb[..., 2])


def check_husl(got, ref, hue_mask, tol=TOL):
    got = np.asarray(got, np.float64)
    dh = hue_diff(got[..., 0], ref[..., 0])[hue_mask]
    ds = np.abs(got[..., 1] - ref[..., 1])
    dl = np.abs(got[..., 2] - ref[..., 2])
    assert dh.size == 0 or dh.max() < tol, "hue off by %g" % dh.max()
    assert ds.max() < tol, "saturation off by %g" % ds.max()
    assert dl.max() < tol, "lightness off by %g" % dl.max()
    assert np.all(got[..., 0] >= 0) and np.all(got[..., 0] <= 360.0)


# ---- fixtures made by the reference -------------------------------------------

@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_to_husl_golden(cu, golden, dtype):
    rgb, ref = golden["rgb"], golden["hsl"]
    got = cu._rgb_to_husl(rgb, dtype=dtype)
    assert got.dtype == dtype and got.shape == ref.shape
    check_husl(got, ref, ~_grey(rgb))
    # greys: S ~ 0, the same lightness, and a hue next to the reference's (which is atan2 of
    # float64 rounding noise there: 17.6 .. 21.1 degrees; reference tests allow 5 degrees
    # at low saturation, tests/test.py:130,147)
    g = _grey(rgb)
    assert np.abs(got[g, 1]).max() < TOL
    assert hue_diff(got[g, 0], ref[g, 0]).max() < 2.5


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_to_hue_golden(cu, golden, dtype):
    rgb = golden["rgb"]
    got = cu._rgb_to_hue(rgb, dtype=dtype)
    ng = ~_grey(rgb)
    assert got.shape == rgb.shape[:-1] and got.dtype == dtype
    assert hue_diff(got[ng], golden["hue"][ng]).max() < TOL


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_to_rgb_golden(cu, golden, dtype):
    hsl = golden["hsl_in"].astype(dtype)
    got = cu._husl_to_rgb(hsl)
    assert got.dtype == np.uint8
    ref = golden["rgb_u8"]
    d = np.abs(got.astype(int) - ref.astype(int))
    # a float32 kernel may land on the other side of a .5 rounding boundary
    # for an arbitrary HUSL triplet: +-1 LSB allowed, rarely used
    assert d.max() <= 1
    assert (d != 0).mean() < 2e-3
    f = cu._husl_to_rgb(hsl, dtype=np.float64)
    assert np.abs(f - golden["rgb_float"]).max() < 2e-4
    # exact round trip of reference-made HUSL (tests/test.py:373-377)
    assert np.array_equal(cu._husl_to_rgb(golden["hsl"].astype(dtype)), golden["rgb"])


def test_float_rgb_and_grayscale_inputs(cu, golden):
    frgb, ref = golden["frgb"], golden["frgb_hsl"]
    for dt in (np.float64, np.float32):
        got = cu._rgb_to_husl(frgb.astype(dt))
        tol = TOL if dt == np.float64 else 5e-3       # float32 INPUT quantises the colour itself
        check_husl(got, ref, ref[:, 1] > 0.1, tol)
    grey = golden["grey"]
    g3 = np.broadcast_to(grey[..., None], grey.shape + (3,))   # 0-stride view -> 1-channel kernel
    got = cu._rgb_to_husl(g3)
    assert got.shape == grey.shape + (3,)
    assert np.abs(got[..., 2] - golden["grey_hsl"][..., 2]).max() < TOL
    assert np.abs(got[..., 1] - golden["grey_hsl"][..., 1]).max() < TOL
    got_m = cu._rgb_to_husl(np.repeat(grey[..., None], 3, -1))  # materialised 3-channel
    assert np.array_equal(got_m[..., 1:], got[..., 1:])


def test_known_answers_through_public_api(cu, golden_meta):
    import warnings
    import nphusl
    lit = golden_meta["literals"]
    with nphusl.cuda_enabled():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hsl = nphusl.to_husl([0.5, 0.2, 0.0, 0.5])                  # tests/test.py:671-673
            assert hsl.shape == (3,)
            assert np.allclose(hsl, lit["to_husl_rgba_0.5_0.2_0_0.5"], atol=TOL)
            a = nphusl.to_husl(np.array([0.1, 0.1]))                    # tests/test.py:677-684
            b = nphusl.to_husl(np.array([[.1, .1, .1]] * 2))
            assert np.allclose(a[..., 1:], b[..., 1:], atol=1e-6)
        assert nphusl.to_rgb([360.0, 100.0, 100.0]).tolist() == [255, 255, 255]  # tests/test.py:665-667
        assert nphusl.to_husl(np.array([255, 0, 0], np.uint8))[0] < 13            # tests/test.py:656-661
        assert nphusl.to_husl(np.array([30, 30, 30], np.uint8))[1] < 0.01
        h8 = nphusl.to_husl(np.array([0x80] * 3, np.uint8))
        hf = nphusl.to_husl(np.array([0.5] * 3))
        assert abs(h8[2] - hf[2]) < 0.3
        w = nphusl.to_husl(np.array([255, 255, 255], np.uint8))
        assert abs(w[0] - lit["white_hue"]) < TOL and w[1] == 0 and w[2] == 100
        assert nphusl.to_husl(np.array([0, 0, 0], np.uint8)).tolist() == [0, 0, 0]


# ---- the full 2^24 cube vs the oracle (BASELINE config 2) -----------------------

@pytest.fixture(scope="module")
def cube_oracle(cube):
    return orc.rgb_to_husl(cube)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cube_to_husl(cu, cube, cube_oracle, cube_slices, dtype):
    got = cu._rgb_to_husl(cube, dtype=dtype)
    ng = ~_grey(cube)
    check_husl(got, cube_oracle, ng)
    # and directly against the reference-made samples of the cube
    idx = cube_slices["sample_idx"]
    s = got.reshape(-1, 3)[idx]
    check_husl(s, cube_slices["sample_hsl"], ~_grey(cube.reshape(-1, 3)[idx]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cube_round_trip_exact(cu, cube, dtype):
    hsl = cu._rgb_to_husl(cube, dtype=dtype)
    back = cu._husl_to_rgb(hsl)
    bad = np.count_nonzero(back != cube)
    assert bad == 0, "%d channel values differ after the round trip" % bad


def test_cube_to_rgb_from_oracle_husl(cu, cube, cube_oracle):
    assert np.array_equal(cu._husl_to_rgb(cube_oracle), cube)


def test_cube_to_hue(cu, cube, cube_oracle):
    got = cu._rgb_to_hue(cube)
    ng = ~_grey(cube)
    assert hue_diff(got[ng], cube_oracle[..., 0][ng]).max() < TOL
    husl = cu._rgb_to_husl(cube)
    assert np.array_equal(got[ng], husl[..., 0][ng])    # README.md:44


# ---- LUT mode: bit-exact vs the reference's C path ------------------------------

def test_lut_tables_generated_on_device(cu, simd_golden, golden_meta):
    cu.lut_generate(1024, 1024, 1024)
    t = cu.lut_download()
    assert hashlib.sha256(t["light"].tobytes()).hexdigest() == golden_meta["light_table_sha256"]
    assert np.array_equal(t["linear"], simd_golden["linear_table"])
    assert np.array_equal(t["y_idx_step"], simd_golden["y_idx_step"])
    assert np.array_equal(t["y_thresh"], simd_golden["y_thresh"])
    assert t["h_idx_step"] == float(simd_golden["h_idx_step"])
    assert t["l_idx_step"] == float(simd_golden["l_idx_step"])
    ref_chroma, _, _ = orc.chroma_table()
    diff = np.count_nonzero(t["chroma"] != ref_chroma)
    # device pow/sin/cos are not bit-identical to glibc: entries sitting on a
    # .5 rounding boundary may differ by one count
    assert diff <= 64, "%d chroma entries differ" % diff
    assert np.abs(t["chroma"].astype(int) - ref_chroma.astype(int)).max() <= 1


@pytest.mark.parametrize("mode,key,variant", [("lut", "lut_strict_cube_sha256", "lut"),
                                              ("lut_interp", "lut_interp_strict_cube_sha256", "lut_interp")])
def test_lut_mode_bit_exact_on_cube(cu, cube, golden_meta, mode, key, variant):
    lt, st, th = orc.light_table()
    ct, hs, ls = orc.chroma_table()
    cu.lut_upload(lt, st, th, ct, hs, ls, orc.linear_table6())
    got = cu._rgb_to_husl(cube, mode=mode)
    ref = orc.simd_rgb_to_husl(cube, variant)
    nbad = np.count_nonzero(got.view(np.uint64) != ref.view(np.uint64))
    assert nbad == 0, "%d values not bit-identical" % nbad
    assert hashlib.sha256(got.tobytes()).hexdigest() == golden_meta[key]


def test_lut_mode_golden_samples(cu, simd_golden):
    g = simd_golden
    lt, st, th = orc.light_table()
    ct, hs, ls = orc.chroma_table()
    cu.lut_upload(lt, st, th, ct, hs, ls, g["linear_table"])
    assert np.array_equal(cu._rgb_to_husl(g["rgb"], mode="lut"), g["hsl_lut_strict"])
    assert np.array_equal(cu._rgb_to_husl(g["rgb"], mode="lut_interp"), g["hsl_lut_interp_strict"])


# ---- edge cases: empty, ragged, unaligned, chunked ---------------------------------

def test_empty_and_ragged_sizes(cu):
    rng = np.random.default_rng(3)
    assert cu._rgb_to_husl(np.empty((0, 3), np.uint8)).shape == (0, 3)
    assert cu._husl_to_rgb(np.empty((0, 3))).shape == (0, 3)
    assert cu._rgb_to_hue(np.empty((0, 3), np.uint8)).shape == (0,)
    for n in (1, 2, 3, 31, 127, 128, 129, 255, 257, 1000, 4099, 128 * 37 + 5):
        rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
        ref = orc.rgb_to_husl(rgb)
        for dt in (np.float32, np.float64):
            got = cu._rgb_to_husl(rgb, dtype=dt)
            check_husl(got, ref, ref[:, 1] > 0.1)
            assert np.array_equal(cu._husl_to_rgb(got), rgb)
            hue = cu._rgb_to_hue(rgb, dtype=dt)
            assert np.array_equal(hue, got[:, 0])


def test_unaligned_device_pointers(cu):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(4)
    n = 5000
    rgb = rng.integers(0, 256, (n + 8, 3), dtype=np.uint8)
    t = torch.from_numpy(rgb.reshape(-1)).cuda()
    for off in (0, 1, 2, 3, 5):
        view = t[off * 3 + off: off * 3 + off + n * 3].view(n, 3)   # odd byte offsets
        host = view.cpu().numpy()
        got = cu._rgb_to_husl(view, dtype=np.float32)
        got_h = np.asarray(got)
        ref = cu._rgb_to_husl(host, dtype=np.float32)
        assert np.array_equal(got_h, ref)


def test_streaming_chunks_do_not_change_results(cu):
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)   # BASELINE config 1 shape
    whole = cu._rgb_to_husl(rgb)
    try:
        cu.set_chunk_pixels(100_000)
        parts = cu._rgb_to_husl(rgb)
        back = cu._husl_to_rgb(parts)
    finally:
        cu.set_chunk_pixels(0)
    assert np.array_equal(whole, parts)
    assert np.array_equal(back, rgb)


def test_public_api_chunksize_and_out(cu):
    import nphusl
    rng = np.random.default_rng(6)
    rgb = rng.integers(0, 256, (300, 200, 3), dtype=np.uint8)
    with nphusl.cuda_enabled():
        whole = nphusl.to_husl(rgb)
        out = np.empty(rgb.shape, np.float64)
        r = nphusl.to_husl(rgb, out=out)
        assert r is out and np.array_equal(out, whole)
        tiled = nphusl.to_husl(rgb, chunksize=64)
        assert np.array_equal(tiled, whole)
        out2 = np.zeros(rgb.shape, np.float64)
        nphusl.to_husl(rgb, chunksize=97, out=out2)
        assert np.array_equal(out2, whole)
        hue = nphusl.to_hue(rgb, chunksize=50)
        assert hue.shape == rgb.shape[:2] and np.array_equal(hue, whole[..., 0])
        back = nphusl.to_rgb(whole, chunksize=128)
        assert back.dtype == np.uint8 and np.array_equal(back, rgb)


# ---- device-resident frames (__cuda_array_interface__) ---------------------------------

def test_device_arrays_never_touch_the_host(cu):
    torch = pytest.importorskip("torch")
    import nphusl
    rng = np.random.default_rng(7)
    rgb = rng.integers(0, 256, (2, 270, 480, 3), dtype=np.uint8)
    ref = cu._rgb_to_husl(rgb, dtype=np.float32)
    t = torch.from_numpy(rgb).cuda()
    with nphusl.cuda_enabled():
        d = nphusl.to_husl(t, dtype=np.float32)
        assert isinstance(d, cu.DeviceArray) and d.shape == rgb.shape
        assert np.array_equal(d.copy_to_host(), ref)
        # zero-copy view in torch, edit lightness on the device (README.md:77-81), convert back
        th = torch.as_tensor(d, device="cuda")
        assert th.data_ptr() == d.ptr
        out = torch.empty_like(t)
        nphusl.to_rgb(th, out=out)
        torch.cuda.synchronize()
        assert torch.equal(out, t)
        hue = torch.empty(rgb.shape[:-1], dtype=torch.float64, device="cuda")
        nphusl.to_hue(t, out=hue)
        torch.cuda.synchronize()
        assert np.array_equal(hue.cpu().numpy(), cu._rgb_to_hue(rgb))
    # stream argument: asynchronous on a torch stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        o = torch.empty(rgb.shape, dtype=torch.float32, device="cuda")
        cu._rgb_to_husl(t, out=o, stream=s)
    s.synchronize()
    assert np.array_equal(o.cpu().numpy(), ref)


def test_strided_device_views(cu):
    """Cropped views, channel slices, BGR-as-RGB, pitched rows and permuted planes
    of device tensors are converted where they lie (strided kernels): same values
    as the contiguous copy, nothing outside the view is written."""
    torch = pytest.importorskip("torch")
    import nphusl
    rng = np.random.default_rng(20)
    frame = rng.integers(0, 256, (97, 203, 4), dtype=np.uint8)
    t = torch.from_numpy(frame).cuda()
    views = {
        "crop": (t[10:80, 33:150, :3], frame[10:80, 33:150, :3]),
        "rgba[..., :3]": (t[..., :3], frame[..., :3]),
        "every other column": (t[:, ::2, :3], frame[:, ::2, :3]),
        "planes permuted": (t[..., :3].permute(2, 0, 1).contiguous().permute(1, 2, 0), frame[..., :3]),
    }
    for name, (v, host) in views.items():
        host = np.ascontiguousarray(host)
        assert not v.is_contiguous(), name
        want = cu._rgb_to_husl(host, dtype=np.float32)
        got = cu._rgb_to_husl(v, dtype=np.float32)
        assert isinstance(got, cu.DeviceArray) and got.shape == host.shape, name
        assert np.array_equal(got.copy_to_host(), want), name
        assert np.array_equal(cu._rgb_to_hue(v, dtype=np.float64).copy_to_host(), cu._rgb_to_hue(host)), name
        with nphusl.cuda_enabled():                                   # and through the public API
            assert np.array_equal(nphusl.to_husl(v).copy_to_host(), cu._rgb_to_husl(host)), name
    # BGR memory presented as RGB: negative channel stride (torch has none; raw __cuda_array_interface__)
    bgr = torch.from_numpy(np.ascontiguousarray(frame[..., 2::-1])).cuda()

    class Raw:
        device = 0
        __cuda_array_interface__ = {"shape": (97, 203, 3), "typestr": "|u1", "data": (bgr.data_ptr() + 2, False),
                                    "version": 3, "strides": (203 * 3, 3, -1)}
    assert np.array_equal(cu._rgb_to_husl(Raw(), dtype=np.float32).copy_to_host(),
                          cu._rgb_to_husl(np.ascontiguousarray(frame[..., :3]), dtype=np.float32))
    # float RGB view
    f = torch.from_numpy(frame[..., :3] / 255.0).cuda()
    assert np.array_equal(cu._rgb_to_husl(f[5:50, 7:99]).copy_to_host(), cu._rgb_to_husl(np.ascontiguousarray(frame[5:50, 7:99, :3] / 255.0)))
    # strided HUSL in, strided uint8 out: write a region of a larger frame, the rest stays untouched
    hsl = torch.from_numpy(cu._rgb_to_husl(np.ascontiguousarray(frame[..., :3]), dtype=np.float32)).cuda()
    canvas = torch.full((97, 203, 4), 0xA5, dtype=torch.uint8, device="cuda")
    cu._husl_to_rgb(hsl[10:80, 33:150], out=canvas[10:80, 33:150, :3])
    torch.cuda.synchronize()
    c = canvas.cpu().numpy()
    assert np.array_equal(c[10:80, 33:150, :3], frame[10:80, 33:150, :3])
    c[10:80, 33:150, :3] = 0xA5
    assert (c == 0xA5).all()
    # fused edit of a region of interest, in place
    work = t.clone()
    roi = work[20:60, 40:140, :3]
    r = cu.edit_rgb(roi, out=roi, lightness=(0, 50), l=(1, 20))
    torch.cuda.synchronize()
    w = work.cpu().numpy()
    assert r is roi
    assert np.array_equal(w[20:60, 40:140, :3], cu.edit_rgb(np.ascontiguousarray(frame[20:60, 40:140, :3]), lightness=(0, 50), l=(1, 20)))
    w[20:60, 40:140, :3] = frame[20:60, 40:140, :3]
    assert np.array_equal(w, frame)
    # a batch of cropped frames (three strided pixel axes)
    vid = torch.from_numpy(rng.integers(0, 256, (4, 50, 70, 3), dtype=np.uint8)).cuda()
    crop = vid[:, 5:45, 10:61]
    assert np.array_equal(cu._rgb_to_husl(crop, dtype=np.float32).copy_to_host(),
                          cu._rgb_to_husl(crop.contiguous().cpu().numpy(), dtype=np.float32))
    # what cannot be expressed as frames x rows x cols is refused, not copied behind the caller's back
    five = torch.zeros((4, 6, 8, 10, 12, 3), dtype=torch.uint8, device="cuda")[:, ::2, ::2, ::2, ::2]
    with pytest.raises(ValueError):
        cu._rgb_to_husl(five)
    with pytest.raises(ValueError):
        cu._rgb_to_husl(t[:, ::2, :3], mode="lut")
    with pytest.raises(ValueError):
        cu.rgb_to_husl_planar(t[:, ::2, :3])                          # planar entry points: contiguous only


# ---- BASELINE full-size workloads: size-independent properties -----------------------

def test_4k_batch_round_trip_and_edit(cu):
    """BASELINE config 4: batch of 3840x2160 frames, to_husl -> lightness edit
    -> to_rgb, device-resident."""
    torch = pytest.importorskip("torch")
    g = torch.Generator(device="cuda").manual_seed(0)
    frames = torch.randint(0, 256, (4, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
    hsl = torch.empty(frames.shape, dtype=torch.float32, device="cuda")
    back = torch.empty_like(frames)
    cu._rgb_to_husl(frames, out=hsl)
    cu._husl_to_rgb(hsl, out=back)
    torch.cuda.synchronize()
    assert torch.equal(back, frames)                       # exact at full size
    assert float(hsl[..., 1].max()) <= 100.0 + TOL and float(hsl[..., 2].max()) <= 100.0
    # spot-check a strided sample against the oracle
    smp = frames.view(-1, 3)[::9973].cpu().numpy()
    check_husl(hsl.view(-1, 3)[::9973].cpu().numpy(), orc.rgb_to_husl(smp), orc.rgb_to_husl(smp)[:, 1] > 0.1)
    # lightness edit: darker frame must convert to an RGB that maps back to the edited L
    hsl[..., 2] *= 0.5
    cu._husl_to_rgb(hsl, out=back)
    again = torch.empty_like(hsl)
    cu._rgb_to_husl(back, out=again)
    torch.cuda.synchronize()
    dl = (again[..., 2] - hsl[..., 2]).abs().max().item()
    assert dl < 0.5          # uint8 quantisation of RGB, measured in L units


def test_launch_counter_moves(cu):
    n0 = cu.launch_count()
    cu._rgb_to_husl(np.zeros((1000, 3), np.uint8))
    assert cu.launch_count() > n0


def test_non_blocking_pinned_calls(cu):
    """blocking=False: pinned host buffers are only enqueued; results are valid
    after the stream is synchronised; two streams may be in flight at once."""
    rng = np.random.default_rng(8)
    n = 5
    src = [cu.pinned_empty((1080, 1920, 3), np.uint8) for _ in range(n)]
    dst = [cu.pinned_empty((1080, 1920, 3), np.uint8) for _ in range(n)]
    for a, b in zip(src, dst):
        a[...] = rng.integers(0, 256, a.shape, dtype=np.uint8)
        b[...] = 0
    streams = [cu.Stream(), cu.Stream()]
    hsl = [cu.DeviceArray((1080, 1920, 3), np.float32), cu.DeviceArray((1080, 1920, 3), np.float32)]
    try:
        cu.set_chunk_pixels(300_000)          # several chunks per call
        for k in range(n):
            s = streams[k & 1]
            cu._rgb_to_husl(src[k], out=hsl[k & 1], stream=s, blocking=False)
            cu._husl_to_rgb(hsl[k & 1], out=dst[k], stream=s, blocking=False)
        for s in streams:
            s.synchronize()
    finally:
        cu.set_chunk_pixels(0)
    for a, b in zip(src, dst):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    # pageable memory ignores blocking=False and completes before returning
    rgb = rng.integers(0, 256, (1000, 3), dtype=np.uint8)
    out = cu._rgb_to_husl(rgb, blocking=False)
    assert np.array_equal(cu._husl_to_rgb(out, blocking=False), rgb)


# ---- callers' side on the GPU: planar HUSL, fused edit ---------------------------------------

@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_planar_layout_equals_interleaved(cu, dtype):
    rng = np.random.default_rng(9)
    for shape in ((1000, 3), (37, 129, 3), (1080, 1920, 3)):
        rgb = rng.integers(0, 256, shape, dtype=np.uint8)
        inter = cu._rgb_to_husl(rgb, dtype=dtype)
        planar = cu.rgb_to_husl_planar(rgb, dtype=dtype)
        assert planar.shape == (3,) + shape[:-1] and planar.dtype == dtype
        assert np.array_equal(planar, np.moveaxis(inter, -1, 0))
        assert np.array_equal(cu.husl_planar_to_rgb(planar), rgb)


def test_planar_device_arrays(cu):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(10)
    rgb = rng.integers(0, 256, (2, 540, 960, 3), dtype=np.uint8)
    t = torch.from_numpy(rgb).cuda()
    planes = torch.empty((3, 2, 540, 960), dtype=torch.float32, device="cuda")
    r = cu.rgb_to_husl_planar(t, out=planes)
    assert r is planes
    torch.cuda.synchronize()
    assert np.array_equal(planes.cpu().numpy(), np.moveaxis(cu._rgb_to_husl(rgb, dtype=np.float32), -1, 0))
    planes[2] *= 0.5                                   # contiguous plane edit
    out = torch.empty_like(t)
    cu.husl_planar_to_rgb(planes, out=out)
    torch.cuda.synchronize()
    hsl = cu._rgb_to_husl(rgb, dtype=np.float32)
    hsl[..., 2] *= np.float32(0.5)
    assert np.array_equal(out.cpu().numpy(), cu._husl_to_rgb(hsl))


def test_fused_edit_equals_three_call_pipeline(cu):
    """nphusl.edit == to_rgb(edit(to_husl(rgb, float32))) bit for bit; and within
    +-1 LSB of the float64 oracle doing the same edit."""
    import nphusl
    rng = np.random.default_rng(11)
    for shape in ((777, 3), (1080, 1920, 3)):
        rgb = rng.integers(0, 256, shape, dtype=np.uint8)
        # README.md:77-81: darken everything that is not bluish
        hsl = cu._rgb_to_husl(rgb, dtype=np.float32)
        bluish = (hsl[..., 0] >= 250) & (hsl[..., 0] <= 290)
        hsl[..., 2][~bluish] *= np.float32(0.5)
        want = cu._husl_to_rgb(hsl)
        got = nphusl.edit(rgb, hue=(250, 290), invert=True, l=(0.5, 0))
        assert got.dtype == np.uint8 and np.array_equal(got, want)
        o = orc.rgb_to_husl(rgb)
        ob = (o[..., 0] >= 250) & (o[..., 0] <= 290)
        o[..., 2][~ob] *= 0.5
        od = np.abs(orc.husl_to_rgb(o).astype(int) - got.astype(int))
        edge = np.abs(np.minimum(np.abs(o[..., 0] - 250), np.abs(o[..., 0] - 290))) < 1e-3   # selection boundary
        assert od[~edge].max() <= 1
        # README.md:94-98: dim pixels to black
        hsl = cu._rgb_to_husl(rgb, dtype=np.float32)
        dark = hsl[..., 2] < 62
        hsl[..., 2][dark] = 0
        want = cu._husl_to_rgb(hsl)
        hi = float(np.nextafter(np.float32(62), np.float32(0)))
        got = nphusl.edit(rgb, lightness=(0, hi), l=(0, 0))
        assert np.array_equal(got, want)
        assert np.all(got[dark] == 0)
        # hue rotation with wrap-around arc selection, saturation scaling
        hsl = cu._rgb_to_husl(rgb, dtype=np.float32)
        sel = (hsl[..., 0] >= 330) | (hsl[..., 0] <= 30)
        h = hsl[..., 0][sel] + np.float32(90)
        h = h - np.float32(360) * np.floor(h * np.float32(1.0 / 360.0))
        hsl[..., 0][sel] = h
        hsl[..., 1][sel] = np.minimum(hsl[..., 1][sel] * np.float32(0.5) + np.float32(10), np.float32(100))
        want = cu._husl_to_rgb(hsl)
        got = nphusl.edit(rgb, hue=(330, 30), h=(1, 90), s=(0.5, 10))
        assert np.array_equal(got, want)
        # identity edit = exact round trip
        assert np.array_equal(nphusl.edit(rgb), rgb)


def test_fused_edit_device_in_place(cu):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(12)
    rgb = rng.integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
    want = cu.edit_rgb(rgb, lightness=(0, 50), l=(1, 20))
    t = torch.from_numpy(rgb).cuda()
    r = cu.edit_rgb(t, out=t, lightness=(0, 50), l=(1, 20))
    torch.cuda.synchronize()
    assert r is t and np.array_equal(t.cpu().numpy(), want)


def test_cuda_graph_capture_of_per_frame_loop(cu):
    """Per-frame pipelines are launch-bound (a 4K frame is ~40 us of kernel
    time): device-pointer calls are plain launches on the caller's stream, so a
    to_husl -> to_rgb frame loop can be captured in a CUDA graph and replayed."""
    torch = pytest.importorskip("torch")
    g = torch.Generator(device="cuda").manual_seed(3)
    frames = torch.randint(0, 256, (6, 720, 1280, 3), dtype=torch.uint8, device="cuda", generator=g)
    cur = torch.empty_like(frames[0])
    hsl = torch.empty(cur.shape, dtype=torch.float32, device="cuda")
    out = torch.empty_like(cur)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        cu._rgb_to_husl(cur, out=hsl, stream=s)           # warm-up: one-time kernel attribute setup
        cu._husl_to_rgb(hsl, out=out, stream=s)
    s.synchronize()
    graph = torch.cuda.CUDAGraph()
    n0 = cu.launch_count()
    with torch.cuda.graph(graph, stream=s):
        cu._rgb_to_husl(cur, out=hsl, stream=torch.cuda.current_stream())
        cu._husl_to_rgb(hsl, out=out, stream=torch.cuda.current_stream())
    assert cu.launch_count() - n0 == 2
    for f in frames:
        cur.copy_(f)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, f)


def test_guard_bands_stay_untouched(cu):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are
    caught the old way: every output lives inside a larger canary-filled buffer
    and the canaries on both sides must survive every kernel family (streaming
    TMA / cp.async / LDG kernels, their tails, generic, planar, fused edit)."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(13)
    GUARD = 4096

    def guarded(nbytes, tdtype):
        raw = torch.full((GUARD + nbytes + GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
        view = raw[GUARD:GUARD + nbytes].view(tdtype)
        return raw, view

    def intact(raw, nbytes):
        return bool((raw[:GUARD] == 0xA5).all()) and bool((raw[GUARD + nbytes:] == 0xA5).all())

    for n in (1, 127, 128, 129, 128 * 37 + 5, 128 * 4000, 128 * 4000 + 91):
        rgb = torch.from_numpy(rng.integers(0, 256, (n, 3), dtype=np.uint8)).cuda()
        for tdt, isz in ((torch.float32, 4), (torch.float64, 8)):
            raw, hsl = guarded(n * 3 * isz, tdt)
            cu._rgb_to_husl(rgb, out=hsl.view(n, 3))
            rawb, back = guarded(n * 3, torch.uint8)
            cu._husl_to_rgb(hsl.view(n, 3), out=back.view(n, 3))
            rawh, hue = guarded(n * isz, tdt)
            cu._rgb_to_hue(rgb, out=hue)
            rawp, planes = guarded(n * 3 * isz, tdt)
            cu.rgb_to_husl_planar(rgb, out=planes.view(3, n))
            rawq, back2 = guarded(n * 3, torch.uint8)
            cu.husl_planar_to_rgb(planes.view(3, n), out=back2.view(n, 3))
            torch.cuda.synchronize()
            assert intact(raw, n * 3 * isz) and intact(rawb, n * 3) and intact(rawh, n * isz)
            assert intact(rawp, n * 3 * isz) and intact(rawq, n * 3)
            assert torch.equal(back.view(n, 3), rgb) and torch.equal(back2.view(n, 3), rgb)
            # channels-first planes both ways (cp.async ring kernel for float32 planes)
            chw = rgb.t().contiguous()
            rawc, cplanes = guarded(n * 3 * isz, tdt)
            cu.rgb_to_husl_planar(chw, out=cplanes.view(3, n), channels_first=True)
            rawd, cback = guarded(n * 3, torch.uint8)
            cu.husl_planar_to_rgb(cplanes.view(3, n), out=cback.view(3, n), channels_first=True)
            torch.cuda.synchronize()
            assert intact(rawc, n * 3 * isz) and intact(rawd, n * 3) and torch.equal(cback.view(3, n), chw)
        rawe, ed = guarded(n * 3, torch.uint8)
        cu.edit_rgb(rgb, out=ed.view(n, 3))
        rawf, edc = guarded(n * 3, torch.uint8)
        cu.edit_rgb(rgb.t().contiguous(), out=edc.view(3, n), channels_first=True)
        torch.cuda.synchronize()
        assert intact(rawe, n * 3) and torch.equal(ed.view(n, 3), rgb)
        assert intact(rawf, n * 3) and torch.equal(edc.view(3, n), rgb.t())


def test_shard_over_devices(cu):
    """In-process fan-out of a host batch over every visible GPU (one host thread
    per GPU, disjoint output slices, no collective) equals the single-GPU result."""
    from nphusl import _shard
    ndev = cu.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    rng = np.random.default_rng(14)
    frames = rng.integers(0, 256, (2 * ndev + 1, 540, 960, 3), dtype=np.uint8)
    devs = list(range(ndev))
    hsl = _shard.to_husl(frames, devices=devs, dtype=np.float32)
    assert np.array_equal(hsl, cu._rgb_to_husl(frames, dtype=np.float32, device=0))
    assert np.array_equal(_shard.to_rgb(hsl, devices=devs), frames)
    assert np.array_equal(_shard.to_hue(frames, devices=devs), cu._rgb_to_hue(frames, device=ndev - 1))
    # a tensor living on GPU 1 is converted on GPU 1 and the caller's current device is left alone
    torch = pytest.importorskip("torch")
    t = torch.from_numpy(frames[0]).to("cuda:1")
    assert torch.cuda.current_device() == 0
    d = cu._rgb_to_husl(t, dtype=np.float32)
    assert d.device == 1 and torch.cuda.current_device() == 0
    assert np.array_equal(d.copy_to_host(), hsl[0])
    # a host array converted INTO a tensor that lives on GPU 1: the call goes where `out` lives
    o1 = torch.empty(frames[0].shape, dtype=torch.float32, device="cuda:1")
    cu._rgb_to_husl(frames[0], out=o1)
    torch.cuda.synchronize(1)
    assert np.array_equal(o1.cpu().numpy(), hsl[0]) and torch.cuda.current_device() == 0


def test_out_of_gamut_husl_wraps_like_the_reference(cu):
    """to_rgb never clamps (reference transform.py:16-20): negative
    channels are mirrored by abs(), values past 255 wrap modulo 256."""
    rng = np.random.default_rng(15)
    n = 20000
    hsl = np.stack([rng.uniform(0, 360, n), rng.uniform(100, 180, n), rng.uniform(5, 95, n)], -1)
    want = orc.husl_to_rgb(hsl).astype(int)
    f = orc.husl_to_rgb_float(hsl)
    assert (f.max() > 1.0) and (f.min() < 0.0)                       # really out of gamut
    # near a gamut singularity (v' -> 0) a channel explodes to 1e6 and beyond, where the cast
    # to uint8 is undefined behaviour in C and numpy alike: compare the sane 99 %
    sane = (np.abs(f) < 16.0).all(-1)
    assert sane.mean() > 0.9
    for dt in (np.float64, np.float32):
        got = cu._husl_to_rgb(hsl.astype(dt)).astype(int)
        d = np.abs(got - want)[sane]
        d = np.minimum(d, 256 - d)                                   # wrap-around distance
        assert d.max() <= 1 and (d != 0).mean() < 5e-3
    # float RGB output is the unclamped value itself
    gf = cu._husl_to_rgb(hsl, dtype=np.float64)
    assert (np.abs(gf - f)[sane] / np.maximum(1.0, np.abs(f[sane]))).max() < 5e-4


def test_hue_outside_0_360_is_periodic(cu, cube):
    """Callers rotate hues without wrapping (`hsl[..., 0] += 120`): the reference feeds
    the angle to sin / cos (_cython_opt.pyx:219-221, nphusl.py:170-171), so to_rgb is
    periodic in H.  Every colour of a cube slice must come back to the same uint8 RGB
    when its hue is given one or two turns up or down."""
    rgb = np.ascontiguousarray(cube.reshape(-1, 3)[::17])  # ~1 M colours spread over the cube
    for dt in (np.float64, np.float32):
        hsl = np.asarray(cu._rgb_to_husl(rgb, dtype=dt))
        for turns in (-2, -1, 1, 2):
            h2 = hsl.copy()
            h2[:, 0] += 360.0 * turns
            back = np.asarray(cu._husl_to_rgb(h2))
            if dt == np.float64:
                assert np.array_equal(back, rgb), "turns=%d" % turns
            else:                                         # float32 H + 720 has lost 2 bits of the angle itself
                d = np.abs(back.astype(int) - rgb.astype(int))
                assert d.max() <= 1 and (d != 0).mean() < 1e-3, "turns=%d" % turns
        want = orc.husl_to_rgb(np.asarray(hsl, np.float64) + np.array([720.0, 0, 0]))
        assert np.array_equal(want, rgb)                  # the oracle (reference math) is periodic too


def test_rgba_device_arrays_are_premultiplied_on_load(cu):
    """4-channel device input: rgb * alpha in the load stage of the kernel
    (reference transform.py:221-230 does it as a host pass).  Float RGBA follows
    the reference exactly (tests/test.py:547-556); uint8 RGBA is premultiplied in
    float, c/255 * a/255 (the reference's integer branch is broken)."""
    import warnings
    torch = pytest.importorskip("torch")
    import nphusl
    rng = np.random.default_rng(16)
    rgba = rng.random((33, 47, 4))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = nphusl.to_husl(rgba)                                   # host: NumPy premultiply, then the GPU
        got = nphusl.to_husl(torch.from_numpy(rgba).cuda())
        hue = nphusl.to_hue(torch.from_numpy(rgba).cuda())
    assert isinstance(got, cu.DeviceArray) and got.shape == (33, 47, 3)
    g = got.copy_to_host()
    assert np.abs(g - want)[..., 1:].max() < 1e-4
    assert hue_diff(g[..., 0], want[..., 0]).max() < 1e-4
    assert np.array_equal(hue.copy_to_host(), g[..., 0])
    u8 = rng.integers(0, 256, (1000, 4), dtype=np.uint8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got8 = nphusl.to_husl(torch.from_numpy(u8).cuda()).copy_to_host()
        host8 = nphusl.to_husl(u8)                                    # host uint8 RGBA: the same values
    ref8 = orc.rgb_to_husl((u8[:, :3] / 255.0) * (u8[:, 3:] / 255.0))
    sat = ref8[:, 1] > 0.1
    assert hue_diff(got8[sat, 0], ref8[sat, 0]).max() < TOL
    assert np.abs(got8[:, 1:] - ref8[:, 1:]).max() < TOL
    assert hue_diff(host8[sat, 0], got8[sat, 0]).max() < 2e-4
    assert np.abs(host8[:, 1:] - got8[:, 1:]).max() < 2e-4


def test_lut_mode_with_other_table_sizes(cu):
    """make_lookup_tables takes the table sizes as arguments (reference
    make_lookup_tables:4-6); the device generator and the LUT kernels follow."""
    rng = np.random.default_rng(17)
    rgb = rng.integers(0, 256, (50000, 3), dtype=np.uint8)
    seg, nh, nl = 256, 300, 200
    cu.lut_generate(seg, nh, nl)
    t = cu.lut_download()
    lt, st, th = orc.light_table(seg)
    ct, hs, ls = orc.chroma_table(nh, nl)
    assert t["light"].shape == (3 * seg,) and t["chroma"].shape == (nh, nl)
    assert np.array_equal(t["light"], lt) and np.array_equal(t["y_idx_step"], st) and np.array_equal(t["y_thresh"], th)
    assert t["h_idx_step"] == hs and t["l_idx_step"] == ls
    assert np.abs(t["chroma"].astype(int) - ct.astype(int)).max() <= 1
    assert np.count_nonzero(t["chroma"] != ct) <= 16
    # with identical (oracle-made) tables the kernel is bit-exact for any size
    cu.lut_upload(lt, st, th, ct, hs, ls, orc.linear_table6())
    for mode, variant in (("lut", "lut"), ("lut_interp", "lut_interp")):
        got = cu._rgb_to_husl(rgb, mode=mode)
        want = orc.simd_rgb_to_husl(rgb, variant, tables=(lt, st, th, ct, hs, ls))
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    with pytest.raises(cu.CudaError):
        cu._rgb_to_husl(rgb.astype(np.float32) / 255, mode="lut")      # LUT modes take uint8, like _simd.c


def test_float32_channels_near_grey_near_the_switch_near_the_ends(cu):
    """float32 RGB goes through the float64-free formulation (husl_math.cuh: diff_from_f32) in the tiled kernel,
    the one-pixel-per-thread kernel (odd sizes) and the strided kernel.  The cases a float32-only chain could
    lose -- channels a few ulps apart, both sides of the 0.04045 segment switch, near 0, near 1 -- against the
    float64 oracle at the bar of 1e-3 (hue where S > 0.1), both output types; to_hue equals the H channel."""
    rng = np.random.default_rng(43)
    n = 128 * 400 + 77

    def ulps(x, k):
        return (x.view(np.int32) + k.astype(np.int32)).view(np.float32)
    g = rng.random(n, dtype=np.float32) * np.float32(0.98) + np.float32(0.01)
    sets = {"uniform": rng.random((n, 3), dtype=np.float32)}
    for name, kmax in (("2ulp", 2), ("1e4ulp", 10 ** 4), ("1e6ulp", 10 ** 6)):
        a = np.stack([ulps(g, rng.integers(-kmax, kmax + 1, n)), g, ulps(g, rng.integers(-kmax, kmax + 1, n))], 1)
        sets[name] = np.clip(a, 0, 1).astype(np.float32)
    t = np.float32(0.04045)
    sets["switch"] = (t + (rng.random((n, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(0.02)).astype(np.float32)
    sets["switch_tight"] = ulps(np.full((n, 3), t, np.float32), rng.integers(-50, 51, (n, 3)))
    sets["dark"] = rng.random((n, 3), dtype=np.float32) * np.float32(0.05)
    sets["bright"] = np.float32(1) - rng.random((n, 3), dtype=np.float32) * np.float32(0.02)
    for name, a in sets.items():
        a = np.ascontiguousarray(a)
        ref = orc.rgb_to_husl(a.astype(np.float64))
        # away from the L > 99.99 / L < 0.01 clamps, where a float32 Y and a float64 Y may fall on different sides
        mid = (ref[:, 2] > 0.02) & (ref[:, 2] < 99.98)
        for odt in (np.float32, np.float64):
            got = np.asarray(cu._rgb_to_husl(a, dtype=odt), np.float64)
            assert got.dtype == np.float64 and got.shape == a.shape
            check_husl(got[mid], ref[mid], ref[mid, 1] > 0.1)
            hue = cu._rgb_to_hue(a, dtype=odt)
            assert np.array_equal(np.asarray(hue, np.float64), got[:, 0]), name
    # NaN channels stay NaN
    bad = np.array([[np.nan, 0.5, 0.5], [0.5, np.nan, 0.5], [0.5, 0.5, np.nan]] * 50, np.float32)
    assert np.isnan(cu._rgb_to_husl(bad, dtype=np.float32)).any(axis=1).all()


def test_lut_mode_pixels_the_streaming_pass_hands_back(cu):
    """The LUT tile kernel's straight-line pass leaves black, white and u == 0 pixels (L index 0: u = v = 0)
    to a second, complete pass over the lane that met one.  Whole tiles of them (letterbox bars, blown
    highlights), single ones between ordinary pixels, and tables where more or fewer colours land in that
    class (a light table whose first entries are 0, and one whose first entry is NOT 0: black must still be
    answered with the constants of _simd.c:205-212) -- against the oracle, bit for bit, and against the
    one-pixel-per-thread kernel (an odd base address keeps the data off the tile kernel)."""
    rng = np.random.default_rng(41)
    n = 128 * 300
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    rgb[128 * 3:128 * 9] = 0                       # whole tiles of black
    rgb[128 * 20:128 * 23] = 255                   # ... and of white
    rgb[128 * 40:128 * 44] = (0, 0, 1)             # L index rounds to 0: u == v == 0 without being black
    dark = rng.integers(0, 3, (128 * 30, 3), dtype=np.uint8)
    rgb[128 * 60:128 * 90] = dark                  # every colour of {0, 1, 2}^3, mixed within lanes
    rgb[128 * 100 + 5] = 0; rgb[128 * 101 + 77] = 255; rgb[128 * 102 + 127] = (0, 0, 1)
    grey = rng.integers(0, 256, 128 * 20, dtype=np.uint8)
    rgb[128 * 120:128 * 140] = grey[:, None]       # greys: u', v' at the white point
    lt, st, th = orc.light_table()
    ct, hs, ls = orc.chroma_table()
    lt_zeros = lt.copy(); lt_zeros[:40] = 0        # many dark colours get L == 0
    lt_lift = lt.copy(); lt_lift[0] = 7            # L(y = 0) != 0: black is no longer a fixed point of the chain
    torch = pytest.importorskip("torch")
    buf = np.zeros(rgb.size + 1, np.uint8)
    buf[1:] = rgb.reshape(-1)
    odd = torch.from_numpy(buf).cuda()[1:].view(n, 3)
    for light in (lt, lt_zeros, lt_lift):
        cu.lut_upload(light, st, th, ct, hs, ls, orc.linear_table6())
        for mode in ("lut", "lut_interp"):
            want = orc.simd_rgb_to_husl(rgb, mode, tables=(light, st, th, ct, hs, ls))
            got = cu._rgb_to_husl(rgb, mode=mode)
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), mode
            # the same pixels through the one-pixel-per-thread kernel
            got1 = np.asarray(cu._rgb_to_husl(odd, mode=mode))
            assert np.array_equal(got1.view(np.uint64), want.view(np.uint64)), mode
    cu.lut_upload(lt, st, th, ct, hs, ls, orc.linear_table6())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_channels_first_images(cu, dtype):
    """(3, H, W) uint8 planes (torch / vision layout) <-> (3, H, W) HUSL planes:
    same values as the interleaved path, exact round trip, ragged sizes."""
    rng = np.random.default_rng(18)
    for hw in ((7, 13), (37, 129), (1080, 1920)):
        chw = rng.integers(0, 256, (3,) + hw, dtype=np.uint8)
        hwc = np.ascontiguousarray(np.moveaxis(chw, 0, -1))
        planes = cu.rgb_to_husl_planar(chw, dtype=dtype, channels_first=True)
        assert planes.shape == (3,) + hw and planes.dtype == dtype
        assert np.array_equal(planes, np.moveaxis(cu._rgb_to_husl(hwc, dtype=dtype), -1, 0))
        back = cu.husl_planar_to_rgb(planes, channels_first=True)
        assert back.shape == chw.shape and np.array_equal(back, chw)
    torch = pytest.importorskip("torch")
    t = torch.from_numpy(chw).cuda()
    out = torch.empty(chw.shape, dtype=torch.float32 if dtype == np.float32 else torch.float64, device="cuda")
    cu.rgb_to_husl_planar(t, out=out, channels_first=True)
    back = torch.empty_like(t)
    cu.husl_planar_to_rgb(out, out=back, channels_first=True)
    torch.cuda.synchronize()
    assert torch.equal(back, t)


def test_cube_through_planar_chw_and_fused_paths(cu, cube, cube_oracle):
    """All 2^24 colours through the other layouts: planar HUSL, channels-first
    planes and the fused identity edit must reproduce the interleaved result /
    the input exactly."""
    inter = cu._rgb_to_husl(cube, dtype=np.float32)
    planar = cu.rgb_to_husl_planar(cube, dtype=np.float32)
    assert np.array_equal(planar, np.moveaxis(inter, -1, 0))
    assert np.array_equal(cu.husl_planar_to_rgb(planar), cube)
    chw = np.ascontiguousarray(np.moveaxis(cube, -1, 0))
    p2 = cu.rgb_to_husl_planar(chw, dtype=np.float32, channels_first=True)
    assert np.array_equal(p2, planar)
    assert np.array_equal(cu.husl_planar_to_rgb(p2, channels_first=True), chw)
    assert np.array_equal(cu.edit_rgb(cube), cube)
    check_husl(np.moveaxis(planar, 0, -1), cube_oracle, ~_grey(cube))


def test_fused_edit_on_channels_first_planes(cu):
    """edit_rgb(channels_first=True) on (3, H, W) uint8 device planes equals the
    interleaved fused edit pixel for pixel, also in place and at ragged sizes."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(19)
    for hw in ((5, 9), (61, 131), (1080, 1920)):
        hwc = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
        want = cu.edit_rgb(hwc, hue=(250, 290), invert=True, l=(0.5, 0))
        t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(hwc, -1, 0))).cuda()
        got = cu.edit_rgb(t, channels_first=True, hue=(250, 290), invert=True, l=(0.5, 0))
        assert isinstance(got, cu.DeviceArray) and got.shape == (3,) + hw
        assert np.array_equal(np.moveaxis(got.copy_to_host(), 0, -1), want)
        r = cu.edit_rgb(t, out=t, channels_first=True, hue=(250, 290), invert=True, l=(0.5, 0))
        torch.cuda.synchronize()
        assert r is t and np.array_equal(np.moveaxis(t.cpu().numpy(), 0, -1), want)
    with pytest.raises(ValueError):
        cu.edit_rgb(hwc, channels_first=True)                      # host arrays: interleaved only


def test_decoder_adjacent_frames_never_touch_the_host(cu):
    """a JPEG decoded ON the GPU (torchvision.io.decode_jpeg(device="cuda"), nvJPEG)
    arrives as (3, H, W) uint8 planes; the channels-first kernels convert / edit it in place of the
    reference's imageio -> numpy -> to_husl path and the result is re-encoded on the GPU.  Checked
    against the host path on the same decoded pixels."""
    torch = pytest.importorskip("torch")
    tvio = pytest.importorskip("torchvision.io")
    rng = np.random.default_rng(21)
    # a smooth synthetic picture (JPEG-friendly), encoded on the CPU
    yy, xx = np.mgrid[0:360, 0:640]
    pic = np.stack([(127 + 120 * np.sin(xx / 40.0)), (127 + 120 * np.cos(yy / 30.0)), (xx + yy) % 256], 0).astype(np.uint8)
    pic[:, 100:140, 200:300] = rng.integers(0, 256, (3, 40, 100), dtype=np.uint8)
    jpeg = tvio.encode_jpeg(torch.from_numpy(pic), quality=92)
    try:
        frame = tvio.decode_jpeg(jpeg, device="cuda")                  # (3, H, W) uint8 on the device
    except Exception as e:                                             # torchvision built without nvJPEG
        pytest.skip("GPU JPEG decode unavailable: %r" % (e,))
    assert frame.is_cuda and frame.dtype == torch.uint8 and tuple(frame.shape) == pic.shape
    host = np.ascontiguousarray(np.moveaxis(frame.cpu().numpy(), 0, -1))   # the same decoded pixels, HWC on the host
    # planes -> HUSL planes -> planes: exact round trip, values equal to the interleaved host path
    planes = cu.rgb_to_husl_planar(frame, dtype=np.float32, channels_first=True)
    assert np.array_equal(np.moveaxis(planes.copy_to_host(), 0, -1), cu._rgb_to_husl(host, dtype=np.float32))
    back = torch.empty_like(frame)
    cu.husl_planar_to_rgb(planes, out=back, channels_first=True)
    torch.cuda.synchronize()
    assert torch.equal(back, frame)
    # fused edit on the decoded planes, in place, then re-encode on the GPU
    want = cu.edit_rgb(host, hue=(250, 290), invert=True, l=(0.5, 0))
    cu.edit_rgb(frame, out=frame, channels_first=True, hue=(250, 290), invert=True, l=(0.5, 0))
    torch.cuda.synchronize()
    assert np.array_equal(np.moveaxis(frame.cpu().numpy(), 0, -1), want)
    try:
        out_jpeg = tvio.encode_jpeg(frame, quality=92)
    except Exception as e:
        pytest.skip("GPU JPEG encode unavailable: %r" % (e,))
    redecoded = tvio.decode_jpeg(out_jpeg.cpu())
    assert tuple(redecoded.shape) == pic.shape
    assert np.abs(redecoded.numpy().astype(int) - np.moveaxis(want, -1, 0).astype(int)).mean() < 6.0


def test_host_results_are_pooled_page_locked_arrays(cu):
    """Host-array calls return plain ndarrays carved out of the page-locked result pool: the D2H
    copy lands in them directly, dropping them recycles the block, and feeding one back into
    to_rgb is a direct DMA.  Values are those of the pageable path."""
    import ctypes
    import gc
    import nphusl
    rng = np.random.default_rng(22)
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)

    def pinned(a):
        loc, dev, pin = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        cu.load().nphusl_cuda_pointer_info(ctypes.c_void_p(a.ctypes.data), ctypes.byref(loc), ctypes.byref(dev), ctypes.byref(pin))
        return loc.value == cu.HOST and pin.value == 1

    cu.set_pinned_results(False)
    want = nphusl.to_husl(img)
    assert not pinned(want)
    cu.set_pinned_results(True)
    s0 = cu.pinned_pool_stats()
    with nphusl.cuda_enabled():
        hsl = nphusl.to_husl(img)
        assert type(hsl) is np.ndarray and pinned(hsl) and np.array_equal(hsl, want)
        ptr = hsl.ctypes.data
        assert np.array_equal(nphusl.to_rgb(hsl), img)            # pinned in, (small) result out
        hue = nphusl.to_hue(img)
        assert np.array_equal(hue, want[..., 0])
        del hsl
        gc.collect()
        again = nphusl.to_husl(img)                               # the block is recycled
    s1 = cu.pinned_pool_stats()
    assert again.ctypes.data == ptr and s1["hits"] > s0["hits"]
    assert np.array_equal(again, want)
    tiny = nphusl.to_husl(img[:4, :4])                            # small results stay pageable
    assert not pinned(tiny) and np.array_equal(tiny, want[:4, :4])


def test_device_results_come_from_a_recycling_pool(cu):
    """A call that returns a NEW DeviceArray must not pay cudaMalloc + cudaFree (0.6 ms and a device
    synchronisation): dropped results go back to a pool and the next result reuses the block."""
    import gc
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(23)
    img = rng.integers(0, 256, (540, 960, 3), dtype=np.uint8)
    t = torch.from_numpy(img).cuda()
    want = cu._rgb_to_husl(img, dtype=np.float32)
    a = cu._rgb_to_husl(t, dtype=np.float32)
    ptr = a.ptr
    assert np.array_equal(a.copy_to_host(), want)
    del a
    gc.collect()
    s0 = cu.device_pool_stats()
    b = cu._rgb_to_husl(t, dtype=np.float32)
    s1 = cu.device_pool_stats()
    assert b.ptr == ptr and s1["hits"] == s0["hits"] + 1 and s1["misses"] == s0["misses"]
    assert np.array_equal(b.copy_to_host(), want)
    c = cu._rgb_to_husl(t, dtype=np.float32)                      # b is still alive: a different block
    assert c.ptr != b.ptr and np.array_equal(c.copy_to_host(), want)
    hue = cu._rgb_to_hue(t)                                       # other size class
    assert np.array_equal(hue.copy_to_host(), cu._rgb_to_hue(img))
    view = b[10:20]                                               # views keep the owner alive, never release
    del b
    gc.collect()
    assert np.array_equal(view.copy_to_host(), want[10:20])
    del view, c, hue
    gc.collect()
    assert cu.device_pool_stats()["cached_bytes"] > 0
    cu.empty_cache()
    assert cu.device_pool_stats()["cached_bytes"] == 0
    assert np.array_equal(cu._rgb_to_husl(t, dtype=np.float32).copy_to_host(), want)


def test_past_the_reference_size_limit(cu):
    """The reference's C path indexes with `int size` (elements, _simd.c:128,195) and overflows past
    715 Mpx.  750 Mpx here: element offsets beyond 2^31 in every kernel family on the
    path (TMA / cp.async streaming kernels, hue, fused edit), exact round trip, spot checks vs the oracle."""
    torch = pytest.importorskip("torch")
    n = 750_000_128                                           # > 2^31 / 3 pixels, a multiple of the 128-px tile plus nothing
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~32 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(5)
    rgb = torch.randint(0, 256, (n, 3), dtype=torch.uint8, device="cuda", generator=g)
    back = torch.empty_like(rgb)
    for tdt, ndt in ((torch.float32, np.float32), (torch.float64, np.float64)):
        hsl = torch.empty((n, 3), dtype=tdt, device="cuda")
        cu._rgb_to_husl(rgb, out=hsl)
        cu._husl_to_rgb(hsl, out=back)
        torch.cuda.synchronize()
        assert torch.equal(back, rgb)
        for lo in (0, n // 2 + 77, n - 1000):                 # first, middle and the very last pixels
            want = orc.rgb_to_husl(rgb[lo:lo + 1000].cpu().numpy())
            got = hsl[lo:lo + 1000].cpu().numpy().astype(np.float64)
            assert hue_diff(got[:, 0], want[:, 0])[want[:, 1] > 0.1].max() < TOL
            assert np.abs(got[:, 1:] - want[:, 1:]).max() < TOL
        del hsl
    hue = torch.empty((n,), dtype=torch.float32, device="cuda")
    cu._rgb_to_hue(rgb, out=hue)
    torch.cuda.synchronize()
    want = orc.rgb_to_husl(rgb[n - 1000:].cpu().numpy())
    assert hue_diff(hue[n - 1000:].cpu().numpy().astype(np.float64), want[:, 0])[want[:, 1] > 0.1].max() < TOL
    del hue
    back.zero_()
    cu.edit_rgb(rgb, out=back)                                # identity edit: the fused kernel at the same offsets
    torch.cuda.synchronize()
    assert torch.equal(back, rgb)


def test_lut_mode_with_float_and_double_tables(cu, cube, golden_meta):
    """The reference generator's table element type (make_lookup_tables:4): float tables built ON
    THE DEVICE equal the reference generator's, and LUT / LUT+interp modes with them are bit-exact
    against the reference C build that bakes them in (sha256 over the 2^24 cube); double tables and
    another int scale against the oracle."""
    cu.lut_generate(1024, 1024, 1024, table_type="float")
    t = cu.lut_download()
    assert t["table_type"] == "float" and t["light"].dtype == np.float32 and t["light_scale"] == 1.0
    lt, st, th = orc.light_table(1024, "float")
    ct, hs, ls = orc.chroma_table(1024, 1024, "float")
    assert hashlib.sha256(t["light"].tobytes()).hexdigest() == golden_meta["light_table_float_sha256"]
    assert np.array_equal(t["y_idx_step"], st) and np.array_equal(t["y_thresh"], th)
    # device pow / sin / cos vs glibc: a handful of entries may sit on the other side of a decimal
    assert np.abs(t["chroma"] - ct).max() <= 0.0011 and np.count_nonzero(t["chroma"] != ct) <= 64
    cu.lut_upload(lt, st, th, ct, hs, ls, orc.linear_table6())            # the reference's exact tables
    for mode, key in (("lut", "lutf_strict_cube_sha256"), ("lut_interp", "lutf_interp_strict_cube_sha256")):
        got = cu._rgb_to_husl(cube, mode=mode)
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == golden_meta[key]
    rng = np.random.default_rng(24)
    rgb = rng.integers(0, 256, (60000, 3), dtype=np.uint8)
    # double tables, odd sizes
    ltd, std_, thd = orc.light_table(300, "double")
    ctd, hsd, lsd = orc.chroma_table(200, 260, "double")
    cu.lut_upload(ltd, std_, thd, ctd, hsd, lsd, orc.linear_table6())
    assert cu.lut_download()["table_type"] == "double"
    for mode in ("lut", "lut_interp"):
        want = orc.simd_rgb_to_husl(rgb, mode, tables=(ltd, std_, thd, ctd, hsd, lsd))
        assert np.array_equal(cu._rgb_to_husl(rgb, mode=mode).view(np.uint64), want.view(np.uint64))
    cu.lut_generate(300, 200, 260, table_type="double")
    td = cu.lut_download()
    assert np.abs(td["light"] - ltd).max() <= 1.1e-4 and np.abs(td["chroma"] - ctd).max() <= 1.1e-3
    # back to the default build for whoever comes next
    cu.lut_generate()
    assert cu.lut_download()["table_type"] == "uint16"


@pytest.mark.parametrize("shape,pitch_pad", [((2160, 3840), 0), ((270, 484), 28), ((66, 130), 0), ((7, 9), 3)])
def test_nv12_decoder_surfaces(cu, shape, pitch_pad):
    """NV12 as the load stage: uint8 RGB bit-exact against the numpy
    statement of the integer colour conversion, HUSL / hue equal to to_husl of that RGB bit
    for bit (and within 1e-3 of the oracle), fused edit equal to edit_rgb of that RGB;
    pitched planes, widths that are not a multiple of 128 / 4, odd sizes; nothing written
    past the outputs."""
    import torch
    h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    hc, wc = (h + 1) // 2, (w + 1) // 2
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    uv = rng.integers(0, 256, (hc, 2 * wc), dtype=np.uint8)
    ybuf = torch.zeros((h, w + pitch_pad), dtype=torch.uint8, device="cuda")
    uvbuf = torch.zeros((hc, 2 * wc + pitch_pad), dtype=torch.uint8, device="cuda")
    ybuf[:, :w] = torch.from_numpy(y).cuda()
    uvbuf[:, :2 * wc] = torch.from_numpy(uv).cuda()
    yd, uvd = ybuf[:, :w], uvbuf[:, :2 * wc]
    for matrix, full in (("bt709", False), ("bt601", True)):
        want_rgb = orc.nv12_to_rgb(y, uv, matrix, full)
        guard = torch.full((h * w * 3 + 256,), 0xA5, dtype=torch.uint8, device="cuda")
        out = guard[128:128 + h * w * 3].view(h, w, 3)
        cu.nv12_to_rgb(yd, uvd, out=out, matrix=matrix, full_range=full)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want_rgb)
        assert bool((guard[:128] == 0xA5).all()) and bool((guard[128 + h * w * 3:] == 0xA5).all())
        rgb_d = torch.from_numpy(want_rgb).cuda()
        for dt in (np.float32, np.float64):
            hsl = np.asarray(cu.nv12_to_husl(yd, uvd, dtype=dt, matrix=matrix, full_range=full))
            assert np.array_equal(hsl, np.asarray(cu._rgb_to_husl(rgb_d, dtype=dt)))
            hue = np.asarray(cu.nv12_to_hue(yd, uvd, dtype=dt, matrix=matrix, full_range=full))
            assert hue.shape == (h, w) and np.array_equal(hue, hsl[..., 0])
        check_husl(hsl, orc.rgb_to_husl(want_rgb), ~_grey(want_rgb))
        ed = dict(hue=(200, 320), lightness=(10, 90), l=(0.5, 5.0), s=(1.2, 0.0))
        got = np.asarray(cu.nv12_edit_rgb(yd, uvd, matrix=matrix, full_range=full, **ed))
        assert np.array_equal(got, np.asarray(cu.edit_rgb(rgb_d, **ed)))
    if h % 2 == 0 and w % 2 == 0 and pitch_pad == 0:
        # one NVDEC-style allocation: luma rows, then the chroma rows, same pitch
        surf = torch.cat([torch.from_numpy(y), torch.from_numpy(uv)]).cuda()
        assert np.array_equal(np.asarray(cu.nv12_to_rgb(surf)), orc.nv12_to_rgb(y, uv))


def test_nv12_argument_errors(cu):
    import torch
    y = torch.zeros((4, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        cu.nv12_to_rgb(y, torch.zeros((2, 6), dtype=torch.uint8, device="cuda"))      # chroma plane too narrow
    with pytest.raises(ValueError):
        cu.nv12_to_rgb(np.zeros((4, 8), np.uint8), np.zeros((2, 8), np.uint8))         # host planes
    with pytest.raises(ValueError):
        cu.nv12_to_husl(y, torch.zeros((2, 8), dtype=torch.uint8, device="cuda"), matrix="bt2020")


@pytest.mark.parametrize("n", [128 * 37 + 5, 1 << 20])
def test_tiled_float_rgb_kernels_equal_the_generic_ones(cu, n):
    """float RGB in [0,1] <-> HUSL (the Cython backend's own signatures, _cython_opt.pyx:22,177):
    the tiled kernels (aligned buffers) give bit for bit what the one-pixel-per-thread kernels
    give (a misaligned view of the same data forces those), and both sit within 1e-3 of the oracle."""
    import torch
    rng = np.random.default_rng(n)
    frgb = rng.random((n, 3))
    frgb[:64] = np.repeat(rng.random((64, 1)), 3, axis=1)          # greys
    frgb[64:70] = [[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1e-4, 0, 0]]
    want = orc.rgb_to_husl(frgb)
    sat = want[:, 1] > 0.5
    for tin in (torch.float32, torch.float64):
        pad = torch.zeros(3 * n + 16, dtype=tin, device="cuda")
        src = torch.from_numpy(frgb).to("cuda", tin)
        pad[1:3 * n + 1] = src.reshape(-1)
        off = pad[1:3 * n + 1].view(n, 3)                          # 4 / 8 bytes off a 16-byte boundary
        for tout in (np.float32, np.float64):
            a = np.asarray(cu._rgb_to_husl(src, dtype=tout))
            b = np.asarray(cu._rgb_to_husl(off, dtype=tout))
            assert np.array_equal(a, b)
            if tin == torch.float64:
                check_husl(a, want, sat)
            assert np.array_equal(np.asarray(cu._rgb_to_hue(src, dtype=tout)), np.asarray(cu._rgb_to_hue(off, dtype=tout)))
    hsl = np.stack([rng.uniform(0, 360, n), rng.uniform(0, 100, n), rng.uniform(0, 100, n)], -1)
    hsl[:4] = [[10, 50, 0], [10, 50, 100], [200, 0, 50], [359.9, 100, 99.995]]
    wantf = orc.husl_to_rgb_float(hsl)
    for tin in (torch.float32, torch.float64):
        pad = torch.zeros(3 * n + 16, dtype=tin, device="cuda")
        src = torch.from_numpy(hsl).to("cuda", tin)
        pad[1:3 * n + 1] = src.reshape(-1)
        off = pad[1:3 * n + 1].view(n, 3)
        for tout in (np.float32, np.float64):
            a = np.asarray(cu._husl_to_rgb(src, dtype=tout))
            b = np.asarray(cu._husl_to_rgb(off, dtype=tout))
            assert np.array_equal(a, b)
            if tin == torch.float64:
                assert np.abs(a - wantf).max() < 2e-5


@pytest.mark.parametrize("x0,x1", [(40, 300), (0, 384), (128, 132)])
def test_dense_row_views_take_the_row_tiled_kernels(cu, x0, x1):
    """Crops / pitched frames / batches of ROIs whose rows are dense and aligned go through
    the row-tiled kernels (husl_rows.cuh): bit for bit what the contiguous copy gives, for
    to_husl (f32 / f64), to_hue, to_rgb into a region of a larger canvas and the in-place
    fused edit; nothing outside the views is written."""
    import torch
    rng = np.random.default_rng(x0 * 1000 + x1)
    vid = rng.integers(0, 256, (3, 64, 400, 3), dtype=np.uint8)
    t = torch.from_numpy(vid).cuda()
    crop, hcrop = t[:, 5:61, x0:x1], np.ascontiguousarray(vid[:, 5:61, x0:x1])
    assert not crop.is_contiguous() or x1 - x0 == 400
    n0 = cu.launch_count()
    for dt in (np.float32, np.float64):
        got = np.asarray(cu._rgb_to_husl(crop, dtype=dt))
        assert np.array_equal(got, np.asarray(cu._rgb_to_husl(hcrop, dtype=dt)))
        assert np.array_equal(np.asarray(cu._rgb_to_hue(crop, dtype=dt)), got[..., 0])
        # strided out= : a region of a larger HUSL canvas
        canvas = torch.full((3, 64, 400, 3), -7.0, dtype=torch.float32 if dt == np.float32 else torch.float64, device="cuda")
        cu._rgb_to_husl(crop, out=canvas[:, 5:61, x0:x1])
        c = canvas.cpu().numpy()
        assert np.array_equal(c[:, 5:61, x0:x1], got)
        c[:, 5:61, x0:x1] = -7.0
        assert (c == -7.0).all()
        # and back: HUSL region -> uint8 region of a canvas
        hsl_full = torch.from_numpy(np.asarray(cu._rgb_to_husl(vid, dtype=dt))).cuda()
        out = torch.full((3, 64, 400, 3), 0xA5, dtype=torch.uint8, device="cuda")
        cu._husl_to_rgb(hsl_full[:, 5:61, x0:x1], out=out[:, 5:61, x0:x1])
        o = out.cpu().numpy()
        assert np.array_equal(o[:, 5:61, x0:x1], hcrop)
        o[:, 5:61, x0:x1] = 0xA5
        assert (o == 0xA5).all()
    assert cu.launch_count() - n0 >= 10
    # fused edit of the regions, in place, for two edit shapes
    for ed in (dict(lightness=(0, 50), l=(1, 20)), dict(hue=(250, 290), invert=True, l=(0.5, 0), s=(0.9, 1))):
        work = t.clone()
        roi = work[:, 5:61, x0:x1]
        cu.edit_rgb(roi, out=roi, **ed)
        w = work.cpu().numpy()
        assert np.array_equal(w[:, 5:61, x0:x1], np.asarray(cu.edit_rgb(hcrop, **ed)))
        w[:, 5:61, x0:x1] = vid[:, 5:61, x0:x1]
        assert np.array_equal(w, vid)
    # a pitched frame (row padding), one frame, rows = 1 .. many
    pitched = torch.zeros((37, 300 + 20, 3), dtype=torch.uint8, device="cuda")
    img = rng.integers(0, 256, (37, 300, 3), dtype=np.uint8)
    pitched[:, :300] = torch.from_numpy(img).cuda()
    assert np.array_equal(np.asarray(cu._rgb_to_husl(pitched[:, :300], dtype=np.float32)), np.asarray(cu._rgb_to_husl(img, dtype=np.float32)))
    assert np.array_equal(np.asarray(cu._rgb_to_husl(pitched[:1, :300], dtype=np.float32)), np.asarray(cu._rgb_to_husl(img[:1], dtype=np.float32)))


def test_rgba_and_bgra_memory_through_the_row_kernels(cu):
    """rgba[..., :3] (4 stored bytes per pixel, alpha ignored) and BGRA memory presented as RGB
    (channel stride -1) with aligned rows: one 128-bit load per lane, same values as the packed copy."""
    import torch
    rng = np.random.default_rng(44)
    rgba = rng.integers(0, 256, (2, 33, 260, 4), dtype=np.uint8)
    t = torch.from_numpy(rgba).cuda()
    packed = np.ascontiguousarray(rgba[..., :3])
    for dt in (np.float32, np.float64):
        want = np.asarray(cu._rgb_to_husl(packed, dtype=dt))
        assert np.array_equal(np.asarray(cu._rgb_to_husl(t[..., :3], dtype=dt)), want)
        assert np.array_equal(np.asarray(cu._rgb_to_hue(t[..., :3], dtype=dt)), want[..., 0])
        assert np.array_equal(np.asarray(cu._rgb_to_husl(t[:, 3:30, 8:200, :3], dtype=dt)), want[:, 3:30, 8:200])

        class Bgra:                               # the same memory read as B, G, R, A: R is byte 2, channels run backwards
            device = 0
            __cuda_array_interface__ = {"shape": (2, 33, 260, 3), "typestr": "|u1", "data": (t.data_ptr() + 2, False),
                                        "version": 3, "strides": (33 * 260 * 4, 260 * 4, 4, -1)}
        swapped = np.ascontiguousarray(rgba[..., 2::-1])
        assert np.array_equal(np.asarray(cu._rgb_to_husl(Bgra(), dtype=dt)), np.asarray(cu._rgb_to_husl(swapped, dtype=dt)))


def test_in_place_edit_of_rgba_and_bgra_frames(cu):
    """Fused edit of RGBA / BGRA frames in place through rgba[..., :3]: the colour bytes equal
    edit_rgb of the packed copy, every alpha byte is left as it was."""
    import torch
    rng = np.random.default_rng(45)
    rgba = rng.integers(0, 256, (2, 33, 260, 4), dtype=np.uint8)
    for ed in (dict(lightness=(0, 50), l=(1, 20)), dict(hue=(250, 290), invert=True, l=(0.5, 0)), dict(h=(1, 90), s=(0.5, 0))):
        t = torch.from_numpy(rgba).cuda()
        roi = t[..., :3]
        assert cu.edit_rgb(roi, out=roi, **ed) is roi
        got = t.cpu().numpy()
        assert np.array_equal(got[..., :3], np.asarray(cu.edit_rgb(np.ascontiguousarray(rgba[..., :3]), **ed)))
        assert np.array_equal(got[..., 3], rgba[..., 3])
        # a cropped region of the RGBA frames, in place: the rest of the frames untouched
        t = torch.from_numpy(rgba).cuda()
        roi = t[:, 4:30, 8:200, :3]
        cu.edit_rgb(roi, out=roi, **ed)
        got = t.cpu().numpy()
        assert np.array_equal(got[:, 4:30, 8:200, :3], np.asarray(cu.edit_rgb(np.ascontiguousarray(rgba[:, 4:30, 8:200, :3]), **ed)))
        got[:, 4:30, 8:200, :3] = rgba[:, 4:30, 8:200, :3]
        assert np.array_equal(got, rgba)
        # the same memory as BGRA: R is byte 2, channels run backwards
        t = torch.from_numpy(rgba).cuda()

        class Bgra:
            device = 0
            __cuda_array_interface__ = {"shape": (2, 33, 260, 3), "typestr": "|u1", "data": (t.data_ptr() + 2, False),
                                        "version": 3, "strides": (33 * 260 * 4, 260 * 4, 4, -1)}
        v = Bgra()
        cu.edit_rgb(v, out=v, **ed)
        got = t.cpu().numpy()
        want = np.asarray(cu.edit_rgb(np.ascontiguousarray(rgba[..., 2::-1]), **ed))
        assert np.array_equal(got[..., 2::-1], want) and np.array_equal(got[..., 3], rgba[..., 3])


def test_strided_view_fuzz(cu):
    """Random (frames, rows, cols) views over one device buffer -- aligned and unaligned offsets,
    padded and NEGATIVE row / frame strides (flipped images), 3- and 4-byte pixels, tiny and wide
    rows, single rows, many frames of few rows -- through to_husl, to_hue, to_rgb and the fused
    edit: always what the packed copy of the same pixels gives, and never a byte outside the view."""
    import torch
    rng = np.random.default_rng(2024)
    buf_np = rng.integers(0, 256, 1 << 22, dtype=np.uint8)
    buf = torch.from_numpy(buf_np).cuda()
    base = buf.data_ptr()

    def raw(ptr, shape, strides, typestr="|u1"):
        class V:
            device = 0
            __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": strides}
        return V()

    cases = 0
    for trial in range(60):
        pxb = int(rng.choice([3, 3, 4]))
        cols = int(rng.choice([4, 8, 12, 124, 128, 132, 256, 300, 301, 1000, 8192 * 2 + 4, 5]))
        rows = int(rng.choice([1, 2, 3, 7, 40]))
        frames = int(rng.choice([1, 1, 2, 5, 33]))
        if frames * rows * cols > 400000:
            frames, rows = 1, min(rows, 3)
        pad = int(rng.choice([0, 0, 4, 16, 1, 7]))
        rs = cols * pxb + pad
        fs = rows * rs + int(rng.choice([0, 0, 16, 3]))
        flip_rows, flip_frames = bool(rng.integers(0, 2)), bool(rng.integers(0, 4) == 0)
        start = int(rng.choice([0, 16, 4, 2, 1, 64]))
        span = frames * fs + 64
        if start + span > buf_np.size:
            continue
        # element (f, y, x, c) lives at start + f*fs + y*rs + x*pxb + c of the buffer
        idx = (start + np.arange(frames)[:, None, None, None] * fs + np.arange(rows)[None, :, None, None] * rs
               + np.arange(cols)[None, None, :, None] * pxb + np.arange(3)[None, None, None, :])
        if flip_rows:
            idx = idx[:, ::-1]
        if flip_frames:
            idx = idx[::-1]
        packed = np.ascontiguousarray(buf_np[idx])
        p0 = base + int(idx[0, 0, 0, 0])
        strides = (-fs if flip_frames else fs, -rs if flip_rows else rs, pxb, 1)
        v = raw(p0, (frames, rows, cols, 3), strides)
        want = np.asarray(cu._rgb_to_husl(packed, dtype=np.float32))
        assert np.array_equal(np.asarray(cu._rgb_to_husl(v, dtype=np.float32)), want), (trial, pxb, cols, rows, frames, pad, start)
        assert np.array_equal(np.asarray(cu._rgb_to_hue(v, dtype=np.float64)), np.asarray(cu._rgb_to_hue(packed)))
        # inverse into the same view of a canvas, and the fused edit in place: bytes outside the view stay
        canvas = torch.full((buf_np.size,), 0x5A, dtype=torch.uint8, device="cuda")
        cv = raw(canvas.data_ptr() + int(idx[0, 0, 0, 0]), (frames, rows, cols, 3), strides)
        cu._husl_to_rgb(torch.from_numpy(want).cuda(), out=cv)
        torch.cuda.synchronize()
        c = canvas.cpu().numpy()
        assert np.array_equal(c[idx], packed), (trial, "to_rgb")
        c[idx] = 0x5A
        assert (c == 0x5A).all(), (trial, "to_rgb wrote outside the view")
        work = buf.clone()
        wv = raw(work.data_ptr() + int(idx[0, 0, 0, 0]), (frames, rows, cols, 3), strides)
        ed = dict(lightness=(20, 80), l=(0.7, 3.0), h=(1, 30))
        cu.edit_rgb(wv, out=wv, **ed)
        torch.cuda.synchronize()
        w = work.cpu().numpy()
        assert np.array_equal(w[idx], np.asarray(cu.edit_rgb(packed, **ed))), (trial, "edit")
        w[idx] = buf_np[idx]
        assert np.array_equal(w, buf_np), (trial, "edit wrote outside the view")
        cases += 1
    assert cases >= 40


@pytest.mark.parametrize("shape,pad", [((2160, 3840), 0), ((66, 132), 12), ((10, 6), 2)])
def test_nv12_encode_side(cu, shape, pad):
    """uint8 RGB -> NV12 bit-exact against the numpy statement; the fused video loop NV12 -> edit -> NV12
    equals rgb_to_nv12(nv12_edit_rgb(...)) bit for bit, also in place and into pitched destination planes
    (whose padding stays untouched), and with a different output matrix / range."""
    import torch
    h, w = shape
    rng = np.random.default_rng(h + w)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    rgb_d = torch.from_numpy(rgb).cuda()
    for matrix, full in (("bt709", False), ("bt601", True)):
        wy, wuv = orc.rgb_to_nv12(rgb, matrix, full)
        ybuf = torch.full((h, w + pad), 0xA5, dtype=torch.uint8, device="cuda")
        uvbuf = torch.full((h // 2, w + pad), 0xA5, dtype=torch.uint8, device="cuda")
        gy, guv = cu.rgb_to_nv12(rgb_d, out_y=ybuf[:, :w], out_uv=uvbuf[:, :w], matrix=matrix, full_range=full)
        torch.cuda.synchronize()
        assert np.array_equal(ybuf.cpu().numpy()[:, :w], wy) and np.array_equal(uvbuf.cpu().numpy()[:, :w], wuv)
        assert bool((ybuf[:, w:] == 0xA5).all()) and bool((uvbuf[:, w:] == 0xA5).all())
    # the video loop
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    uv = rng.integers(0, 256, (h // 2, w), dtype=np.uint8)
    yd, uvd = torch.from_numpy(y).cuda(), torch.from_numpy(uv).cuda()
    for ed in (dict(l=(0.5, 0)), dict(hue=(250, 290), invert=True, l=(0.5, 0), s=(0.8, 0)), dict(lightness=(0, 62), l=(0, 0), h=(1, 40))):
        edited = cu.nv12_edit_rgb(yd, uvd, **ed)                                   # fused NV12 -> edit -> RGB
        wy, wuv = (np.asarray(a) for a in cu.rgb_to_nv12(edited))
        gy, guv = cu.nv12_edit_nv12(yd, uvd, **ed)
        assert np.array_equal(np.asarray(gy), wy) and np.array_equal(np.asarray(guv), wuv)
        oy, ouv = orc.rgb_to_nv12(np.asarray(edited))
        assert np.array_equal(wy, oy) and np.array_equal(wuv, ouv)
        # other output matrix / range, in place
        y2, uv2 = yd.clone(), uvd.clone()
        cu.nv12_edit_nv12(y2, uv2, out_y=y2, out_uv=uv2, out_matrix="bt601", out_full_range=True, **ed)
        oy, ouv = orc.rgb_to_nv12(np.asarray(edited), "bt601", True)
        assert np.array_equal(y2.cpu().numpy(), oy) and np.array_equal(uv2.cpu().numpy(), ouv)
    with pytest.raises(ValueError):
        cu.rgb_to_nv12(torch.zeros((5, 8, 3), dtype=torch.uint8, device="cuda"))    # odd height


def test_nv12_size_fuzz(cu):
    """Random surface sizes and pitches (vector path, partial tiles, odd sizes -> one-pixel kernels):
    NV12 -> RGB against the numpy statement, NV12 -> HUSL against to_husl of that RGB, and for even
    sizes the encode side against the numpy statement."""
    import torch
    rng = np.random.default_rng(77)
    for trial in range(40):
        w = int(rng.choice([2, 4, 6, 8, 36, 124, 128, 130, 132, 256, 260, 1000, 1001, 37]))
        h = int(rng.choice([1, 2, 3, 4, 10, 11, 64, 300]))
        pad_y, pad_uv = int(rng.choice([0, 0, 4, 12, 3])), int(rng.choice([0, 0, 4, 8, 1]))
        hc, wc2 = (h + 1) // 2, 2 * ((w + 1) // 2)
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        uv = rng.integers(0, 256, (hc, wc2), dtype=np.uint8)
        yb = torch.zeros((h, w + pad_y), dtype=torch.uint8, device="cuda")
        ub = torch.zeros((hc, wc2 + pad_uv), dtype=torch.uint8, device="cuda")
        yb[:, :w] = torch.from_numpy(y).cuda()
        ub[:, :wc2] = torch.from_numpy(uv).cuda()
        yd, ud = yb[:, :w], ub[:, :wc2]
        matrix, full = ("bt709", "bt601")[trial & 1], bool(trial & 2)
        want = orc.nv12_to_rgb(y, uv, matrix, full)
        got = np.asarray(cu.nv12_to_rgb(yd, ud, matrix=matrix, full_range=full))
        assert np.array_equal(got, want), (trial, w, h, pad_y, pad_uv)
        hsl = np.asarray(cu.nv12_to_husl(yd, ud, matrix=matrix, full_range=full))
        assert np.array_equal(hsl, np.asarray(cu._rgb_to_husl(want, dtype=np.float32))), (trial, w, h)
        if h % 2 == 0 and w % 2 == 0:
            gy, guv = cu.rgb_to_nv12(torch.from_numpy(want).cuda(), matrix=matrix, full_range=full)
            oy, ouv = orc.rgb_to_nv12(want, matrix, full)
            assert np.array_equal(np.asarray(gy), oy) and np.array_equal(np.asarray(guv), ouv), (trial, w, h)
            ed = dict(lightness=(10, 90), l=(0.8, 2.0))
            ey, euv = cu.nv12_edit_nv12(yd, ud, matrix=matrix, full_range=full, **ed)
            oy, ouv = orc.rgb_to_nv12(np.asarray(cu.edit_rgb(want, **ed)), matrix, full)
            assert np.array_equal(np.asarray(ey), oy) and np.array_equal(np.asarray(euv), ouv), (trial, w, h, "loop")


def test_concurrent_host_threads_on_one_device(cu):
    """SURVEY 8b threading: the shim is re-entrant per (device, stream).  Eight host threads hammer one
    GPU at once -- device arrays on their own streams, pageable host arrays (which share the per-device
    staging pipeline), pinned non-blocking calls, fused edits, LUT mode, new-array results from the
    recycling pools -- and every result is what the single-threaded call gives."""
    import threading
    import torch
    rng = np.random.default_rng(99)
    imgs = [rng.integers(0, 256, (97 + 13 * i, 211, 3), dtype=np.uint8) for i in range(8)]
    want_hsl = [np.asarray(cu._rgb_to_husl(a)) for a in imgs]
    want_edit = [np.asarray(cu.edit_rgb(a, lightness=(0, 50), l=(1, 20))) for a in imgs]
    cu.lut_generate()
    want_lut = [np.asarray(cu._rgb_to_husl(a, mode="lut")) for a in imgs]
    errors = []

    def worker(i):
        try:
            a = imgs[i]
            s = cu.Stream()
            d = torch.from_numpy(a).cuda()
            for rep in range(12):
                kind = (i + rep) % 5
                if kind == 0:                                   # device in, new device result, own stream
                    h = cu._rgb_to_husl(d, stream=s)
                    back = cu._husl_to_rgb(h, stream=s)
                    s.synchronize()
                    assert np.array_equal(np.asarray(h), want_hsl[i]) and np.array_equal(np.asarray(back), a)
                elif kind == 1:                                 # pageable host arrays through the staging pipeline
                    h = cu._rgb_to_husl(a)
                    assert np.array_equal(h, want_hsl[i]) and np.array_equal(cu._husl_to_rgb(h), a)
                elif kind == 2:                                 # pinned, non-blocking
                    pin, out = cu.pinned_empty(a.shape, np.uint8), cu.pinned_empty(a.shape, np.float64)
                    pin[...] = a
                    cu._rgb_to_husl(pin, out=out, stream=s, blocking=False)
                    s.synchronize()
                    assert np.array_equal(out, want_hsl[i])
                elif kind == 3:                                 # fused edit, host and device
                    assert np.array_equal(cu.edit_rgb(a, lightness=(0, 50), l=(1, 20)), want_edit[i])
                    e = cu.edit_rgb(d, stream=s, lightness=(0, 50), l=(1, 20))
                    s.synchronize()
                    assert np.array_equal(np.asarray(e), want_edit[i])
                else:                                           # LUT mode shares the per-device tables
                    assert np.array_equal(cu._rgb_to_husl(a, mode="lut"), want_lut[i])
            s.close()
        except BaseException as ex:                             # noqa: BLE001 -- reported by the main thread
            errors.append((i, repr(ex)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not any(t.is_alive() for t in threads), "a worker hung"
    assert not errors, errors


# ---- round 2: the edit inside the three-call pipeline, NV12 host planes, frame streams, pool ordering ----

README_EDIT = dict(hue=(250, 290), invert=True, l=(0.5, 0))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_to_rgb_applies_the_edit_on_load(cu, dtype):
    """to_rgb(hsl, edit=...) == to_rgb(edit_husl(hsl)) == the fused nphusl.edit(rgb) bit for bit, for host and
    device HUSL arrays, whole tiles, ragged tails and unaligned bases; `hsl` itself is not modified; and the
    result is within +-1 LSB of the float64 oracle doing the README edit (configs[3])."""
    torch = pytest.importorskip("torch")
    import nphusl
    rng = np.random.default_rng(31)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    for shape in ((5, 3), (131, 3), (4741, 3), (3, 360, 640, 3)):
        rgb = rng.integers(0, 256, shape, dtype=np.uint8)
        want = nphusl.edit(rgb, **README_EDIT)
        hsl = cu._rgb_to_husl(rgb, dtype=dtype)
        keep = hsl.copy()
        with nphusl.cuda_enabled():
            got = nphusl.to_rgb(hsl, edit=README_EDIT)                       # host array, dict form
        assert got.dtype == np.uint8 and np.array_equal(got, want) and np.array_equal(hsl, keep)
        d = torch.from_numpy(hsl).cuda()
        out = torch.empty(shape, dtype=torch.uint8, device="cuda")
        with nphusl.cuda_enabled():
            r = nphusl.to_rgb(d, out=out, edit=cu.Edit(**README_EDIT))       # device array, Edit form
        assert r is out and np.array_equal(out.cpu().numpy(), want) and np.array_equal(d.cpu().numpy(), keep)
        # an unaligned view of the same data goes through the generic kernel
        flat = torch.empty(hsl.size + 1, dtype=tdt, device="cuda")
        flat[1:].copy_(d.reshape(-1))
        un = cu._husl_to_rgb(flat[1:].view(*shape), edit=README_EDIT)
        assert np.array_equal(un.copy_to_host(), want)
        # the explicit pass over the array, in place and into another array
        e = cu.edit_husl(d.clone(), **README_EDIT)
        assert np.array_equal(cu._husl_to_rgb(e).copy_to_host(), want)
        o2 = torch.empty_like(d)
        assert cu.edit_husl(d, cu.Edit(**README_EDIT), out=o2) is o2 and torch.equal(o2, e) and np.array_equal(d.cpu().numpy(), keep)
        ed = e.cpu().numpy()
        sel = ~((keep[..., 0] >= 250) & (keep[..., 0] <= 290))
        assert np.array_equal(ed[..., :2], keep[..., :2])                    # H and S pass through bit for bit
        assert np.array_equal(ed[..., 2][~sel], keep[..., 2][~sel]) and np.array_equal(ed[..., 2][sel], keep[..., 2][sel] * dtype(0.5))
        # float RGB output takes the edit as well
        f = cu._husl_to_rgb(d, dtype=np.float32, edit=README_EDIT).copy_to_host()
        assert np.abs(np.round(f * 255.0) - want).max() <= 1
    o = orc.rgb_to_husl(rgb)
    ob = (o[..., 0] > 250) & (o[..., 0] < 290)                               # README.md:77-81 verbatim
    o[..., 2][~ob] *= 0.5
    od = np.abs(orc.husl_to_rgb(o).astype(int) - want.astype(int))
    edge = np.minimum(np.abs(o[..., 0] - 250), np.abs(o[..., 0] - 290)) < 1e-3
    assert od[~edge].max() <= 1 and (od[~edge] != 0).mean() < 2e-3
    with pytest.raises(TypeError):
        cu._husl_to_rgb(hsl, edit="darken")
    with pytest.raises(ValueError):
        cu.edit_husl(hsl, **README_EDIT)                                     # host arrays: use numpy, or to_rgb(edit=)


def test_nv12_edit_nv12_with_host_planes(cu):
    """NV12 planes in HOST memory (1.5 B/px over the link each way): equal to the device-resident call;
    pinned planes with blocking=False only enqueue."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(32)
    h, w = 360, 640
    y = rng.integers(16, 236, (h, w), dtype=np.uint8)
    uv = rng.integers(16, 241, (h // 2, w), dtype=np.uint8)
    wy, wuv = cu.nv12_edit_nv12(torch.from_numpy(y).cuda(), torch.from_numpy(uv).cuda(), **README_EDIT)
    wy, wuv = wy.copy_to_host(), wuv.copy_to_host()
    gy, guv = cu.nv12_edit_nv12(y, uv, **README_EDIT)                        # pageable in, pooled results out
    assert gy.shape == (h, w) and guv.shape == (h // 2, w) and np.array_equal(gy, wy) and np.array_equal(guv, wuv)
    py, puv = cu.pinned_empty((h, w), np.uint8), cu.pinned_empty((h // 2, w), np.uint8)
    oy, ouv = cu.pinned_empty((h, w), np.uint8), cu.pinned_empty((h // 2, w), np.uint8)
    py[...], puv[...] = y, uv
    oy[...], ouv[...] = 0, 0
    s = cu.Stream()
    ry, ruv = cu.nv12_edit_nv12(py, puv, cu.Edit(**README_EDIT), out_y=oy, out_uv=ouv, stream=s, blocking=False)
    s.synchronize()
    assert ry is oy and ruv is ouv and np.array_equal(oy, wy) and np.array_equal(ouv, wuv)
    with pytest.raises(ValueError):
        cu.nv12_edit_nv12(y, uv[:, :-2], **README_EDIT)
    with pytest.raises(ValueError):
        cu.nv12_edit_nv12(y, uv, out_y=np.empty((h, w + 2), np.uint8), out_uv=ouv, **README_EDIT)


def test_stream_frames_converts_a_cyclic_video_exactly(cu):
    """_shard.stream_frames (BASELINE configs[4] at toy size): frame i of the stream is pool[i % P]; every
    frame's round trip arrives in the host ring bit for bit, in order, whatever the batch / pool / shard
    arithmetic; with an edit the frames equal the fused edit of the source frame."""
    import nphusl
    from nphusl import _shard
    rng = np.random.default_rng(33)
    P, H, W = 5, 96, 256
    pool = cu.pinned_empty((P, H, W, 3), np.uint8)
    pool[...] = rng.integers(0, 256, pool.shape, dtype=np.uint8)
    ndev = min(2, cu.device_count())
    for n, batch, first, edit, schedule in ((23, 4, 0, None, "static"), (7, 3, 9, None, "static"), (11, 8, 2, cu.Edit(**README_EDIT), "static"),
                                            (29, 4, 3, None, "dynamic"), (9, 2, 0, cu.Edit(**README_EDIT), "dynamic")):
        seen = {}

        def sink(dev, i, frames):
            for j, f in enumerate(frames):
                seen[i + j] = f.copy()
        called = []
        res = _shard.stream_frames(pool, n, devices=list(range(ndev)), first_frame=first, batch=batch, dtype=np.float32,
                                   edit=edit, ready=lambda: called.append(1), on_batch=sink, schedule=schedule)
        assert called == [1] and res["frames"] == n and sorted(seen) == list(range(first, first + n))
        assert sum(v["frames"] for v in res["per_device"].values()) == n and res["seconds"] > 0
        for i, f in seen.items():
            want = pool[i % P] if edit is None else nphusl.edit(pool[i % P], **README_EDIT)
            assert np.array_equal(f, want), i
        for dev, (i, p, m, view) in res["last"].items():
            assert p == i % P and np.array_equal(view[0], seen[i])
    # a pageable pool works too (staged through the library's pinned slots), without a sink
    res = _shard.stream_frames(np.array(pool), 6, devices=[0], batch=4)
    i, p, m, view = res["last"][0]
    assert (i, p, m) == (5, 0, 1) and np.array_equal(view[0], pool[0])
    with pytest.raises(ValueError):
        _shard.stream_frames(pool[..., 0], 4)


def test_dropped_device_results_are_not_reused_under_a_running_kernel(cu):
    """`x = to_husl(a, stream=s1); y = to_rgb(x, stream=s1); del x; to_husl(b, stream=s2)` must not let
    the s2 call overwrite x's block while the s1 kernel still reads it.  The s1 stream is kept busy so the race
    window is wide; with the stream-ordered pool y is always the round trip of a."""
    torch = pytest.importorskip("torch")
    import gc
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(0, 256, (4, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
    b = torch.randint(0, 256, (4, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
    busy = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cu.empty_cache()
    w0 = cu.device_pool_stats()["cross_stream_waits"]
    for _ in range(8):
        with torch.cuda.stream(s1):
            for _ in range(6):
                busy.add_(1)                                          # ~0.1 ms each: s1 runs well behind the host
        x = cu._rgb_to_husl(a, dtype=np.float32, stream=s1)
        y = cu._husl_to_rgb(x, stream=s1)
        ptr = x.ptr
        del x
        gc.collect()
        z = cu._rgb_to_husl(b, dtype=np.float32, stream=s2)          # same size class: takes x's block ...
        assert z.ptr == ptr                                           # ... but only behind an event wait on s1's tail
        r = cu._husl_to_rgb(z, stream=s2)
        torch.cuda.synchronize()
        assert np.array_equal(y.copy_to_host(), a.cpu().numpy())
        assert np.array_equal(r.copy_to_host(), b.cpu().numpy())
        del y, z, r
        gc.collect()
    assert cu.device_pool_stats()["cross_stream_waits"] > w0


def test_grayscale_hue_stream_kernel_equals_the_general_kernel(cu):
    """to_hue of a grayscale image (config 3) runs k_gray_hue: one constant per lit pixel.  It must answer exactly
    what the general kernels answer for the same image materialised as R = G = B: uint8 / float32 / float64 greys
    incl. 0, -0, negatives, NaN and the values around the linear / power switch (tiled float kernel + tail), and
    -- against the one-pixel-per-thread kernel, whose code it abbreviates -- float-denormal and overflowing inputs."""
    import warnings
    import nphusl
    rng = np.random.default_rng(34)
    ordinary = np.array([0.0, -0.0, 1e-30, 0.04045, np.nextafter(0.04045, 1), 0.5, 1.0, 2.0, 1e6, -0.25, np.nan])
    exotic = np.array([1e-320, 1e-45, 1e-39, 5e-38, np.inf, -np.inf, 1e14, 1e16, 1e30, 3.5e38, 3.7e38, 1e39, 1e300])
    for dt in (np.float64, np.float32, np.uint8):
        with np.errstate(over="ignore", invalid="ignore"):
            grey = (rng.integers(0, 256, (37, 129)) if dt == np.uint8 else rng.random((37, 129))).astype(dt)
            if dt != np.uint8:
                grey.reshape(-1)[:ordinary.size] = ordinary.astype(dt)
        g3 = np.ascontiguousarray(np.repeat(grey[..., None], 3, -1))
        for odt in (np.float64, np.float32):
            want = cu._rgb_to_hue(g3, dtype=odt)                                       # 3 stored channels: general kernels
            got = cu._rgb_to_hue(np.broadcast_to(grey[..., None], grey.shape + (3,)), dtype=odt)   # 0-stride view: 1 channel
            assert got.shape == grey.shape and got.dtype == odt
            bad = np.flatnonzero(got.reshape(-1) != want.reshape(-1))
            assert bad.size == 0, (dt, odt, grey.reshape(-1)[bad][:8], got.reshape(-1)[bad][:8], want.reshape(-1)[bad][:8])
            if dt != np.uint8:
                with np.errstate(over="ignore"):
                    ex = exotic.astype(dt)                                             # 13 pixels: the one-pixel-per-thread kernel
                w = cu._rgb_to_hue(np.ascontiguousarray(np.repeat(ex[:, None], 3, -1)), dtype=odt)
                g = cu._rgb_to_hue(np.broadcast_to(ex[:, None], ex.shape + (3,)), dtype=odt)
                assert np.array_equal(g, w, equal_nan=True), (dt, odt, ex, g, w)
        if dt != np.uint8:
            with warnings.catch_warnings(), nphusl.cuda_enabled():
                warnings.simplefilter("ignore")
                hue = nphusl.to_hue(grey)                                                # public grayscale route
            assert np.array_equal(hue, cu._rgb_to_hue(g3))
            lit = hue[np.isfinite(grey) & (grey > 1e-30) & (grey < 1e7)]
            assert np.all(np.abs(lit - 19.916405993809086) < 1e-3)


def test_graph_api_records_and_replays_a_frame_loop(cu):
    """nphusl._cuda.Graph / graph(): the per-frame loop to_husl -> to_rgb(edit) recorded once through the C ABI
    (no torch involved) and replayed with one call per frame; results equal the direct calls."""
    import nphusl
    rng = np.random.default_rng(35)
    frames = rng.integers(0, 256, (5, 270, 480, 3), dtype=np.uint8)
    cur = cu.DeviceArray(frames.shape[1:], np.uint8)
    hsl = cu.DeviceArray(frames.shape[1:], np.float32)
    out = cu.DeviceArray(frames.shape[1:], np.uint8)
    edit = cu.Edit(**README_EDIT)
    g = cu.Graph()
    n0 = cu.launch_count()
    with g.capture() as s, nphusl.cuda_enabled():
        nphusl.to_husl(cur, out=hsl, stream=s)
        nphusl.to_rgb(hsl, out=out, stream=s, edit=edit)
    assert cu.launch_count() - n0 == 4                          # recorded (tiles + tail per call), counted once
    lib = cu.load()
    for f in frames:
        cu._check(lib.nphusl_cuda_memcpy(cur.ptr, f.ctypes.data, f.nbytes, 1, cur.device, cu._stream_ptr(g.stream)), "h2d")
        g.replay()
        g.synchronize()
        assert np.array_equal(out.copy_to_host(), nphusl.edit(f, **README_EDIT))
    hue = cu.DeviceArray(frames.shape[1:3], np.float32)
    g2 = cu.graph(lambda s: cu._rgb_to_hue(cur, out=hue, stream=s))
    g2()
    g2.synchronize()
    assert np.array_equal(hue.copy_to_host(), cu._rgb_to_hue(frames[-1], dtype=np.float32))
    with pytest.raises(cu.CudaError):
        cu.Graph().replay()
    g.close()
    g2.close()


def test_device_generated_tables_vs_the_reference_tables_on_the_cube(cu, cube):
    """`lut_generate()` + mode="lut" is NOT bit-exact against _simd.c -- the device's
    pow / sin / cos differ from glibc's in the last bit, so a chroma entry that sits on a .5 rounding boundary
    can come out one count (0.01 chroma) off.  This pins by how much: which table entries differ, and what
    that does to the result over all 2^24 colours relative to the reference's own tables (the bit-exact
    configuration of test_lut_mode_bit_exact_on_cube)."""
    lt, st, th = orc.light_table()
    ct, hs, ls = orc.chroma_table()
    cu.lut_upload(lt, st, th, ct, hs, ls, orc.linear_table6())
    ref = cu._rgb_to_husl(cube, mode="lut")                      # == _simd.c bit for bit
    cu.lut_generate(1024, 1024, 1024)
    t = cu.lut_download()
    got = cu._rgb_to_husl(cube, mode="lut")
    assert np.array_equal(t["light"], lt) and np.array_equal(t["linear"], orc.linear_table6())
    d = t["chroma"].astype(int) - ct.astype(int)
    n_entries = int(np.count_nonzero(d))
    assert n_entries <= 64 and np.abs(d).max() <= 1              # of 1 048 576 entries, one count each
    # hue and lightness never read the chroma table
    assert np.array_equal(got[..., 0].view(np.uint64), ref[..., 0].view(np.uint64))
    assert np.array_equal(got[..., 2].view(np.uint64), ref[..., 2].view(np.uint64))
    ds = np.abs(got[..., 1] - ref[..., 1])
    changed = int(np.count_nonzero(ds))
    # an entry is shared by the colours that round to the same (hue, lightness) cell: a few hundred each at most;
    # one count of max chroma moves S by S / (100 * max chroma): well below the LUT mode's own error (App. C: 1.7)
    assert changed <= 4000 * max(n_entries, 1) and (changed > 0) == (n_entries > 0)
    assert ds.max() <= 0.5, ds.max()
    print("device-generated chroma table: %d entries differ by one count; %d of 16 777 216 colours change S, max |dS| = %.4f"
          % (n_entries, changed, ds.max()))


# ---- the banded map of the reference's callers (examples.py:56-103) ----

def test_banded_map_hue_watermelon_and_melonize(cu):
    """nphusl._cuda.Bands: to_rgb(hsl, edit=Bands(...)) and the fused edit_rgb(rgb, edit=Bands(...)) equal
    to_rgb of the callers' numpy loops over the float32 HUSL array, bit for bit -- whole tiles, tails,
    host and device arrays, float32 and float64 HUSL; band edges, negative / out-of-range / NaN values
    belong to no band; `hsl` is left untouched."""
    torch = pytest.importorskip("torch")
    import nphusl
    from nphusl import chunk
    rng = np.random.default_rng(77)
    rgb = rng.integers(0, 256, (2, 270, 481, 3), dtype=np.uint8)
    hsl = cu._rgb_to_husl(rgb, dtype=np.float32)

    # hue_watermelon verbatim (examples.py:56-71), on the float32 HUSL array
    hue = hsl[..., 0]
    hue_out = hue.copy()
    pink, green, chunksize = 360, 130, 45
    for low, high in chunk(360, chunksize):
        select = np.logical_and(hue > low, hue < high)
        is_odd = low % (chunksize * 2)
        hue_out[select] = pink if is_odd else green
    melon = hsl.copy()
    melon[..., 0] = hue_out
    assert np.array_equal(melon, bands_spec(hsl, 0, 45, 360, 130, 360))
    want = cu._husl_to_rgb(melon)
    wm = cu.Bands("h", 45, 360, 130, 360)
    keep = hsl.copy()
    with nphusl.cuda_enabled():
        assert np.array_equal(nphusl.to_rgb(hsl, edit=wm), want) and np.array_equal(hsl, keep)
        assert np.array_equal(nphusl.edit(rgb, edit=wm), want)                    # fused, host arrays
    d_rgb = torch.from_numpy(rgb).cuda()
    out = torch.empty_like(d_rgb)
    assert cu.edit_rgb(d_rgb, edit=wm, out=out) is out and np.array_equal(out.cpu().numpy(), want)
    d64 = torch.from_numpy(hsl.astype(np.float64)).cuda()
    assert np.array_equal(cu._husl_to_rgb(d64, edit=wm).copy_to_host(), want)      # float64 HUSL (TMA ring)
    assert torch.equal(d64, torch.from_numpy(hsl.astype(np.float64)).cuda())

    # melonize (examples.py:74-103): lightness bands of every chunksize it walks through, stripe of +-2 at S = 100
    d32 = torch.from_numpy(hsl).cuda()
    for cs in (1, 2, 3, 7, 10, 33, 45, 64, 99, 100, 250):
        b = cu.Bands("l", cs, 100, 130, 360, stripe=(2, 100))
        spec = bands_spec(hsl, 2, cs, 100, 130, 360, (2, 100))
        if cs in (7, 100):      # the callers' loop verbatim
            lit = hsl[..., 2]
            o = hsl.copy()
            for low, high in chunk(100, cs):
                o[..., 0][np.logical_and(lit > low, lit < high)] = pink if low % (cs * 2) else green
                ave = (low + high) / 2
                o[..., 1][np.logical_and(lit > (ave - 2), lit < (ave + 2))] = 100
            assert np.array_equal(o, spec)
        w = cu._husl_to_rgb(spec)
        assert np.array_equal(cu._husl_to_rgb(d32, edit=b).copy_to_host(), w), cs
        assert np.array_equal(cu.edit_rgb(rgb, edit=b), w), cs
    # ragged sizes and an unaligned base (generic kernels)
    for n in (1, 5, 127, 129, 4741):
        sub = np.ascontiguousarray(hsl.reshape(-1, 3)[:n])
        b = cu.Bands("s", 12.5, 100, 10, 200, stripe=(1.5, 40))
        w = cu._husl_to_rgb(bands_spec(sub, 1, 12.5, 100, 10, 200, (1.5, 40)))
        assert np.array_equal(cu._husl_to_rgb(sub, edit=b), w)
        assert np.array_equal(cu.edit_rgb(np.ascontiguousarray(rgb.reshape(-1, 3)[:n]), edit=b), w)
        flat = torch.empty(3 * n + 1, dtype=torch.float32, device="cuda")
        flat[1:].copy_(torch.from_numpy(sub).reshape(-1))
        assert np.array_equal(cu._husl_to_rgb(flat[1:].view(n, 3), edit=b).copy_to_host(), w)
    # values on band edges, outside [0, total), NaN: passed on as they are
    edge = np.zeros((256, 3), np.float32)
    edge[:, 0] = np.concatenate([np.arange(0, 405, 45), np.arange(0, 405, 45) + 1e-3, np.arange(0, 405, 45) - 1e-3,
                                 [-10, -0.0, 359.99997, 360.00003, 720, np.nan, np.inf, -np.inf],
                                 rng.uniform(-50, 410, 256 - 35)]).astype(np.float32)
    edge[:, 1] = 80
    edge[:, 2] = rng.uniform(20, 80, 256).astype(np.float32)
    fin = np.isfinite(edge[:, 0])
    got = cu._husl_to_rgb(edge, edit=wm)
    assert np.array_equal(got[fin], cu._husl_to_rgb(bands_spec(edge, 0, 45, 360, 130, 360))[fin])
    # bad descriptors and entry points that take an Edit only
    with pytest.raises(ValueError):
        cu.Bands("x", 45)
    with pytest.raises(ValueError):
        cu.Bands("h", 0)
    with pytest.raises(cu.CudaError):
        cu._husl_to_rgb(hsl, edit=cu.Bands("h", 1e-4, 360))                     # more than 2^20 bands
    with pytest.raises(TypeError):
        cu.edit_husl(d32, edit=wm)
    with pytest.raises(cu.CudaError):
        cu._husl_to_rgb(hsl, dtype=np.float32, edit=wm)                         # uint8 RGB only


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_edit_husl_pass_group_sizes_alignment_and_guard_bands(cu, dtype):
    """The explicit edit pass (k_husl_edit_wide: 96-byte groups with 256-bit accesses + one-pixel-per-thread
    rest) against numpy on the same float32 arithmetic: every size around the group size, bases at every
    element offset (only 32-byte aligned ones take the wide kernel), in place and into another array, a
    full edit (hue wrap, saturation clamp, ranges); nothing outside the array is written."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(5)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    ed = dict(hue=(300, 40), saturation=(10, 95), lightness=(5, 90), h=(1.0, 77.5), s=(1.5, -5), l=(0.8, 3))
    f = np.float32

    def spec(a):
        h, s, l = (a[:, i].astype(f) for i in range(3))   # noqa: E741
        sel = ((h >= 300) | (h <= 40)) & (s >= 10) & (s <= 95) & (l >= 5) & (l <= 90)
        hn = h * f(1.0) + f(77.5)
        hn = hn - f(360) * np.floor(hn * f(1.0 / 360.0))
        hn = np.where((hn >= 360) | (hn < 0), f(0), hn)
        sn = np.clip(s * f(1.5) + f(-5), 0, 100).astype(f)
        ln = np.clip(l * f(0.8) + f(3), 0, 100).astype(f)
        out = a.copy()
        for i, (old, new) in enumerate(((h, hn), (s, sn), (l, ln))):
            new = np.where(sel, new, old)
            out[:, i] = np.where(new == old, a[:, i], new.astype(dtype))
        return out

    for n in (1, 3, 4, 7, 8, 9, 31, 32, 33, 257, 4099):
        for off in range(0, 9):
            a = np.stack([rng.uniform(0, 360, n), rng.uniform(0, 100, n), rng.uniform(0, 100, n)], axis=1).astype(dtype)
            want = spec(a)
            pad = 16
            buf = torch.full((pad + off + 3 * n + pad,), -7.0, dtype=tdt, device="cuda")
            view = buf[pad + off: pad + off + 3 * n].view(n, 3)
            view.copy_(torch.from_numpy(a))
            other = torch.full((pad + off + 3 * n + pad,), -9.0, dtype=tdt, device="cuda")
            oview = other[pad + off: pad + off + 3 * n].view(n, 3)
            assert cu.edit_husl(view, out=oview, **ed) is oview
            assert np.array_equal(oview.cpu().numpy(), want) and np.array_equal(view.cpu().numpy(), a)
            cu.edit_husl(view, **ed)                                        # in place
            assert np.array_equal(view.cpu().numpy(), want), (n, off)
            for b, fill in ((buf, -7.0), (other, -9.0)):
                g = b.cpu().numpy()
                assert np.all(g[:pad + off] == fill) and np.all(g[pad + off + 3 * n:] == fill), (n, off)


def test_float64_rgb_out_of_float64_husl_in_the_same_buffer(cu):
    """to_rgb(hsl, dtype=float64, out=hsl): the float-RGB result written over the HUSL array it came from (k_inv_flt_wide
    reads a thread's 96 bytes before it writes them; tails through the one-pixel-per-thread kernel) equals the
    out-of-place call, and both are within 2e-5 of the reference-made fixture values."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(91)
    for n in (3, 4, 127, 128, 4741, 270 * 480):
        hsl = np.stack([rng.uniform(0, 360, n), rng.uniform(0, 100, n), rng.uniform(0, 100, n)], 1)
        d = torch.from_numpy(hsl).cuda()
        want = cu._husl_to_rgb(d, dtype=np.float64).copy_to_host()
        r = cu._husl_to_rgb(d, dtype=np.float64, out=d)
        assert r is d and np.array_equal(d.cpu().numpy(), want)
        ref = orc.husl_to_rgb_float(hsl) if hasattr(orc, "husl_to_rgb_float") else None
        if ref is not None:
            assert np.abs(want - ref).max() < 2e-5
        assert want.min() >= 0.0 and want.max() <= 1.0
