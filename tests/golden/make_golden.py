#!/usr/bin/env python
"""Generate tests/golden/*.npz|json with THE REFERENCE ITSELF (run in the
authoring container only: needs /root/reference and oracle/_ref).

The fixtures are what travels to the GPU box, where /root/reference does not
exist.  Everything here is produced by reference code:

* the NumPy float64 path imported from /root/reference (nphusl.numpy_enabled),
* its _simd.c compiled strictly (oracle/_ref/libsimd_lut_strict.so etc.),
* its generated lookup tables (baked into that .so).

Nothing in this script calls the oracle or the CUDA backend.
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as _o  # only for ref_simd / ref_tables (ctypes access to oracle/_ref)  # noqa: E402

REF = os.environ.get("REF", "/root/reference")


def import_reference():
    np.float = float
    np.int = int
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import nphusl
    assert os.path.dirname(nphusl.__file__).startswith(REF)
    return nphusl


def cube():
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    return np.stack([r, g, b], axis=-1).reshape(4096, 4096, 3)


def float_lut_fixtures(rgb, meta):
    """The float-table build of the reference (`make_lookup_tables float 1024 1024`,
    oracle/_ref/libsimd_lutf*_strict.so): sampled outputs, its tables, full-cube digests."""
    lt, st, th, ct, hs, ls = _o.ref_tables("lutf_strict")
    np.savez_compressed(
        os.path.join(HERE, "simd_float_golden.npz"),
        rgb=rgb,
        hsl_lutf_strict=_o.ref_simd(rgb, "lutf_strict"),
        hsl_lutf_interp_strict=_o.ref_simd(rgb, "lutf_interp_strict"),
        light_table=lt, y_idx_step=st, y_thresh=th, h_idx_step=hs, l_idx_step=ls,
        chroma_rows_every_64=ct[::64].copy())
    c = cube()
    meta["lutf_strict_cube_sha256"] = hashlib.sha256(_o.ref_simd(c, "lutf_strict").tobytes()).hexdigest()
    meta["lutf_interp_strict_cube_sha256"] = hashlib.sha256(_o.ref_simd(c, "lutf_interp_strict").tobytes()).hexdigest()
    meta["light_table_float_sha256"] = hashlib.sha256(lt.tobytes()).hexdigest()
    meta["chroma_table_float_sha256"] = hashlib.sha256(ct.tobytes()).hexdigest()


def main():
    if "--float-lut-only" in sys.argv:       # add the float-table fixtures to an existing set
        rgb = np.load(os.path.join(HERE, "simd_golden.npz"))["rgb"]
        with open(os.path.join(HERE, "golden_meta.json")) as f:
            meta = json.load(f)
        float_lut_fixtures(rgb, meta)
        with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(json.dumps({k: v for k, v in meta.items() if "float" in k or "lutf" in k}, indent=1))
        return
    ref = import_reference()
    rng = np.random.default_rng(20261019)
    warnings.simplefilter("ignore")

    # ---- sampled colours ---------------------------------------------------
    edge = [(0, 0, 0), (255, 255, 255), (255, 255, 254), (254, 255, 255), (255, 254, 255),
            (0, 0, 1), (1, 0, 0), (0, 1, 0), (1, 1, 1), (255, 0, 0), (0, 255, 0), (0, 0, 255),
            (255, 255, 0), (0, 255, 255), (255, 0, 255), (128, 128, 128), (128, 128, 129),
            (127, 128, 128), (30, 30, 30), (26, 26, 26), (10, 30, 40), (10, 10, 11), (11, 10, 10),
            (254, 254, 254), (254, 254, 255), (2, 1, 1), (200, 199, 200)]
    greys = [(i, i, i) for i in range(256)]
    rnd = rng.integers(0, 256, size=(16384 - len(edge) - len(greys), 3))
    rgb = np.array(edge + greys + rnd.tolist(), dtype=np.uint8)
    with ref.numpy_enabled():
        hsl = ref.to_husl(rgb)
        hue = ref.to_hue(rgb)
        back = ref.to_rgb(hsl)
    # arbitrary HUSL triplets (in gamut), incl. the L extremes and S = 0 / 100
    n = 8192
    hsl_in = np.stack([rng.uniform(0, 360, n), rng.uniform(0, 100, n), rng.uniform(0, 100, n)], -1)
    hsl_in[:64, 2] = rng.uniform(99.9, 100.0, 64)
    hsl_in[64:128, 2] = rng.uniform(0.0, 0.05, 64)
    hsl_in[128:192, 1] = 100.0
    hsl_in[192:256, 1] = 0.0
    hsl_in[256] = (360.0, 100.0, 100.0)
    hsl_in[257] = (0.0, 0.0, 0.0)
    hsl_in[258] = (0.0, 100.0, 50.0)
    with ref.numpy_enabled():
        rgb_f = ref.nphusl._husl_to_rgb(hsl_in)
        rgb_u8 = ref.to_rgb(hsl_in)
    # float RGB and grayscale inputs
    frgb = rng.uniform(0, 1, size=(4096, 3))
    frgb[:16] = rng.uniform(0, 0.002, size=(16, 3))      # below the L_MIN clamp
    frgb[16:32] = 1.0 - rng.uniform(0, 2e-5, size=(16, 3))  # above the L_MAX clamp
    grey = rng.uniform(0, 1, size=(64, 50))
    with ref.numpy_enabled():
        frgb_hsl = ref.to_husl(frgb)
        grey_hsl = ref.to_husl(grey)
        grey_hue = ref.to_hue(grey)
    np.savez_compressed(
        os.path.join(HERE, "husl_golden.npz"),
        rgb=rgb, hsl=hsl, hue=hue, rgb_back=back,
        hsl_in=hsl_in, rgb_float=rgb_f, rgb_u8=rgb_u8,
        frgb=frgb, frgb_hsl=frgb_hsl, grey=grey, grey_hsl=grey_hsl, grey_hue=grey_hue)

    # ---- _simd.c (LUT) on the same colours + tables -------------------------
    lt, st, th, ct, hs, ls = _o.ref_tables()
    np.savez_compressed(
        os.path.join(HERE, "simd_golden.npz"),
        rgb=rgb,
        hsl_lut_strict=_o.ref_simd(rgb, "lut_strict"),
        hsl_lut_interp_strict=_o.ref_simd(rgb, "lut_interp_strict"),
        hsl_exact_fast=_o.ref_simd(rgb, "exact"),
        light_table=lt, y_idx_step=st, y_thresh=th, h_idx_step=hs, l_idx_step=ls,
        chroma_rows_every_64=ct[::64].copy(), linear_table=_o.ref_linear_table())

    # ---- full-cube digests ---------------------------------------------------
    c = cube()
    with ref.numpy_enabled():
        ch = ref.to_husl(c)
    flat = ch.reshape(256, 65536, 3)
    hsl_lut = _o.ref_simd(c, "lut_strict")
    hsl_int = _o.ref_simd(c, "lut_interp_strict")
    meta = {
        "generated_by": "tests/golden/make_golden.py against " + REF,
        "cube_order": "r slowest, b fastest (tests/make_rgb16mil.py:8-17), 4096x4096x3",
        "numpy_cube_slice_sums_file": "cube_numpy_slices.npz",
        "lut_strict_cube_sha256": hashlib.sha256(np.ascontiguousarray(hsl_lut).tobytes()).hexdigest(),
        "lut_interp_strict_cube_sha256": hashlib.sha256(np.ascontiguousarray(hsl_int).tobytes()).hexdigest(),
        "light_table_sha256": hashlib.sha256(lt.tobytes()).hexdigest(),
        "chroma_table_sha256": hashlib.sha256(ct.tobytes()).hexdigest(),
        "numpy_cube_roundtrip_exact": bool(np.array_equal(ref.to_rgb(ch), c)),
        "literals": {
            "to_husl_rgba_0.5_0.2_0_0.5": [28.42153418, 100.0, 14.39285828],   # tests/test.py:671-673
            "white_hue": 19.916405993809086,                                     # _simd.c:79
            "to_rgb_360_100_100": [255, 255, 255],                               # tests/test.py:665-667
            "cython_max_chroma_0.25_40": 0.37678582700482516,                    # tests/test.py:559-563
        },
    }
    # per-r-slice sums and maxima of the NumPy float64 cube result (greys'
    # hue excluded: numerically meaningless)
    rgbf = c.reshape(256, 65536, 3)
    g = (rgbf[..., 0] == rgbf[..., 1]) & (rgbf[..., 1] == rgbf[..., 2])
    hmask = np.where(g, 0.0, flat[..., 0])
    np.savez_compressed(os.path.join(HERE, "cube_numpy_slices.npz"),
                        sum_h=hmask.sum(1), sum_s=flat[..., 1].sum(1), sum_l=flat[..., 2].sum(1),
                        max_s=flat[..., 1].max(1), min_l=flat[..., 2].min(1),
                        # every 4099th colour, full precision, for spot checks
                        sample_idx=np.arange(0, 1 << 24, 4099),
                        sample_hsl=ch.reshape(-1, 3)[::4099].copy())
    float_lut_fixtures(rgb, meta)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
