#!/usr/bin/env python
"""Pin the CPU oracle against the reference itself (run in the authoring
container, where /root/reference exists; results -> oracle/PIN_RESULTS.json).

Checks, over the full 2^24 RGB cube in tests/make_rgb16mil.py order (r slowest,
b fastest) unless noted:

 1. oracle float64 path  vs  the reference's NumPy path imported from
    /root/reference (nphusl.numpy_enabled): to_husl, to_hue, to_rgb;
 2. oracle float64 path  vs  the reference's compiled Cython kernels
    (oracle/_ref/refhost/_cython_opt, nphusl/_cython_opt.pyx);
 3. oracle _simd.c restatement (exact / lut / lut_interp)  vs  the reference's
    own _simd.c compiled strictly (oracle/_ref/libsimd_*_strict.so, gcc -O2) --
    must be BIT-EXACT -- and vs the -O3 -ffast-math default build (ulp report);
 4. oracle table generators vs the tables baked into the reference build
    (sha256 from SURVEY.md App. B);
 5. the reference tests' known-answer literals (tests/test.py:656-684).

TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle  # noqa: E402

REF = os.environ.get("REF", "/root/reference")


def cube():
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    return np.stack([r, g, b], axis=-1).reshape(4096, 4096, 3)


def hue_diff(a, b):
    d = np.abs(a - b)
    return np.minimum(d, 360.0 - d)


def import_reference():
    np.float = float   # SURVEY App. B fix 1 (aliases removed in numpy >= 1.24)
    np.int = int
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import nphusl
    assert os.path.dirname(nphusl.__file__).startswith(REF)
    return nphusl


def main():
    res = {"reference": REF, "numpy": np.__version__}
    rgb = cube()
    flat = rgb.reshape(-1, 3)
    grey = (flat[:, 0] == flat[:, 1]) & (flat[:, 1] == flat[:, 2])
    ref = import_reference()

    # ---- 1. NumPy path -------------------------------------------------
    t = time.time()
    o_hsl = oracle.rgb_to_husl(rgb).reshape(-1, 3)
    res["oracle_to_husl_s"] = round(time.time() - t, 2)
    t = time.time()
    with ref.numpy_enabled(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r_hsl = ref.to_husl(rgb).reshape(-1, 3)
    res["ref_numpy_to_husl_s"] = round(time.time() - t, 2)
    dh = hue_diff(o_hsl[:, 0], r_hsl[:, 0])
    res["numpy_to_husl"] = {
        "max_dH_nongrey": float(dh[~grey].max()),
        "max_dH_grey": float(dh[grey].max()),
        "max_dS": float(np.abs(o_hsl[:, 1] - r_hsl[:, 1]).max()),
        "max_dL": float(np.abs(o_hsl[:, 2] - r_hsl[:, 2]).max()),
    }
    with ref.numpy_enabled(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r_hue = ref.to_hue(rgb).reshape(-1)
    o_hue = oracle.rgb_to_hue(rgb).reshape(-1)
    res["numpy_to_hue"] = {"max_dH_nongrey": float(hue_diff(o_hue, r_hue)[~grey].max()),
                           "oracle_hue_equals_oracle_husl_H": bool(np.array_equal(o_hue, o_hsl[:, 0]))}
    with ref.numpy_enabled(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r_rgb = ref.to_rgb(r_hsl.reshape(rgb.shape))
    o_rgb = oracle.husl_to_rgb(r_hsl.reshape(rgb.shape))
    res["numpy_to_rgb"] = {
        "ref_roundtrip_exact": bool(np.array_equal(r_rgb, rgb)),
        "oracle_vs_ref_mismatches": int((o_rgb != r_rgb).sum()),
        "oracle_roundtrip_exact": bool(np.array_equal(oracle.husl_to_rgb(o_hsl.reshape(rgb.shape)), rgb)),
    }
    # float RGB out of the inverse (before the uint8 tail), subsample
    sub = r_hsl[::257]
    with ref.numpy_enabled(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r_f = ref.nphusl._husl_to_rgb(sub)
    res["numpy_husl_to_rgb_float"] = {"max_abs": float(np.abs(oracle.husl_to_rgb_float(sub) - r_f).max())}

    # ---- 2. Cython kernels ----------------------------------------------
    try:
        cy = oracle.ref_cython()
        f = flat.astype(np.float64) / 255.0
        c_hsl = np.asarray(cy._rgb_to_husl_2d(f))
        dh = hue_diff(o_hsl[:, 0], c_hsl[:, 0])
        res["cython_to_husl"] = {
            "max_dH_nongrey": float(dh[~grey].max()), "max_dH_grey": float(dh[grey].max()),
            "max_dS": float(np.abs(o_hsl[:, 1] - c_hsl[:, 1]).max()),
            "max_dL": float(np.abs(o_hsl[:, 2] - c_hsl[:, 2]).max())}
        c_rgbf = np.asarray(cy._husl_to_rgb_2d(np.ascontiguousarray(r_hsl)))
        o_rgbf = oracle.husl_to_rgb_float(r_hsl)
        res["cython_husl_to_rgb_float"] = {"max_abs": float(np.abs(c_rgbf - o_rgbf).max())}
        res["cython_max_chroma_0.25_40"] = [float(cy._test_max_chroma(0.25, 40.0)), oracle.max_chroma(0.25, 40.0)]
    except Exception as e:  # pragma: no cover
        res["cython"] = "unavailable: %r" % (e,)

    # ---- 3. _simd.c ------------------------------------------------------
    tabs = oracle.ref_tables()
    for variant, strict, fast in (("lut", "lut_strict", "lut"),
                                  ("lut_interp", "lut_interp_strict", "lut_interp"),
                                  ("exact", None, "exact")):
        mine = oracle.simd_rgb_to_husl(rgb, variant, tables=tabs).reshape(-1, 3)
        entry = {}
        if strict:
            theirs = oracle.ref_simd(rgb, strict).reshape(-1, 3)
            same = (mine.view(np.uint64) == theirs.view(np.uint64)) | (np.isnan(mine) & np.isnan(theirs))
            entry["bit_exact_vs_strict_O2"] = bool(same.all())
            entry["mismatching_values"] = int((~same).sum())
        theirs = oracle.ref_simd(rgb, fast).reshape(-1, 3)
        with np.errstate(invalid="ignore"):
            d = np.abs(mine - theirs)
        entry["vs_fastmath_O3"] = {"values_differing": int((mine.view(np.uint64) != theirs.view(np.uint64)).sum()),
                                   "max_abs": [float(np.nanmax(d[:, k])) for k in range(3)]}
        res["simd_" + variant] = entry

    # ---- 3b. _simd.c built with FLOAT lookup tables (make_lookup_tables float) --------
    try:
        ftabs = oracle.ref_tables("lutf_strict")
        flt, fst, fth = oracle.light_table(1024, "float")
        fct, fhs, fls = oracle.chroma_table(1024, 1024, "float")
        entry = {"light_equal_ref": bool(np.array_equal(flt, ftabs[0])),
                 "chroma_entries_differing": int((fct != ftabs[3]).sum())}
        for variant, build in (("lut", "lutf_strict"), ("lut_interp", "lutf_interp_strict")):
            mine = oracle.simd_rgb_to_husl(rgb, variant, tables=ftabs).reshape(-1, 3)
            theirs = oracle.ref_simd(rgb, build).reshape(-1, 3)
            same = (mine.view(np.uint64) == theirs.view(np.uint64)) | (np.isnan(mine) & np.isnan(theirs))
            entry[variant + "_mismatching_values_vs_strict_O2"] = int((~same).sum())
        res["simd_float_tables"] = entry
    except OSError as e:  # pragma: no cover
        res["simd_float_tables"] = "unavailable: %r" % (e,)

    # ---- 4. tables -------------------------------------------------------
    lt, st, th = oracle.light_table()
    ct, hs, ls = oracle.chroma_table()
    res["tables"] = {
        "light_sha256": hashlib.sha256(lt.tobytes()).hexdigest(),
        "light_equal_ref": bool(np.array_equal(lt, tabs[0])),
        "chroma_sha256": hashlib.sha256(ct.tobytes()).hexdigest(),
        "chroma_equal_ref": bool(np.array_equal(ct, tabs[3])),
        "chroma_entries_differing": int((ct != tabs[3]).sum()),
        "steps_equal_ref": bool(np.array_equal(st, tabs[1]) and np.array_equal(th, tabs[2])
                                and hs == tabs[4] and ls == tabs[5]),
        "linear6_equal_ref": bool(np.array_equal(oracle.linear_table6(), oracle.ref_linear_table())),
    }

    # ---- 5. literals (tests/test.py:656-684, _simd.c:79) ------------------
    res["literals"] = {
        "rgba_quadruplet": oracle.rgb_to_husl(np.array([0.5, 0.2, 0.0]) * 0.5).tolist(),  # ~[28.42153418, 100, 14.39285828]
        "white": oracle.rgb_to_husl(np.array([255, 255, 255], dtype=np.uint8)).tolist(),   # H = 19.916405993809086
        "to_rgb_360_100_100": oracle.husl_to_rgb(np.array([360.0, 100.0, 100.0])).tolist(),
        "red_hue": float(oracle.rgb_to_husl(np.array([255, 0, 0], dtype=np.uint8))[0]),
        "grey30_sat": float(oracle.rgb_to_husl(np.array([30, 30, 30], dtype=np.uint8))[1]),
        "grey_0.1_quantised": oracle.rgb_to_husl(np.array([26, 26, 26], dtype=np.uint8)).tolist(),  # L 9.30666400
    }
    out = os.path.join(HERE, "PIN_RESULTS.json")
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
