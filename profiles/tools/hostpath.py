#!/usr/bin/env python
"""Host-array (pageable numpy) path: where does the time go?  Sweeps the
streaming chunk size under the current NPHUSL_COPY_THREADS, with a fresh
output array per call (page faults included, what a drop-in caller sees) and
with a preallocated `out=`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
from nphusl import _cuda  # noqa: E402


def best(fn, reps=7):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


rng = np.random.default_rng(0)
res = {"copy_threads": os.environ.get("NPHUSL_COPY_THREADS", "default")}
for name, shape in (("1080p", (1080, 1920, 3)), ("4k", (2160, 3840, 3))):
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    px = img.size // 3
    hsl = _cuda._rgb_to_husl(img)
    out_h = np.empty(shape, np.float64)
    out_r = np.empty(shape, np.uint8)
    for chunk in (0, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 23):
        _cuda.set_chunk_pixels(chunk)
        res["%s_chunk%d" % (name, chunk)] = {
            "to_husl_fresh": round(px / best(lambda: _cuda._rgb_to_husl(img)) / 1e6),
            "to_husl_out": round(px / best(lambda: _cuda._rgb_to_husl(img, out=out_h)) / 1e6),
            "to_rgb_fresh": round(px / best(lambda: _cuda._husl_to_rgb(hsl)) / 1e6),
            "to_rgb_out": round(px / best(lambda: _cuda._husl_to_rgb(hsl, out=out_r)) / 1e6),
            "to_husl_f32_fresh": round(px / best(lambda: _cuda._rgb_to_husl(img, dtype=np.float32)) / 1e6)}
    # numpy baselines: what allocating / copying a result of that size costs on this host
    res[name + "_np_empty_fill_f64_mpix"] = round(px / best(lambda: np.empty(shape, np.float64).fill(1.0)) / 1e6)
    res[name + "_np_copy_f64_mpix"] = round(px / best(lambda: hsl.copy()) / 1e6)
    res[name + "_np_copyto_f64_mpix"] = round(px / best(lambda: np.copyto(out_h, hsl)) / 1e6)
_cuda.set_chunk_pixels(0)
print(json.dumps(res))
