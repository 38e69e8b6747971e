#!/usr/bin/env python
"""Small, fast pass over every kernel family for compute-sanitizer (memcheck /
racecheck): ragged sizes so the streaming kernels, their tails and the generic
kernels all run; results checked against the round trip."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
from nphusl import _cuda  # noqa: E402

rng = np.random.default_rng(0)
for n in (1, 127, 128 * 40 + 3, 128 * 2000 + 77):
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    for dt in (np.float32, np.float64):
        hsl = _cuda._rgb_to_husl(rgb, dtype=dt)
        assert np.array_equal(_cuda._husl_to_rgb(hsl), rgb)
        hue = _cuda._rgb_to_hue(rgb, dtype=dt)
        assert np.array_equal(hue, hsl[:, 0])
        pl = _cuda.rgb_to_husl_planar(rgb, dtype=dt)
        assert np.array_equal(_cuda.husl_planar_to_rgb(pl), rgb)
        f = _cuda._rgb_to_husl(rgb.astype(dt) / 255)
        assert f.shape == rgb.shape
        assert _cuda._husl_to_rgb(hsl, dtype=np.float32).shape == rgb.shape
    assert np.array_equal(_cuda.edit_rgb(rgb), rgb)
_cuda.lut_generate(64, 64, 64)
_cuda._rgb_to_husl(rgb[:1000], mode="lut")
_cuda._rgb_to_husl(rgb[:1000], mode="lut_interp")
print("sanitize_small ok")
