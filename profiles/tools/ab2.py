#!/usr/bin/env python
"""Round-2 A/B timing: median-of-N CUDA-event time per launch of the kernels worked on this round,
under the library named by NPHUSL_CUDA_LIB (build/var/lib_<name>.so from build_variants.sh).
One JSON line per run; AB_CASES=lut,edit,flt,... selects cases."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
import torch  # noqa: E402
from nphusl import _cuda  # noqa: E402

want = set(os.environ.get("AB_CASES", "lut,edit,flt,pipe").split(","))
g = torch.Generator(device="cuda").manual_seed(0)
rgb = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
px = rgb.numel() // 3
h64 = torch.empty(rgb.shape, dtype=torch.float64, device="cuda")
back = torch.empty_like(rgb)
edit = _cuda.Edit(hue=(250, 290), invert=True, l=(0.5, 0))


def timed(fn, reps=7, inner=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    return round(float(np.median(ts)), 4)


res = {"lib": os.path.basename(os.environ.get("NPHUSL_CUDA_LIB", "default"))}
if "lut" in want:
    _cuda.lut_generate()
    res["lut_ms"] = timed(lambda: _cuda._rgb_to_husl(rgb, out=h64, mode="lut"))
    res["lut_gpix"] = round(px / res["lut_ms"] / 1e6, 1)
    res["lut_interp_ms"] = timed(lambda: _cuda._rgb_to_husl(rgb, out=h64, mode="lut_interp"))
if "edit" in want or "pipe" in want:
    _cuda._rgb_to_husl(rgb, out=h64)
    res["inv64_ms"] = timed(lambda: _cuda._husl_to_rgb(h64, out=back))
    res["inv64_edit_ms"] = timed(lambda: _cuda._husl_to_rgb(h64, out=back, edit=edit))

    def pipe():
        _cuda._rgb_to_husl(rgb, out=h64)
        _cuda._husl_to_rgb(h64, out=back, edit=edit)
    res["pipe_edit_ms"] = timed(pipe)

    def pipe0():
        _cuda._rgb_to_husl(rgb, out=h64)
        _cuda._husl_to_rgb(h64, out=back)
    res["pipe_ms"] = timed(pipe0)
    h32 = torch.empty(rgb.shape, dtype=torch.float32, device="cuda")
    _cuda._rgb_to_husl(rgb, out=h32)
    res["inv32_ms"] = timed(lambda: _cuda._husl_to_rgb(h32, out=back))
    res["inv32_edit_ms"] = timed(lambda: _cuda._husl_to_rgb(h32, out=back, edit=edit))
    del h32
if "hue" in want:
    hue32 = torch.empty(rgb.shape[:-1], dtype=torch.float32, device="cuda")
    hue64 = torch.empty(rgb.shape[:-1], dtype=torch.float64, device="cuda")
    h32 = torch.empty(rgb.shape, dtype=torch.float32, device="cuda")
    res["hue32_ms"] = timed(lambda: _cuda._rgb_to_hue(rgb, out=hue32))
    res["hue64_ms"] = timed(lambda: _cuda._rgb_to_hue(rgb, out=hue64))
    res["fwd32_ms"] = timed(lambda: _cuda._rgb_to_husl(rgb, out=h32))
    res["fwd64_ms"] = timed(lambda: _cuda._rgb_to_husl(rgb, out=h64))
    res["fused_edit_ms"] = timed(lambda: _cuda.edit_rgb(rgb, edit, out=back))
    del hue32, hue64, h32
if "flt" in want:
    f64in = (rgb[:2].to(torch.float64) / 255.0).contiguous()
    f32in = f64in.to(torch.float32)
    for name, src, dt, tdt in (("f32_f32", f32in, np.float32, torch.float32), ("f32_f64", f32in, np.float64, torch.float64),
                               ("f64_f32", f64in, np.float32, torch.float32), ("f64_f64", f64in, np.float64, torch.float64)):
        out = torch.empty(src.shape, dtype=tdt, device="cuda")
        res["flt_" + name + "_ms"] = timed(lambda: _cuda._rgb_to_husl(src, out=out, dtype=dt))
    hue = torch.empty(f32in.shape[:-1], dtype=torch.float32, device="cuda")
    res["flt_f32_hue32_ms"] = timed(lambda: _cuda._rgb_to_hue(f32in, out=hue, dtype=np.float32))
print(json.dumps(res))
