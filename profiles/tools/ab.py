#!/usr/bin/env python
"""A/B timing of the streaming kernels under the current NPHUSL_STREAM /
NPHUSL_VARIANT environment: median-of-N CUDA-event time per launch on the bench
workload (8 device-resident 4K frames), with a parity check of each output."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
import torch  # noqa: E402
from nphusl import _cuda  # noqa: E402

frames = int(os.environ.get("AB_FRAMES", "8"))
g = torch.Generator(device="cuda").manual_seed(0)
rgb = torch.randint(0, 256, (frames, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
px = rgb.numel() // 3
h64 = torch.empty(rgb.shape, dtype=torch.float64, device="cuda")
h32 = torch.empty(rgb.shape, dtype=torch.float32, device="cuda")
hue64 = torch.empty(rgb.shape[:-1], dtype=torch.float64, device="cuda")
hue32 = torch.empty(rgb.shape[:-1], dtype=torch.float32, device="cuda")
back = torch.empty_like(rgb)


def timed(fn, reps=7, inner=20):
    """median over `reps` of (CUDA-event time of `inner` back-to-back launches) / inner:
    back-to-back so the host launch gap is hidden behind the previous kernel"""
    for _ in range(5):
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
    return float(np.median(ts))


res = {"stream": os.environ.get("NPHUSL_STREAM", "tma"), "variant": os.environ.get("NPHUSL_VARIANT", "0")}
cases = [("to_husl_f64", lambda: _cuda._rgb_to_husl(rgb, out=h64), 27),
         ("to_rgb_f64", lambda: _cuda._husl_to_rgb(h64, out=back), 27),
         ("to_husl_f32", lambda: _cuda._rgb_to_husl(rgb, out=h32), 15),
         ("to_rgb_f32", lambda: _cuda._husl_to_rgb(h32, out=back), 15),
         ("to_hue_f64", lambda: _cuda._rgb_to_hue(rgb, out=hue64), 11),
         ("to_hue_f32", lambda: _cuda._rgb_to_hue(rgb, out=hue32), 7)]
for name, fn, b in cases:
    ms = timed(fn)
    res[name] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 1), "hbm_frac": round(px * b / ms / 1e6 / 6547.2, 3)}
    if name.startswith("to_rgb"):
        torch.cuda.synchronize()
        res[name]["exact"] = bool(torch.equal(back, rgb))
        back.zero_()
def pipe64():
    _cuda._rgb_to_husl(rgb, out=h64)
    _cuda._husl_to_rgb(h64, out=back)


def pipe32():
    _cuda._rgb_to_husl(rgb, out=h32)
    _cuda._husl_to_rgb(h32, out=back)


pl32 = torch.empty((3,) + tuple(rgb.shape[:-1]), dtype=torch.float32, device="cuda")
pl64 = torch.empty((3,) + tuple(rgb.shape[:-1]), dtype=torch.float64, device="cuda")
_cuda.rgb_to_husl_planar(rgb, out=pl32)
_cuda.rgb_to_husl_planar(rgb, out=pl64)
for name, fn, b in (("pipeline_f64", pipe64, 54), ("pipeline_f32", pipe32, 30),
                    ("planar_fwd_f32", lambda: _cuda.rgb_to_husl_planar(rgb, out=pl32), 15),
                    ("planar_inv_f32", lambda: _cuda.husl_planar_to_rgb(pl32, out=back), 15),
                    ("planar_fwd_f64", lambda: _cuda.rgb_to_husl_planar(rgb, out=pl64), 27),
                    ("planar_inv_f64", lambda: _cuda.husl_planar_to_rgb(pl64, out=back), 27),
                    ("chw_fwd_f32", lambda: _cuda.rgb_to_husl_planar(rgb.view(3, -1), out=pl32.view(3, -1), channels_first=True), 15),
                    ("chw_inv_f32", lambda: _cuda.husl_planar_to_rgb(pl32.view(3, -1), out=back.view(3, -1), channels_first=True), 15),
                    ("fused_edit_u8", lambda: _cuda.edit_rgb(rgb, out=back, hue=(250, 290), invert=True, l=(0.5, 0)), 6),
                    ("fused_edit_chw_u8", lambda: _cuda.edit_rgb(rgb.view(3, -1), out=back.view(3, -1), channels_first=True,
                                                                hue=(250, 290), invert=True, l=(0.5, 0)), 6)):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 1), "hbm_frac": round(px * b / ms / 1e6 / 6547.2, 3)}
# float RGB inputs / float RGB outputs (generic one-pixel-per-thread kernels), 2 frames
f2 = rgb[:2]
px2 = f2.numel() // 3
rf32 = f2.to(torch.float32) / 255.0
rf64 = f2.to(torch.float64) / 255.0
o32 = torch.empty(f2.shape, dtype=torch.float32, device="cuda")
o64 = torch.empty(f2.shape, dtype=torch.float64, device="cuda")
g64 = rf64[..., 0].contiguous()
for name, fn, b in (("frgb32_to_husl32", lambda: _cuda._rgb_to_husl(rf32, out=o32), 24),
                    ("frgb64_to_husl64", lambda: _cuda._rgb_to_husl(rf64, out=o64), 48),
                    ("husl64_to_frgb64", lambda: _cuda._husl_to_rgb(o64, out=rf64), 48),
                    ("husl32_to_frgb32", lambda: _cuda._husl_to_rgb(o32, out=rf32), 24)):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 4), "gpix_s": round(px2 / ms / 1e6, 1), "hbm_frac": round(px2 * b / ms / 1e6 / 6547.2, 3)}
# strided device views (one-pixel-per-thread kernels): a crop of every frame, and RGBA seen as RGB
crop = rgb[:, 100:2060, 200:3640]
cpx = crop.numel() // 3
crop32 = torch.empty(crop.shape, dtype=torch.float32, device="cuda")
rgba = torch.empty(rgb.shape[:-1] + (4,), dtype=torch.uint8, device="cuda")
rgba[..., :3] = rgb
for name, fn, b, npx in (("strided_crop_to_husl32", lambda: _cuda._rgb_to_husl(crop, out=crop32), 15, cpx),
                         ("strided_crop_to_rgb", lambda: _cuda._husl_to_rgb(crop32, out=back[:, 100:2060, 200:3640]), 15, cpx),
                         ("strided_rgba3_to_husl32", lambda: _cuda._rgb_to_husl(rgba[..., :3], out=h32), 16, px),
                         ("strided_crop_edit", lambda: _cuda.edit_rgb(crop, out=crop, lightness=(0, 50), l=(1, 20)), 6, cpx),
                         ("rgba_edit_in_place", lambda: _cuda.edit_rgb(rgba[..., :3], out=rgba[..., :3], lightness=(0, 50), l=(1, 20)), 8, px)):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 4), "gpix_s": round(npx / ms / 1e6, 1), "hbm_frac": round(npx * b / ms / 1e6 / 6547.2, 3)}
# NV12 decoder surfaces as the load stage (8 frames stacked as one tall 4K-wide surface)
ysurf = torch.randint(0, 256, (frames * 2160, 3840), dtype=torch.uint8, device="cuda", generator=g)
uvsurf = torch.randint(0, 256, (frames * 1080, 3840), dtype=torch.uint8, device="cuda", generator=g)
ysurf2, uvsurf2 = torch.empty_like(ysurf), torch.empty_like(uvsurf)
hv32, hv64 = h32.view(frames * 2160, 3840, 3), h64.view(frames * 2160, 3840, 3)
bv = back.view(frames * 2160, 3840, 3)
for name, fn, b in (("nv12_to_husl_f32", lambda: _cuda.nv12_to_husl(ysurf, uvsurf, out=hv32), 13.5),
                    ("nv12_to_husl_f64", lambda: _cuda.nv12_to_husl(ysurf, uvsurf, out=hv64), 25.5),
                    ("nv12_to_hue_f32", lambda: _cuda.nv12_to_hue(ysurf, uvsurf, out=hue32.view(frames * 2160, 3840)), 5.5),
                    ("nv12_to_rgb", lambda: _cuda.nv12_to_rgb(ysurf, uvsurf, out=bv), 4.5),
                    ("nv12_edit_rgb", lambda: _cuda.nv12_edit_rgb(ysurf, uvsurf, out=bv, hue=(250, 290), invert=True, l=(0.5, 0)), 4.5),
                    ("nv12_edit_nv12", lambda: _cuda.nv12_edit_nv12(ysurf, uvsurf, out_y=ysurf2, out_uv=uvsurf2, hue=(250, 290), invert=True, l=(0.5, 0)), 3.0),
                    ("rgb_to_nv12", lambda: _cuda.rgb_to_nv12(bv, out_y=ysurf2, out_uv=uvsurf2), 4.5)):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 1), "hbm_frac": round(px * b / ms / 1e6 / 6547.2, 3)}
print(json.dumps(res))
