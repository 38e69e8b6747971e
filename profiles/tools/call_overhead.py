#!/usr/bin/env python
"""Host-side cost of one public-API call on a device-resident frame (what bounds per-frame loops on
small images): wall time per call with the GPU kept busy-free (tiny image), and a cProfile of 2000 calls."""
import cProfile, os, pstats, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
import torch
import nphusl
from nphusl import _cuda
nphusl.enable_cuda()
img = torch.randint(0, 256, (64, 64, 3), dtype=torch.uint8, device="cuda")
hsl = torch.empty(img.shape, dtype=torch.float64, device="cuda")
out = torch.empty_like(img)
stream = torch.cuda.current_stream()
cases = {
    "nphusl.to_husl(dev, out=)": lambda: nphusl.to_husl(img, out=hsl),
    "nphusl.to_rgb(dev, out=)": lambda: nphusl.to_rgb(hsl, out=out),
    "nphusl.to_hue(dev)": lambda: nphusl.to_hue(img),
    "_cuda._rgb_to_husl(dev, out=)": lambda: _cuda._rgb_to_husl(img, out=hsl),
    "_cuda._husl_to_rgb(dev, out=)": lambda: _cuda._husl_to_rgb(hsl, out=out),
    "_cuda.edit_rgb(dev, out=)": lambda: _cuda.edit_rgb(img, out=out, lightness=(0, 50), l=(1, 20)),
    "nphusl.to_husl(dev) -> new DeviceArray": lambda: nphusl.to_husl(img),
}
for name, fn in cases.items():
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5000):
        fn()
    torch.cuda.synchronize()
    print("%-42s %6.1f us/call" % (name, (time.perf_counter() - t0) / 5000 * 1e6))
if len(sys.argv) > 1:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        nphusl.to_husl(img, out=hsl)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
