"""Banded map (nphusl._cuda.Bands) on the bench workload: 8 x 4K frames, CUDA-event ms per launch."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "husl-numpy_b200"))
from nphusl import _cuda as cu   # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    rgb = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
    px = rgb.numel() // 3
    out = torch.empty_like(rgb)
    res = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        hsl = torch.empty(rgb.shape, dtype=dt, device="cuda")
        cu._rgb_to_husl(rgb, out=hsl, dtype=np.float64 if dt == torch.float64 else np.float32)
        for label, ed in (("none", None), ("readme_edit", cu.Edit(hue=(250, 290), invert=True, l=(0.5, 0))),
                          ("watermelon", cu.Bands("h", 45, 360, 130, 360)),
                          ("melonize", cu.Bands("l", 7, 100, 130, 360, stripe=(2, 100)))):
            ms = timed(lambda: cu._husl_to_rgb(hsl, out=out, edit=ed))
            res["to_rgb_%s_%s" % (name, label)] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 1)}
        del hsl
    for label, ed in (("readme_edit", cu.Edit(hue=(250, 290), invert=True, l=(0.5, 0))),
                      ("watermelon", cu.Bands("h", 45, 360, 130, 360)),
                      ("melonize", cu.Bands("l", 7, 100, 130, 360, stripe=(2, 100)))):
        ms = timed(lambda: cu.edit_rgb(rgb, edit=ed, out=out))
        res["fused_%s" % label] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 1)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
