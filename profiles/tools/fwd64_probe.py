"""to_husl u8 -> f64 / to_rgb f64 -> u8 stand-alone and as the to_husl -> to_rgb(edit) step, for the current
setting of an A/B environment knob of nphusl_cuda_api.cu; 8 x 4K frames, CUDA events around the Python calls."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "husl-numpy_b200"))
from nphusl import _cuda as cu   # noqa: E402


def timed(fn, n=30):
    for _ in range(5):
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
    hsl = torch.empty(rgb.shape, dtype=torch.float64, device="cuda")
    out = torch.empty_like(rgb)
    ed = cu.Edit(hue=(250, 290), invert=True, l=(0.5, 0))
    res = {"lib": os.environ.get("NPHUSL_CUDA_LIB", "in-tree")}
    res["to_husl_f64_ms"] = round(timed(lambda: cu._rgb_to_husl(rgb, out=hsl)), 4)
    res["to_rgb_f64_ms"] = round(timed(lambda: cu._husl_to_rgb(hsl, out=out)), 4)
    res["to_rgb_f64_edit_ms"] = round(timed(lambda: cu._husl_to_rgb(hsl, out=out, edit=ed)), 4)

    def step():
        cu._rgb_to_husl(rgb, out=hsl)
        cu._husl_to_rgb(hsl, out=out, edit=ed)
    t = []
    for _ in range(5):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            step()
        b.record(); torch.cuda.synchronize()
        t.append(a.elapsed_time(b) / 20)
    res["step_ms"] = round(float(np.median(t)), 4)
    def step0():
        cu._rgb_to_husl(rgb, out=hsl)
        cu._husl_to_rgb(hsl, out=out)
    t = []
    for _ in range(5):
        for _ in range(3):
            step0()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            step0()
        b.record(); torch.cuda.synchronize()
        t.append(a.elapsed_time(b) / 20)
    res["step_no_edit_ms"] = round(float(np.median(t)), 4)
    assert torch.equal(out, rgb)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
