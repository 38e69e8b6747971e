"""to_husl u8 -> f64 stand-alone and inside the to_husl -> to_rgb(edit) step, for the current
NPHUSL_FWD64_DIRECT setting (A/B knob of nphusl_cuda_api.cu); grayscale f64 -> hue; 8 x 4K frames."""
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
    res = {"direct": os.environ.get("NPHUSL_FWD64_DIRECT", "0")}
    res["to_husl_f64_ms"] = round(timed(lambda: cu._rgb_to_husl(rgb, out=hsl)), 4)

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
    ref = cu._rgb_to_husl(rgb[:1].cpu().numpy())
    assert np.array_equal(hsl[:1].cpu().numpy(), ref)
    g = torch.rand(4000, 4000, dtype=torch.float64, device="cuda")
    ho = torch.empty_like(g)
    g3 = g[..., None].expand(-1, -1, 3)
    res["gray_f64_hue_ms"] = round(timed(lambda: cu._rgb_to_hue(g3, out=ho)), 4)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
