import os, sys, numpy as np, torch
sys.path.insert(0, "husl-numpy_b200")
from nphusl import _cuda
g = torch.rand((4, 2160, 3840), dtype=torch.float64, device="cuda")
v = g.unsqueeze(-1).expand(-1, -1, -1, 3)
px = g.numel()
hue = torch.empty(g.shape, dtype=torch.float64, device="cuda")
hsl = torch.empty(g.shape + (3,), dtype=torch.float64, device="cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps
for name, fn, B in (("gray64_to_hue64", lambda: _cuda._rgb_to_hue(v, out=hue), 16), ("gray64_to_husl64", lambda: _cuda._rgb_to_husl(v, out=hsl), 32)):
    ms = timed(fn)
    print(name, round(ms, 4), "ms", round(px / ms / 1e6, 1), "Gpix/s hbm_frac", round(px * B / ms / 1e6 / 6547.2, 3))
