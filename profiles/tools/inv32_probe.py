import json, os, sys
import numpy as np, torch
sys.path.insert(0, "husl-numpy_b200")
from nphusl import _cuda as cu
def many(fn, n=20, rounds=5):
    t = []
    for _ in range(rounds):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        t.append(a.elapsed_time(b) / n)
    return round(float(np.median(t)), 4)
rgb = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
hsl = torch.empty(rgb.shape, dtype=torch.float32, device="cuda")
out = torch.empty_like(rgb)
cu._rgb_to_husl(rgb, out=hsl, dtype=np.float32)
res = {"d32": os.environ.get("NPHUSL_INV32_DIRECT", "0")}
res["to_rgb_f32_ms"] = many(lambda: cu._husl_to_rgb(hsl, out=out))
assert torch.equal(out, rgb)
def step():
    cu._rgb_to_husl(rgb, out=hsl, dtype=np.float32)
    cu._husl_to_rgb(hsl, out=out)
res["step_f32_ms"] = many(step)
print(json.dumps(res))
