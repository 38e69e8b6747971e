import sys, os
sys.path.insert(0, "husl-numpy_b200")
import torch, numpy as np
from nphusl import _cuda
g = torch.Generator(device="cuda").manual_seed(0)
rgb = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
h64 = torch.empty(rgb.shape, dtype=torch.float64, device="cuda")
_cuda.lut_generate()
for _ in range(3):
    _cuda._rgb_to_husl(rgb, out=h64, mode="lut")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    _cuda._rgb_to_husl(rgb, out=h64, mode="lut")
b.record(); torch.cuda.synchronize()
print("lut ms", a.elapsed_time(b) / 10)
