import sys, numpy as np, torch
sys.path.insert(0, "husl-numpy_b200")
from nphusl import _cuda as cu
rgb = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
hsl = torch.empty(rgb.shape, dtype=torch.float32, device="cuda")
out = torch.empty_like(rgb)
cu._rgb_to_husl(rgb, out=hsl, dtype=np.float32)
for _ in range(3):
    cu._husl_to_rgb(hsl, out=out)
torch.cuda.synchronize()
assert torch.equal(out, rgb)
print("ok")
