#!/usr/bin/env python
"""Launch every streaming kernel of the hot path on the bench workload (8
device-resident 4K frames) a few times -- the command ncu wraps:

  python profiles/prof_kernels.py > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'k_(fwd|inv|edit|nv12|husl|gray|bands|to_nv12)' \
      -o gpurun_out/prof python profiles/prof_kernels.py
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
import torch  # noqa: E402
from nphusl import _cuda  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(0)
rgb = torch.randint(0, 256, (frames, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
h64 = torch.empty(rgb.shape, dtype=torch.float64, device="cuda")
h32 = torch.empty(rgb.shape, dtype=torch.float32, device="cuda")
hue64 = torch.empty(rgb.shape[:-1], dtype=torch.float64, device="cuda")
hue32 = torch.empty(rgb.shape[:-1], dtype=torch.float32, device="cuda")
back = torch.empty_like(rgb)
planes = torch.empty((3,) + tuple(rgb.shape[:-1]), dtype=torch.float32, device="cuda")
ysurf = torch.randint(0, 256, (frames * 2160, 3840), dtype=torch.uint8, device="cuda", generator=g)
uvsurf = torch.randint(0, 256, (frames * 1080, 3840), dtype=torch.uint8, device="cuda", generator=g)
f64in = (rgb[:2].to(torch.float64) / 255.0).contiguous()
f64out = torch.empty_like(f64in)
f32in, f32out = f64in.to(torch.float32), torch.empty(f64in.shape, dtype=torch.float32, device="cuda")
grey = torch.rand((2 * 2160, 3840), dtype=torch.float64, device="cuda", generator=g)
grey_hue = torch.empty_like(grey)
edit = _cuda.Edit(hue=(250, 290), invert=True, l=(0.5, 0))
_cuda.lut_generate()
import nphusl  # noqa: E402
import warnings  # noqa: E402
warnings.simplefilter("ignore")
for _ in range(reps):
    _cuda._rgb_to_husl(rgb, out=h64, mode="lut")  # the reference's default build, bit-exact LUT mode
    _cuda._rgb_to_husl(rgb, out=h64, mode="lut_interp")
    _cuda._rgb_to_husl(f32in, out=f32out, dtype=np.float32)   # float32 RGB -> float32 HUSL
    nphusl.to_hue(grey, out=grey_hue)             # grayscale float64 -> hue (config 3)
    _cuda._rgb_to_husl(rgb, out=h64)
    _cuda._husl_to_rgb(h64, out=back, edit=edit)  # to_rgb with the edit applied on load (bench headline step)
    _cuda._rgb_to_husl(rgb, out=h32)
    _cuda._husl_to_rgb(h32, out=back, edit=edit)
    _cuda.edit_husl(h64, edit)
    _cuda._rgb_to_husl(rgb, out=h64)
    _cuda._husl_to_rgb(h64, out=back)
    _cuda._rgb_to_husl(rgb, out=h32)
    _cuda._husl_to_rgb(h32, out=back)
    _cuda._rgb_to_hue(rgb, out=hue64)
    _cuda._rgb_to_hue(rgb, out=hue32)
    _cuda.rgb_to_husl_planar(rgb, out=planes)
    _cuda.husl_planar_to_rgb(planes, out=back)
    _cuda.nv12_to_husl(ysurf, uvsurf, out=h64.view(-1, 3840, 3))
    _cuda.nv12_to_husl(ysurf, uvsurf, out=h32.view(-1, 3840, 3))
    _cuda.nv12_edit_rgb(ysurf, uvsurf, out=back.view(-1, 3840, 3), l=(0.5, 0))
    _cuda._rgb_to_husl(f64in, out=f64out)         # float64 RGB in [0,1] (2 frames), tiled float kernels
    _cuda._husl_to_rgb(f64out, out=f64in, dtype=np.float64)
    _cuda._husl_to_rgb(h64, out=back, edit=_cuda.Bands("h", 45, 360, 130, 360))                       # hue_watermelon on load
    _cuda.edit_rgb(rgb, edit=_cuda.Bands("l", 7, 100, 130, 360, stripe=(2, 100)), out=back)           # a melonize frame, fused
    _cuda.edit_rgb(rgb, edit=edit, out=back)      # the README edit, fused
    _cuda.edit_rgb(rgb, out=back)                 # identity edit: exact round trip
torch.cuda.synchronize()
assert torch.equal(back, rgb)
print("ok", _cuda.launch_count(), "launches")
