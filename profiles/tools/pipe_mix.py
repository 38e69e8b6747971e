#!/usr/bin/env python
"""to_husl from library A, to_rgb from library B (same process, same buffers): which of the two
kernels decides the float64 pipeline step time?  usage: pipe_mix.py libA.so libB.so"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
from nphusl import _cuda

def load(path):
    lib = ctypes.CDLL(path)
    for name in ("nphusl_cuda_rgb_to_husl", "nphusl_cuda_husl_to_rgb"):
        res, args = _cuda.SIGNATURES[name]
        getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    return lib

A, B = load(sys.argv[1]), load(sys.argv[2])
shape = (8, 2160, 3840, 3)
g = torch.Generator(device="cuda").manual_seed(0)
rgb = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)
hsl = torch.empty(shape, dtype=torch.float64, device="cuda")
back = torch.empty_like(rgb)
n = rgb.numel() // 3

def step():
    assert A.nphusl_cuda_rgb_to_husl(rgb.data_ptr(), 0, 1, 3, hsl.data_ptr(), 2, 1, n, 0, 0, None) == 0
    assert B.nphusl_cuda_husl_to_rgb(hsl.data_ptr(), 2, 1, back.data_ptr(), 0, 1, n, 0, None) == 0

for _ in range(5):
    step()
torch.cuda.synchronize()
K, G = 600, 100
evs = [torch.cuda.Event(enable_timing=True) for _ in range(K // G + 1)]
evs[0].record()
for k in range(K):
    step()
    if (k + 1) % G == 0:
        evs[(k + 1) // G].record()
torch.cuda.synchronize()
assert torch.equal(back, rgb)
print(os.path.basename(sys.argv[1]), "+", os.path.basename(sys.argv[2]), " ".join("%.4f" % (evs[i].elapsed_time(evs[i + 1]) / G) for i in range(K // G)))
