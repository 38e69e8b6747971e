#!/usr/bin/env python
"""Interleaved A/B of library variants in ONE process: every round times each variant's kernel
back to back, so clock drift (sw_power_cap on compute-bound kernels) hits all variants alike.
usage: ab_interleaved.py case lib1.so lib2.so ...   (case: edit | edit_l | edit_chw | hue32 | fwd32 | inv32)"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
from nphusl import _cuda

case, paths = sys.argv[1], sys.argv[2:]
NAMES = ("nphusl_cuda_rgb_edit_rgb", "nphusl_cuda_rgb_planes_edit", "nphusl_cuda_rgb_to_hue", "nphusl_cuda_rgb_to_husl",
         "nphusl_cuda_husl_to_rgb", "nphusl_cuda_nv12_to_husl", "nphusl_cuda_rgb_to_husl_strided",
         "nphusl_cuda_rgb_to_husl_planar", "nphusl_cuda_rgb_planes_to_husl_planes", "nphusl_cuda_nv12_edit_rgb")


def load(path):
    lib = ctypes.CDLL(path)
    for name in NAMES:
        res, args = _cuda.SIGNATURES[name]
        getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    return lib


libs = [load(p) for p in paths]
shape = (8, 2160, 3840, 3)
g = torch.Generator(device="cuda").manual_seed(0)
rgb = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)
back = torch.empty_like(rgb)
n = rgb.numel() // 3
h32 = torch.empty(shape, dtype=torch.float32, device="cuda")
hue32 = torch.empty(shape[:-1], dtype=torch.float32, device="cuda")
e_arc = _cuda.Edit(hue=(250, 290), invert=True, l=(0.5, 0))
e_l = _cuda.Edit(l=(0.5, 0))
libs[0].nphusl_cuda_rgb_to_husl(rgb.data_ptr(), 0, 1, 3, h32.data_ptr(), 1, 1, n, 0, 0, None)
ysurf = torch.randint(0, 256, (8 * 2160, 3840), dtype=torch.uint8, device="cuda", generator=g)
uvsurf = torch.randint(0, 256, (8 * 1080, 3840), dtype=torch.uint8, device="cuda", generator=g)
nv = _cuda.Nv12(ysurf.data_ptr(), uvsurf.data_ptr(), 3840, 3840, 3840, 8 * 2160, 1, 0)
# crop [:, 100:2060, 200:3640] of the batch -> contiguous float32 HUSL
crop_in = _cuda.View(rgb.data_ptr() + (100 * 3840 + 200) * 3, 0, 8, 1960, 3440, 2160 * 3840 * 3, 3840 * 3, 3, 1)
crop_out = _cuda.View(h32.data_ptr(), 1, 8, 1960, 3440, 1960 * 3440 * 12, 3440 * 12, 12, 4)


def call(lib):
    if case == "edit":
        rc = lib.nphusl_cuda_rgb_edit_rgb(rgb.data_ptr(), 1, back.data_ptr(), 1, n, ctypes.byref(e_arc), 0, None)
    elif case == "edit_l":
        rc = lib.nphusl_cuda_rgb_edit_rgb(rgb.data_ptr(), 1, back.data_ptr(), 1, n, ctypes.byref(e_l), 0, None)
    elif case == "edit_chw":
        p = rgb.data_ptr(); q = back.data_ptr()
        rc = lib.nphusl_cuda_rgb_planes_edit(p, p + n, p + 2 * n, q, q + n, q + 2 * n, n, ctypes.byref(e_arc), 0, None)
    elif case == "hue32":
        rc = lib.nphusl_cuda_rgb_to_hue(rgb.data_ptr(), 0, 1, 3, hue32.data_ptr(), 1, 1, n, 0, None)
    elif case == "fwd32":
        rc = lib.nphusl_cuda_rgb_to_husl(rgb.data_ptr(), 0, 1, 3, h32.data_ptr(), 1, 1, n, 0, 0, None)
    elif case == "inv32":
        rc = lib.nphusl_cuda_husl_to_rgb(h32.data_ptr(), 1, 1, back.data_ptr(), 0, 1, n, 0, None)
    elif case == "planar_fwd32":
        p = h32.data_ptr()
        rc = lib.nphusl_cuda_rgb_to_husl_planar(rgb.data_ptr(), p, p + 4 * n, p + 8 * n, 1, n, 0, None)
    elif case == "chw_fwd32":
        p, q = h32.data_ptr(), rgb.data_ptr()
        rc = lib.nphusl_cuda_rgb_planes_to_husl_planes(q, q + n, q + 2 * n, p, p + 4 * n, p + 8 * n, 1, n, 0, None)
    elif case == "nv12_edit":
        rc = lib.nphusl_cuda_nv12_edit_rgb(ctypes.byref(nv), ctypes.byref(e_arc), back.data_ptr(), 0, None)
    elif case == "nv12_f32":
        rc = lib.nphusl_cuda_nv12_to_husl(ctypes.byref(nv), h32.data_ptr(), 1, 0, 0, None)
    elif case == "crop_fwd32":
        rc = lib.nphusl_cuda_rgb_to_husl_strided(ctypes.byref(crop_in), ctypes.byref(crop_out), 0, 0, None)
    else:
        raise SystemExit("unknown case")
    assert rc == 0


for lib in libs:
    for _ in range(5):
        call(lib)
torch.cuda.synchronize()
ROUNDS, INNER = 15, 10
ts = [[] for _ in libs]
for r in range(ROUNDS):
    for i, lib in enumerate(libs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(INNER):
            call(lib)
        b.record()
        b.synchronize()
        ts[i].append(a.elapsed_time(b) / INNER)
for p, t in zip(paths, ts):
    print("%-8s %-28s median %.4f ms  min %.4f  max %.4f" % (case, os.path.basename(p), float(np.median(t)), min(t), max(t)))
