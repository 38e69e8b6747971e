"""numpy statements of the HUSL-space stages the callers of the reference write by hand (README.md:77-98,
examples.py:9-103), on float32 HUSL arrays -- what the fused / edit-on-load kernels are compared with on the GPU
(tests/test_cuda_parity.py) and what the host build of csrc/husl_edit.cuh is compared with on the CPU
(tests/test_emul_math.py)."""
import numpy as np


def bands_spec(hsl32, src, width, total, h_even, h_odd, stripe=None):
    """hue_watermelon / one frame of melonize (examples.py:56-103) on a float32 HUSL array, written the
    way the callers write it: for low, high in chunk(total, width): select = (x > low) & (x < high) ...
    Band edges and centres in float32 (exact for the integer parameters the callers use)."""
    f = np.float32
    x = hsl32[..., src].copy()
    out = hsl32.copy()
    k = 0
    while f(k) * f(width) < f(total):
        lo, hi = f(k) * f(width), min(f(k + 1) * f(width), f(total))
        out[..., 0][(x > lo) & (x < hi)] = f(h_odd if k & 1 else h_even)
        if stripe:
            c = f(0.5) * (lo + hi)
            out[..., 1][(x > c - f(stripe[0])) & (x < c + f(stripe[0]))] = f(stripe[1])
        k += 1
    return out


def edit_spec(hsl32, hue=(-np.inf, np.inf), saturation=(-np.inf, np.inf), lightness=(-np.inf, np.inf), invert=False,
              h=(1.0, 0.0), s=(1.0, 0.0), l=(1.0, 0.0)):     # noqa: E741
    """nphusl_edit (include/nphusl_cuda.h): select by hue arc / saturation / lightness ranges, affine edit of the selected
    pixels in float32; a channel whose (mul, add) is (1, 0) is left exactly as it is."""
    f = np.float32
    H, S, L = (hsl32[..., i].astype(f) for i in range(3))
    arc = ((H >= f(hue[0])) & (H <= f(hue[1]))) if hue[0] <= hue[1] else ((H >= f(hue[0])) | (H <= f(hue[1])))
    sel = arc & (S >= f(saturation[0])) & (S <= f(saturation[1])) & (L >= f(lightness[0])) & (L <= f(lightness[1]))
    if invert:
        sel = ~sel
    out = hsl32.copy()
    with np.errstate(invalid="ignore"):
        if tuple(h) != (1.0, 0.0):
            hn = H * f(h[0]) + f(h[1])
            hn = hn - f(360) * np.floor(hn * f(1.0 / 360.0))
            hn = np.where((hn >= 360) | (hn < 0), f(0), hn)
            out[..., 0] = np.where(sel, hn, H)
        if tuple(s) != (1.0, 0.0):
            out[..., 1] = np.where(sel, np.fmin(np.fmax(S * f(s[0]) + f(s[1]), f(0)), f(100)), S)
        if tuple(l) != (1.0, 0.0):
            out[..., 2] = np.where(sel, np.fmin(np.fmax(L * f(l[0]) + f(l[1]), f(0)), f(100)), L)
    return out
