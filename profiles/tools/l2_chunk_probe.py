#!/usr/bin/env python
"""Cache-blocked to_husl -> to_rgb: the float64 HUSL intermediate of a row band that fits the
126 MB L2 is consumed by to_rgb before it is evicted.  Times the 8 x 4K batch pipeline for
several band sizes (MB of float64 HUSL per band), plain launches and a CUDA graph replay."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
import torch  # noqa: E402
from nphusl import _cuda  # noqa: E402

frames = 8
g = torch.Generator(device="cuda").manual_seed(0)
rgb = torch.randint(0, 256, (frames * 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
px = rgb.numel() // 3
back = torch.empty_like(rgb)
res = {}
for dt, name, bpp in ((torch.float64, "f64", 24), (torch.float32, "f32", 12)):
    hsl = torch.empty(rgb.shape, dtype=dt, device="cuda")
    rows_total = rgb.shape[0]
    for band_mb in (0, 16, 24, 32, 48, 64, 80, 96, 128, 200):
        rows = rows_total if band_mb == 0 else max(1, int(band_mb * 1e6 / (3840 * bpp)))
        bands = [(r, min(r + rows, rows_total)) for r in range(0, rows_total, rows)]

        def step():
            for a, b in bands:
                _cuda._rgb_to_husl(rgb[a:b], out=hsl[a:b])
                _cuda._husl_to_rgb(hsl[a:b], out=back[a:b])
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
        ms = float(np.median(ts))
        # graph replay of the same step (no host launch cost)
        s = torch.cuda.Stream()
        gms = None
        try:
            with torch.cuda.stream(s):
                def gstep():
                    for a, b in bands:
                        _cuda._rgb_to_husl(rgb[a:b], out=hsl[a:b], stream=s)
                        _cuda._husl_to_rgb(hsl[a:b], out=back[a:b], stream=s)
                gstep()
                s.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=s):
                    gstep()
                for _ in range(3):
                    graph.replay()
                s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(20):
                    graph.replay()
                e1.record(s)
                e1.synchronize()
                gms = e0.elapsed_time(e1) / 20
        except Exception as ex:          # noqa: BLE001
            gms = repr(ex)
        torch.cuda.synchronize()
        ok = bool(torch.equal(back, rgb))
        res["%s_band%d" % (name, band_mb)] = {"bands": len(bands), "ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 1),
                                             "graph_ms": gms if not isinstance(gms, float) else round(gms, 4), "exact": ok}
        print(name, band_mb, res["%s_band%d" % (name, band_mb)], flush=True)
    del hsl
print(json.dumps(res))
