#!/usr/bin/env python
"""Timeline of the reference's data flow (to_husl host -> HOST float64 HUSL, to_rgb HOST HUSL -> host) per 4K frame
with non-blocking calls on two alternating user streams: when does each call start and end (CUDA events on the user
streams), how much of the two transfer directions overlaps.  Measurement tool."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
import torch  # noqa: E402
import nphusl  # noqa: E402
from nphusl import _cuda  # noqa: E402

nphusl.enable_cuda()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(0)
frame = (2160, 3840, 3)
host_in = _cuda.pinned_empty((F,) + frame, np.uint8)
host_in[...] = rng.integers(0, 256, host_in.shape, dtype=np.uint8)
host_hsl = _cuda.pinned_empty((F,) + frame, np.float64)
host_out = _cuda.pinned_empty((F,) + frame, np.uint8)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(record):
    ev = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(F):
        s = streams[i & 1]
        if record:
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record(s)
        nphusl.to_husl(host_in[i], out=host_hsl[i], stream=s, blocking=False)
        if record:
            b.record(s)
        nphusl.to_rgb(host_hsl[i], out=host_out[i], stream=s, blocking=False)
        if record:
            c.record(s)
            ev.append((a, b, c))
    submit = time.perf_counter() - t0
    torch.cuda.synchronize()
    return time.perf_counter() - t0, submit, ev


run(False)
wall, submit, _ = run(False)
_, _, ev = run(True)
base = ev[0][0]
rows = [[round(base.elapsed_time(e), 2) for e in tr] for tr in ev]
px = F * frame[0] * frame[1]
print(json.dumps({"frames": F, "wall_ms": round(wall * 1e3, 2), "submit_ms": round(submit * 1e3, 2), "mpix_s": round(px / wall / 1e6, 1),
                  "round_trip_exact": bool(np.array_equal(host_out, host_in)),
                  "timeline_ms [call start, to_husl done, to_rgb done] per frame": rows}))
