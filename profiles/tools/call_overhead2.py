#!/usr/bin/env python
"""Per-call cost of small device-resident frames through the public API, and what a recorded Graph buys
(1080p device-resident to_husl was 84 Gpix/s per call against 223 at batch size)."""
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
res = {}
edit = _cuda.Edit(hue=(250, 290), invert=True, l=(0.5, 0))


ts = torch.cuda.Stream()                 # everything runs and is timed on this stream (graphs are replayed onto it)
torch.cuda.set_stream(ts)


def rate(fn, px, reps=400):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(ts)
    for _ in range(reps):
        fn()
    b.record(ts)
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return {"us_per_call_device": round(ms * 1e3, 2), "us_per_call_host": round(host * 1e6, 2), "gpix_s": round(px / ms / 1e6, 1)}


for name, shape in (("1080p", (1080, 1920, 3)), ("4k", (2160, 3840, 3))):
    px = shape[0] * shape[1]
    frame = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda")
    back = torch.empty_like(frame)
    for dt, tdt in (("f64", torch.float64), ("f32", torch.float32)):
        hsl = torch.empty(shape, dtype=tdt, device="cuda")
        s = torch.cuda.current_stream()
        res["%s_to_husl_%s_public_api" % (name, dt)] = rate(lambda: nphusl.to_husl(frame, out=hsl, stream=ts), px)
        res["%s_to_husl_%s_backend_callable" % (name, dt)] = rate(lambda: _cuda._rgb_to_husl(frame, out=hsl, stream=ts), px)
        g = _cuda.graph(lambda st: _cuda._rgb_to_husl(frame, out=hsl, stream=st))
        res["%s_to_husl_%s_graph" % (name, dt)] = rate(lambda: g.replay(ts), px)

        def rt():
            nphusl.to_husl(frame, out=hsl, stream=ts)
            nphusl.to_rgb(hsl, out=back, stream=ts)
        res["%s_round_trip_%s_public_api" % (name, dt)] = rate(rt, px)

        def rec(st):
            nphusl.to_husl(frame, out=hsl, stream=st)
            nphusl.to_rgb(hsl, out=back, stream=st)
        g2 = _cuda.graph(rec)
        res["%s_round_trip_%s_graph" % (name, dt)] = rate(lambda: g2.replay(ts), px)
        torch.cuda.synchronize()
        assert torch.equal(back, frame)

        def rec_e(st):
            nphusl.to_husl(frame, out=hsl, stream=st)
            nphusl.to_rgb(hsl, out=back, stream=st, edit=edit)
        g3 = _cuda.graph(rec_e)
        res["%s_edit_pipeline_%s_graph" % (name, dt)] = rate(lambda: g3.replay(ts), px)
        # 8 frames per replay: the loop body of a batch of small frames as ONE graph
        g8 = _cuda.graph(lambda st: [rec(st) for _ in range(8)])
        r8 = rate(lambda: g8.replay(ts), 8 * px, reps=100)
        res["%s_round_trip_%s_graph_of_8" % (name, dt)] = r8
print(json.dumps(res, indent=1))
