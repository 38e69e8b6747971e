#!/usr/bin/env python
"""Why does the to_husl -> to_rgb step differ between bench.py and ab.py?  Times the float64
pipeline under combinations of (buffer allocation order, per-kernel events, API layer)."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
import torch
import nphusl
from nphusl import _cuda
nphusl.enable_cuda()
B = 8
shape = (B, 2160, 3840, 3)
mode = sys.argv[1] if len(sys.argv) > 1 else "bench"
if mode == "bench":      # bench.py's allocation order
    host_in = _cuda.pinned_empty(shape, np.uint8)
    host_in[...] = np.random.default_rng(1000).integers(0, 256, shape, dtype=np.uint8)
    host_out = _cuda.pinned_empty(shape, np.uint8)
    rgb = torch.from_numpy(np.asarray(host_in)).cuda()
    hsl = torch.empty(shape, dtype=torch.float64, device="cuda")
    back = torch.empty_like(rgb)
else:                    # ab.py's
    g = torch.Generator(device="cuda").manual_seed(0)
    rgb = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)
    hsl = torch.empty(shape, dtype=torch.float64, device="cuda")
    h32 = torch.empty(shape, dtype=torch.float32, device="cuda")
    hue64 = torch.empty(shape[:-1], dtype=torch.float64, device="cuda")
    hue32 = torch.empty(shape[:-1], dtype=torch.float32, device="cuda")
    back = torch.empty_like(rgb)
print(mode, "ptrs rgb %x hsl %x back %x" % (rgb.data_ptr(), hsl.data_ptr(), back.data_ptr()))
stream = torch.cuda.current_stream()

def raw():
    _cuda._rgb_to_husl(rgb, out=hsl)
    _cuda._husl_to_rgb(hsl, out=back)

def api():
    nphusl.to_husl(rgb, out=hsl, stream=stream)
    nphusl.to_rgb(hsl, out=back, stream=stream)

def timed(fn, events, K=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(3 * K)]
    a.record()
    for k in range(K):
        if events:
            mids[3 * k].record(stream)
            _cuda._rgb_to_husl(rgb, out=hsl) if fn is raw else nphusl.to_husl(rgb, out=hsl, stream=stream)
            mids[3 * k + 1].record(stream)
            _cuda._husl_to_rgb(hsl, out=back) if fn is raw else nphusl.to_rgb(hsl, out=back, stream=stream)
            mids[3 * k + 2].record(stream)
        else:
            fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K

for name, fn in (("raw", raw), ("api", api)):
    for ev in (False, True):
        print("  %s events=%d  %.4f ms/step" % (name, ev, timed(fn, ev)))

# drift: per-20-step time over a long run, with SM clock / power sampled by NVML
import threading, time
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    h = None
samples = []
stop = False
def poll():
    while not stop and h is not None:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h),
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                        pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)))
        time.sleep(0.001)
th = threading.Thread(target=poll, daemon=True); th.start()
torch.cuda.synchronize()
K, G = 600, 20
evs = [torch.cuda.Event(enable_timing=True) for _ in range(K // G + 1)]
t0 = time.perf_counter()
evs[0].record()
for k in range(K):
    raw()
    if (k + 1) % G == 0:
        evs[(k + 1) // G].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
stop = True
print("  per-%d-step ms/step:" % G, " ".join("%.4f" % (evs[i].elapsed_time(evs[i + 1]) / G) for i in range(K // G)))
inwin = [s for s in samples if t0 <= s[0] <= t1]
if inwin and os.environ.get("PROBE_TRACE"):
    step = max(1, len(inwin) // 30)
    print("  trace (ms since start: sm MHz / mem MHz / W / reasons / C):",
          " | ".join("%d: %d/%d/%.0f/%s/%d" % ((s_[0] - t0) * 1e3, s_[1], s_[4], s_[2], hex(s_[3]), s_[5]) for s_ in inwin[::step]))
if inwin:
    print("  nvml in window: n=%d sm_mhz min/med/max %d/%d/%d power W min/max %.0f/%.0f reasons %s" % (
        len(inwin), min(s[1] for s in inwin), int(np.median([s[1] for s in inwin])), max(s[1] for s in inwin),
        min(s[2] for s in inwin), max(s[2] for s in inwin), sorted(set(hex(s[3]) for s in inwin))))
