#!/usr/bin/env python
"""Headline benchmark of the HUSL hot path (BASELINE.json: "Mpix/s to_husl &
to_rgb, 4K frame batches, 1/2/4/8 B200; % of roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference]

One STEP = one pass of the hot path over one batch of synthetic 3840x2160x3
uint8 frames: ``to_husl`` (uint8 RGB -> float64 HUSL, the reference's dtype)
followed by ``to_rgb`` (float64 HUSL -> uint8 RGB).  The metric counts every
pixel once per step (a pixel that went through both conversions).

* ``value``  : frames already resident in HBM, both kernels back to back on one
  stream, CUDA-event timing, max over ranks (weak scaling: every GPU owns its
  own batch, no data-path collective -- the path is embarrassingly parallel).
* ``e2e``    : the same step through the public API (``nphusl.to_husl`` /
  ``nphusl.to_rgb`` -> ctypes -> C ABI) with HOST buffers: uint8 frames start
  in pinned host memory and uint8 results end in pinned host memory, H2D and
  D2H inside the timed region (the HUSL intermediate stays on the GPU, as in
  README.md:77-98 style pipelines to_husl -> edit -> to_rgb).
* ``roofline``: dominant kernel's algorithmic bytes / its mean CUDA-event
  duration against the measured HBM copy peak (MEASURED_PEAKS.json).
* ``cpu_baseline`` / ``--impl reference``: the reference's own CPU code for the
  same step (``rgb_to_husl_nd`` of nphusl/_simd.c in its default LUT build +
  ``_cython_opt._husl_to_rgb`` + the uint8 tail of transform.py:16-20), built
  from /root/reference into oracle/_ref, all host cores, on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "husl-numpy_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FRAME = (2160, 3840, 3)
FRAME_PX = FRAME[0] * FRAME[1]
METRIC = "Mpix/s to_husl+to_rgb, 4K frame batches"
UNIT = "Mpix/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--frames", type=int, default=8, help="4K frames per GPU per step")
    ap.add_argument("--husl-dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    return ap.parse_args()


def make_frames(n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n,) + FRAME, dtype=np.uint8)


def config(args, n_gpus, frames):
    return {"workload": "configs[3]: batch of %d synthetic 3840x2160x3 uint8 frames per GPU, to_husl -> to_rgb "
                        "(lightness edit is a caller-side op, not timed)" % frames,
            "frames_per_gpu": frames, "pixels_per_step": frames * FRAME_PX * n_gpus,
            "husl_dtype": args.husl_dtype,
            "l2": "inputs larger than L2: %.0f MB read+written per kernel launch vs 126 MB L2"
                  % (frames * FRAME_PX * (3 + 3 * np.dtype(args.husl_dtype).itemsize) / 1e6),
            "precision": "float32 arithmetic on hi+lo float pairs, HUSL stored as %s (the reference's dtype is float64); "
                         "all 2^24 colours within 6.6e-5 of the reference float64 path (tests/test_cuda_parity.py; north_star "
                         "bar: 1e-3); the exact uint8 round trip of the batch is checked on every run" % args.husl_dtype,
            "parallelism": "frames sharded over %d GPU(s), one process per GPU, no collective" % n_gpus}


# ---------------------------------------------------------------------------
# reference CPU implementation (oracle/_ref, built from /root/reference)
# ---------------------------------------------------------------------------

class ReferenceCPU:
    """The reference's own compiled code for the step; see module docstring."""

    def __init__(self):
        import oracle as orc
        self.orc = orc
        self.kind = "reference" if orc.have_ref() else "port"
        self.cy = None
        if self.kind == "reference":
            try:
                self.cy = orc.ref_cython()
            except Exception:
                self.cy = None
        # all host threads the box has: torchrun exports OMP_NUM_THREADS=1, undo that
        # for this CPU leg (libgomp may already be loaded, so set the ICV as well)
        want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        os.environ["OMP_NUM_THREADS"] = str(want)
        try:
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(want)
        except OSError:
            pass
        try:
            self.cores = int(orc.lib().oracle_omp_threads())
        except Exception:
            self.cores = want

    def to_husl(self, rgb):
        if self.kind == "reference":
            return self.orc.ref_simd(rgb, "lut")            # _simd_opt.pyx:20-33 incl. its copy
        return self.orc.simd_rgb_to_husl(rgb, "lut")

    def to_rgb(self, hsl):
        if self.cy is not None:
            f = np.asarray(self.cy._husl_to_rgb(hsl))        # _cython_opt.pyx:177-192
            return np.abs(np.round(f * 255.0)).astype(np.uint8)   # transform.py:16-20
        return self.orc.husl_to_rgb(hsl)

    def step(self, rgb):
        return self.to_rgb(self.to_husl(rgb))

    def describe(self):
        if self.kind == "reference":
            return ("reference rgb_to_husl_nd (nphusl/_simd.c, default LUT build, -O3 -ffast-math -fopenmp) + "
                    + ("_cython_opt._husl_to_rgb (OpenMP) + numpy uint8 tail" if self.cy is not None
                       else "oracle husl_to_rgb (Cython module unavailable)"))
        return "oracle/husl_oracle.c restatement (reference build absent)"


def time_cpu(ref, budget_s, max_frames=1):
    """best-of-N of the step on one 4K frame (timeit.repeat(number=1) style,
    reference tests/performance_test.py:82-83) within ~budget_s seconds."""
    rgb = make_frames(max_frames, 12345)[0]
    ref.step(rgb[:64])                                       # load libraries, spin up OpenMP
    best, t_husl, t_rgb, spent, reps = float("inf"), None, None, 0.0, 0
    while reps < 2 or (spent < budget_s and reps < 50):
        t0 = time.perf_counter()
        hsl = ref.to_husl(rgb)
        t1 = time.perf_counter()
        ref.to_rgb(hsl)
        t2 = time.perf_counter()
        if t2 - t0 < best:
            best, t_husl, t_rgb = t2 - t0, t1 - t0, t2 - t1
        spent += t2 - t0
        reps += 1
    px = rgb.size // 3
    return {"value": px / best / 1e6, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
            "sample": "one 3840x2160 frame, best of %d repeats; %s" % (reps, ref.describe()),
            "to_husl_mpix_s": px / t_husl / 1e6, "to_rgb_mpix_s": px / t_rgb / 1e6}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = ReferenceCPU()
    rgb = make_frames(1, 12345)[0]
    px = rgb.size // 3
    for _ in range(max(1, min(args.warmup, 2))):
        ref.step(rgb)
    steps = max(1, min(args.steps, 8))                       # bounded: ~0.7 s per step per frame
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.step(rgb)
    dt = time.perf_counter() - t0
    val = px * steps / dt / 1e6
    cfg = config(args, 1, args.frames)
    cfg["reference_sample"] = "one 3840x2160 frame per step (bounded sample of the batch), steps capped at 8"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                             "sample": "one 3840x2160 frame per step, %d steps; %s" % (steps, ref.describe())},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
                 "hw_power_brake": nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                break
            time.sleep(0.001)

    def mark(self):
        """-> index of the next sample (to cut the value region out of the whole run)"""
        return len(self.samples)

    def stats(self, lo=0, hi=None):
        part = self.samples[lo:hi]
        return {"sm_mhz": float(np.median(part)) if part else None, "sm_mhz_min": min(part) if part else None,
                "samples": len(part)}

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ---------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------

def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_for(kernel):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def warp_inst_for(kernel):
    """warp instructions per 66 355 200-px launch from the committed ncu capture (None when not recorded)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("warp_inst", {}).get(kernel)
    except Exception:
        return None


def run_cuda(args):
    import torch
    import torch.distributed as dist
    import nphusl
    from nphusl import _cuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not _cuda.available():
        raise SystemExit("bench.py: libnphusl_cuda.so / GPU not available -- there is no CPU fallback")
    torch.cuda.set_device(local)
    _cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nphusl.enable_cuda()

    B, K, W = args.frames, args.steps, max(args.warmup, 3)
    hdt = np.dtype(args.husl_dtype)
    tdt = torch.float64 if hdt == np.float64 else torch.float32
    px = B * FRAME_PX
    bpp = 3 + 3 * hdt.itemsize                       # algorithmic bytes per pixel per kernel

    # inputs: pinned host frames (e2e) and a device-resident copy (value)
    host_in = _cuda.pinned_empty((B,) + FRAME, np.uint8)
    host_in[...] = make_frames(B, 1000 + rank)
    host_out = _cuda.pinned_empty((B,) + FRAME, np.uint8)
    dev_in = torch.from_numpy(np.asarray(host_in)).cuda()
    dev_hsl = torch.empty((B,) + FRAME, dtype=tdt, device="cuda")
    dev_out = torch.empty_like(dev_in)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
        nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)

    # ---- value: device-resident ------------------------------------------------
    for _ in range(W):
        step_dev()
    barrier()
    # The K steps are bracketed by one pair of events; every PROBE-th step additionally carries
    # events around each of its two kernels (the live per-launch durations of `roofline`).  An event
    # between two kernels costs ~3.5 us of lost launch overlap, so not every step is instrumented.
    PROBE = 4
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for k in range(0, K, PROBE)}
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    n0 = _cuda.launch_count()
    barrier()
    m0 = sampler.mark()
    t_beg.record(stream)
    for k in range(K):
        e = ev.get(k)
        if e:
            e[0].record(stream)
        nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
        if e:
            e[1].record(stream)
        nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)
        if e:
            e[2].record(stream)
    t_end.record(stream)
    torch.cuda.synchronize()
    m1 = sampler.mark()
    launches = _cuda.launch_count() - n0
    total_ms = t_beg.elapsed_time(t_end)
    t_fwd = float(np.mean([e[0].elapsed_time(e[1]) for e in ev.values()]))
    t_inv = float(np.mean([e[1].elapsed_time(e[2]) for e in ev.values()]))
    if not torch.equal(dev_out, dev_in):
        raise SystemExit("bench.py: round trip is not exact -- refusing to report a number")
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = px * world * K / (total_ms * 1e-3) / 1e6

    # ---- host link of THIS box (diagnostic beside e2e): plain pinned copies -----------
    def link_gbs():
        nbytes = host_in.nbytes
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        flat_in = torch.from_numpy(np.asarray(host_in).reshape(-1))
        flat_out = torch.from_numpy(np.asarray(host_out).reshape(-1))

        def run(h2d, d2h, reps=3):
            best = 1e9
            for _ in range(reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if h2d:
                    with torch.cuda.stream(s1):
                        scratch.copy_(flat_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        flat_out.copy_(dev_in.reshape(-1), non_blocking=True)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            return nbytes / best / 1e9
        return {"h2d": round(run(True, False), 1), "d2h": round(run(False, True), 1),
                "duplex_each": round(run(True, True), 1)}
    link = link_gbs()

    # ---- e2e: public API, host buffers ------------------------------------------
    # step k runs on user stream k%2 with its own HUSL buffer; calls are
    # non-blocking (pinned memory), so the H2D of step k+1 overlaps the D2H of step k
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    hsl2 = [dev_hsl, torch.empty_like(dev_hsl)]

    def step_e2e(k):
        s = streams[k & 1]
        nphusl.to_husl(host_in, out=hsl2[k & 1], stream=s, blocking=False)     # H2D + kernel, chunk-pipelined
        nphusl.to_rgb(hsl2[k & 1], out=host_out, stream=s, blocking=False)     # kernel + D2H
    # the plain way first: blocking calls, one after the other (reported beside the headline)
    def step_blocking():
        nphusl.to_husl(host_in, out=dev_hsl)
        nphusl.to_rgb(dev_hsl, out=host_out)
    for _ in range(2):
        step_blocking()
    torch.cuda.synchronize()
    tb = time.perf_counter()
    for _ in range(5):
        step_blocking()
    torch.cuda.synchronize()
    e2e_blocking = px * 5 / (time.perf_counter() - tb) / 1e6

    # the reference's own data flow: the HUSL array comes back to the HOST (to_husl returns a
    # float64 ndarray, 24 B/px) and goes up again for to_rgb -- 27 B/px over the link each way
    e2e_host_husl = None
    if rank == 0 and not args.no_variants:
        try:
            host_hsl = _cuda.pinned_empty((B,) + FRAME, hdt)
            def step_host_husl():
                nphusl.to_husl(host_in, out=host_hsl)
                nphusl.to_rgb(host_hsl, out=host_out)
            step_host_husl()
            torch.cuda.synchronize()
            th = time.perf_counter()
            for _ in range(3):
                step_host_husl()
            torch.cuda.synchronize()
            e2e_host_husl = px * 3 / (time.perf_counter() - th) / 1e6
            del host_hsl
        except Exception as e:                         # not enough pinned memory on this box
            e2e_host_husl = repr(e)

    Ke = max(4, min(K, 20))
    for k in range(4):
        step_e2e(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        step_e2e(k)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3         # wall clock around submit + drain of Ke steps
    clocks = sampler.stop()                           # sampled across the value and e2e timed regions
    clocks["value_region"] = sampler.stats(m0, m1)    # the device-resident K steps alone
    if not np.array_equal(np.asarray(host_out), np.asarray(host_in)):
        raise SystemExit("bench.py: e2e round trip is not exact")
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = px * world * Ke / (float(t.item()) * 1e-3) / 1e6

    # ---- per-kernel variants (rank 0, informational) -------------------------------
    variants = {}
    if rank == 0 and not args.no_variants:
        def timed(fn, reps=20):
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        hbm, _ = peaks()
        h32 = torch.empty((B,) + FRAME, dtype=torch.float32, device="cuda")
        hue64 = torch.empty((B,) + FRAME[:2], dtype=torch.float64, device="cuda")
        hue32 = torch.empty((B,) + FRAME[:2], dtype=torch.float32, device="cuda")
        _cuda._rgb_to_husl(dev_in, out=h32, stream=stream)
        # each kernel against the SLOWER of its two limits: HBM bytes at the measured copy peak, or
        # instruction issue (warp instructions of the ncu capture, profiles/traffic.json, scaled to this launch,
        # over SMs x 4 schedulers x max SM clock)
        issue_rate = _cuda.device_info()["sm_count"] * 4 * float(clocks.get("sm_max_mhz") or 1965) * 1e6
        for name, kern, fn, nbytes in (
                ("to_husl_u8_f64", "k_fwd_tma<double>", lambda: _cuda._rgb_to_husl(dev_in, out=dev_hsl, stream=stream), 27),
                ("to_rgb_f64_u8", "k_inv_tma<double>", lambda: _cuda._husl_to_rgb(dev_hsl, out=dev_out, stream=stream), 27),
                ("to_husl_u8_f32", "k_fwd_u8<float>", lambda: _cuda._rgb_to_husl(dev_in, out=h32, stream=stream), 15),
                ("to_rgb_f32_u8", "k_inv_cp<float>", lambda: _cuda._husl_to_rgb(h32, out=dev_out, stream=stream), 15),
                ("to_hue_u8_f64", "k_fwd_u8<double,hue>", lambda: _cuda._rgb_to_hue(dev_in, out=hue64, stream=stream), 11),
                ("to_hue_u8_f32", "k_fwd_u8<float,hue>", lambda: _cuda._rgb_to_hue(dev_in, out=hue32, stream=stream), 7),
                ("fused_edit_u8_u8", "k_edit", lambda: _cuda.edit_rgb(dev_in, out=dev_out, stream=stream, hue=(250, 290),
                                                                      invert=True, l=(0.5, 0)), 6)):
            if hdt != np.float64 and "f64" in name and "hue" not in name:
                continue
            ms = timed(fn)
            variants[name] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 2),
                              "gb_s": round(px * nbytes / ms / 1e6, 1), "hbm_frac": round(px * nbytes / ms / 1e6 / hbm, 3)}
            winst = warp_inst_for(kern)
            if winst:
                hbm_ms = px * nbytes / hbm / 1e6
                issue_ms = winst * (px / 66355200.0) / issue_rate * 1e3
                variants[name].update({"issue_frac": round(issue_ms / ms, 3), "bound": "issue" if issue_ms > hbm_ms else "hbm",
                                       "roofline_frac": round(max(hbm_ms, issue_ms) / ms, 3)})
        # decoder surfaces and float images as inputs (8 frames as one tall NV12 surface; 2 frames of float64 RGB)
        ysurf = torch.randint(0, 256, (B * FRAME[0], FRAME[1]), dtype=torch.uint8, device="cuda")
        uvsurf = torch.randint(0, 256, (B * FRAME[0] // 2, FRAME[1]), dtype=torch.uint8, device="cuda")
        hv64 = dev_hsl.view(B * FRAME[0], FRAME[1], 3) if hdt == np.float64 else None
        hv32 = h32.view(B * FRAME[0], FRAME[1], 3)
        f64in = (dev_in[:2].to(torch.float64) / 255.0).contiguous()
        f64out = torch.empty(f64in.shape, dtype=torch.float64, device="cuda")
        px2 = f64in.numel() // 3
        extra = [("nv12_to_husl_f32", "k_nv12<husl,float>", lambda: _cuda.nv12_to_husl(ysurf, uvsurf, out=hv32, stream=stream), 13.5, px),
                 ("frgb_f64_to_husl_f64", "k_fwd_flt<double,double>", lambda: _cuda._rgb_to_husl(f64in, out=f64out, stream=stream), 48, px2),
                 ("husl_f64_to_frgb_f64", "k_inv_flt<double,double>",
                  lambda: _cuda._husl_to_rgb(f64out, out=f64in, dtype=np.float64, stream=stream), 48, px2)]
        if hv64 is not None:
            extra.insert(0, ("nv12_to_husl_f64", "k_nv12<husl,double>",
                             lambda: _cuda.nv12_to_husl(ysurf, uvsurf, out=hv64, stream=stream), 25.5, px))
        for name, kern, fn, nbytes, npx in extra:
            ms = timed(fn)
            variants[name] = {"ms": round(ms, 4), "gpix_s": round(npx / ms / 1e6, 2),
                              "gb_s": round(npx * nbytes / ms / 1e6, 1), "hbm_frac": round(npx * nbytes / ms / 1e6 / hbm, 3)}
            winst = warp_inst_for(kern)
            if winst:
                hbm_ms = npx * nbytes / hbm / 1e6
                issue_ms = winst * (npx / 66355200.0) / issue_rate * 1e3
                variants[name].update({"issue_frac": round(issue_ms / ms, 3), "bound": "issue" if issue_ms > hbm_ms else "hbm",
                                       "roofline_frac": round(max(hbm_ms, issue_ms) / ms, 3)})
        del h32, hue64, hue32, ysurf, uvsurf, f64in, f64out
        # configs[3] as its callers write it (README.md:77-81): to_husl -> lightness edit -> to_rgb, with the
        # edit as the caller's own device op between the two conversions, and the same result from the fused kernel
        def with_edit():
            nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
            dev_hsl[..., 2] *= 0.5
            nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)
        ms = timed(with_edit, 10)
        variants["to_husl_edit_to_rgb_3_calls"] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 2)}
        ms = timed(lambda: _cuda.edit_rgb(dev_in, out=dev_out, stream=stream, l=(0.5, 0)), 10)
        variants["to_husl_edit_to_rgb_fused"] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 2)}

    # ---- cpu baseline (rank 0, N=1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = time_cpu(ReferenceCPU(), args.cpu_seconds)
        except Exception as e:          # the checker is optional for the number, never for the tests
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        hbm, which = peaks()
        f64 = hdt == np.float64
        dom = (("to_husl", "k_fwd_tma<double>" if f64 else "k_fwd_u8<float>", t_fwd) if t_fwd >= t_inv
               else ("to_rgb", "k_inv_tma<double>" if f64 else "k_inv_cp<float>", t_inv))
        achieved = px * bpp / (dom[2] * 1e-3) / 1e9
        kname = dom[1]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(args, world, B),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": px * 3, "d2h_bytes_per_step": px * 3,
                    "steps": Ke, "link_gbs": link, "blocking_calls_mpix_s_this_rank": e2e_blocking,
                    "host_husl_mpix_s_this_rank": e2e_host_husl, "path": "nphusl.to_husl(pinned host u8 -> device HUSL) + nphusl.to_rgb(device HUSL -> pinned host u8), "
                            "blocking=False on two alternating streams; wall clock incl. final synchronize"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic_for(kname), "peak_source": which,
                         "bytes_per_pixel": bpp, "pixels_per_launch": px,
                         "ms_per_launch": {"to_husl": t_fwd, "to_rgb": t_inv}},
            "kernels": variants,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
