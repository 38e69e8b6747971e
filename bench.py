#!/usr/bin/env python
"""Headline benchmark of the HUSL hot path (BASELINE.json: "Mpix/s to_husl &
to_rgb, 4K frame batches, 1/2/4/8 B200; % of roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--workload batch|video]

One STEP (BASELINE.json configs[3]) = one pass of the hot path over one batch of synthetic
3840x2160x3 uint8 frames: ``to_husl`` (uint8 RGB -> float64 HUSL, the reference's dtype), the
lightness edit of the reference's README.md:77-81 (every pixel whose hue is NOT in 250..290 gets
``L *= 0.5``) and ``to_rgb`` (float64 HUSL -> uint8 RGB).  The metric counts every pixel once per
step.  The edit is inside the timed step in every figure of the line; ``config4`` names the ways
of writing it (edit applied by ``to_rgb`` while loading -- the headline --, an edit pass over the
HUSL array, a caller-side torch op, and the one-kernel fused form).

* ``value``  : frames already resident in HBM, two kernels back to back on one stream,
  CUDA-event timing, max over ranks (weak scaling: every GPU owns its own batch, no data-path
  collective -- the path is embarrassingly parallel).
* ``e2e``    : the same step through the public API (``nphusl.to_husl`` / ``nphusl.to_rgb`` ->
  ctypes -> C ABI) with HOST buffers: uint8 frames start in pinned host memory and uint8 results
  end in pinned host memory, H2D and D2H inside the timed region; the HUSL intermediate stays on
  the GPU.  Beside it, in the same object: ``host_husl`` (the reference's own data flow: the float64
  HUSL array returns to the host and goes up again, 27 B/px each way), ``nv12`` (frames as NV12
  surfaces, 1.5 B/px each way) and ``jpeg`` (compressed frames in and out, nvJPEG on the GPU) --
  every one aggregated over ranks like ``e2e.value``.
* ``video``  : BASELINE.json configs[4] -- 1000 frames of 4K video cycled from a pool of 32
  distinct pinned frames, ``to_husl`` + ``to_rgb``, wall clock incl. H2D / D2H; one process per GPU
  (``video.process_per_gpu``) and ONE process driving every GPU (``video.in_process``, rank 0).
* ``roofline``: dominant kernel's algorithmic bytes / its mean CUDA-event duration against the
  measured HBM copy peak (MEASURED_PEAKS.json).
* ``cpu_baseline`` / ``--impl reference``: the reference package itself (nphusl 1.5.0 compiled from
  /root/reference into oracle/_ref/refpkg, see oracle/build_ref.sh) through ITS public API --
  ``to_husl`` (its C + OpenMP LUT path), the same numpy edit, ``to_rgb`` (its Cython + OpenMP path) --
  on all host cores.
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
# README.md:77-81: `bluish = (hue > 250) & (hue < 290); lightness[~bluish] *= 0.5`
EDIT = dict(hue=(250.0, 290.0), invert=True, l=(0.5, 0.0))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="batch", choices=["batch", "video"],
                    help="batch: configs[3] (the headline line, with configs[4] as its `video` object); "
                         "video: configs[4] alone as the line's value")
    ap.add_argument("--frames", type=int, default=8, help="4K frames per GPU per step")
    ap.add_argument("--husl-dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--video-frames", type=int, default=1000)
    ap.add_argument("--video-pool", type=int, default=32)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="budget of the whole --impl reference run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-video", action="store_true")
    ap.add_argument("--no-jpeg", action="store_true")
    return ap.parse_args()


def make_frames(n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n,) + FRAME, dtype=np.uint8)


def config(args, n_gpus, frames):
    return {"workload": "configs[3]: batch of %d synthetic 3840x2160x3 uint8 frames per GPU, to_husl -> lightness edit "
                        "(README.md:77-81: L *= 0.5 where hue is outside 250..290) -> to_rgb" % frames,
            "frames_per_gpu": frames, "pixels_per_step": frames * FRAME_PX * n_gpus,
            "husl_dtype": args.husl_dtype, "edit": EDIT,
            "l2": "inputs larger than L2: %.0f MB read+written per kernel launch vs 126 MB L2"
                  % (frames * FRAME_PX * (3 + 3 * np.dtype(args.husl_dtype).itemsize) / 1e6),
            "precision": "float32 arithmetic on hi+lo float pairs, HUSL stored as %s (the reference's dtype is float64); "
                         "all 2^24 colours within 6.6e-5 of the reference float64 path (tests/test_cuda_parity.py; north_star "
                         "bar: 1e-3); every run checks the exact uint8 round trip of the batch without the edit and the "
                         "bit-for-bit equality of the edited result with the fused kernel and the three-call form" % args.husl_dtype,
            "parallelism": "frames sharded over %d GPU(s), one process per GPU, no collective" % n_gpus}


# ---------------------------------------------------------------------------
# reference CPU implementation (oracle/_ref/refpkg: the reference package itself)
# ---------------------------------------------------------------------------

def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core it may run on."""
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(want)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(want)
    except OSError:
        pass
    return want


class ReferenceCPU:
    """The reference's own code for the step, through its own public API (module docstring)."""

    def __init__(self):
        self.cores = _all_host_threads()
        import oracle as orc
        self.orc = orc
        self.pkg = None
        if orc.have_ref_package():
            try:
                self.pkg = orc.ref_package()
            except Exception:
                self.pkg = None
        self.kind = "reference" if self.pkg is not None else "port"

    def to_husl(self, rgb):
        if self.pkg is not None:
            return self.pkg.to_husl(rgb)                     # nphusl.py:48-54 -> _simd_opt -> rgb_to_husl_nd
        return self.orc.simd_rgb_to_husl(rgb, "lut")

    @staticmethod
    def edit(hsl):
        hue, lightness = hsl[..., 0], hsl[..., 2]            # README.md:77-81, verbatim
        bluish = np.logical_and(hue > 250, hue < 290)
        lightness[~bluish] *= 0.5
        return hsl

    def to_rgb(self, hsl):
        if self.pkg is not None:
            return self.pkg.to_rgb(hsl)                      # nphusl.py:39-45 -> _cython_opt._husl_to_rgb + to_rgb_int
        return self.orc.husl_to_rgb(hsl)

    def step(self, rgb):
        return self.to_rgb(self.edit(self.to_husl(rgb)))

    def describe(self):
        if self.pkg is not None:
            return ("reference package (oracle/_ref/refpkg/nphusl_ref, compiled from /root/reference): nphusl.to_husl [C + OpenMP "
                    "rgb_to_husl_nd, default LUT build -O3 -ffast-math] -> numpy lightness edit -> nphusl.to_rgb [Cython + OpenMP]; "
                    "SIMD=%s CYTHON=%s" % (sorted(self.pkg.SIMD), sorted(self.pkg.CYTHON)))
        return "oracle/husl_oracle.c restatement (reference build absent)"


def time_cpu(ref, budget_s):
    """best-of-N of the step on one 4K frame (timeit.repeat(number=1) style,
    reference tests/performance_test.py:82-83) within ~budget_s seconds."""
    rgb = make_frames(1, 12345)[0]
    ref.step(rgb[:64])                                       # load libraries, spin up OpenMP
    best, parts, spent, reps = float("inf"), None, 0.0, 0
    while reps < 2 or (spent < budget_s and reps < 50):
        t0 = time.perf_counter()
        hsl = ref.to_husl(rgb)
        t1 = time.perf_counter()
        ref.edit(hsl)
        t2 = time.perf_counter()
        ref.to_rgb(hsl)
        t3 = time.perf_counter()
        if t3 - t0 < best:
            best, parts = t3 - t0, (t1 - t0, t2 - t1, t3 - t2)
        spent += t3 - t0
        reps += 1
    px = rgb.size // 3
    return {"value": px / best / 1e6, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
            "sample": "one 3840x2160 frame, best of %d repeats; %s" % (reps, ref.describe()),
            "to_husl_mpix_s": px / parts[0] / 1e6, "edit_mpix_s": px / parts[1] / 1e6, "to_rgb_mpix_s": px / parts[2] / 1e6}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = ReferenceCPU()
    K, W, B = args.steps, max(args.warmup, 0), args.frames
    batch = make_frames(B, 1000)                             # rank 0's batch of the CUDA arm
    t0 = time.perf_counter()
    ref.step(batch[0])                                       # first call: OpenMP spin-up, page faults
    ref.step(batch[0])
    t_frame = (time.perf_counter() - t0) / 2
    # bounded sample: whole batches when K + W steps of them fit the budget, else the first f frames of the batch
    f = int(max(1, min(B, args.ref_seconds / max(K + W, 1) / max(t_frame, 1e-3))))
    # the reference's compiled paths take images, not batches (_cython_opt.pyx:177-192: 2-D / 3-D arrays):
    # the f frames of a step are handed over as ONE tall (f * 2160, 3840, 3) image, one call per step
    rgb = batch[:f].reshape(f * FRAME[0], FRAME[1], 3)
    px = f * FRAME_PX
    for _ in range(W):
        ref.step(rgb)
    t0 = time.perf_counter()
    for _ in range(K):
        ref.step(rgb)
    dt = time.perf_counter() - t0
    val = px * K / dt / 1e6
    cfg = config(args, 1, args.frames)
    cfg["reference_sample"] = ("%d of the %d frames of the batch per step (the whole batch when K + W steps of it fit "
                               "--ref-seconds %.0f s)" % (f, B, args.ref_seconds))
    cfg["parallelism"] = "host cores of the box (rank 0), %d OpenMP threads" % ref.cores
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": dt / max(K, 1) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                             "sample": "%d x 3840x2160 frames per step, %d steps; %s" % (f, K, ref.describe())},
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


def _traffic_file():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def traffic_for(kernel):
    return _traffic_file().get(kernel)


def warp_inst_for(kernel):
    """warp instructions per 66 355 200-px launch from the committed ncu capture (None when not recorded)"""
    return _traffic_file().get("warp_inst", {}).get(kernel)


class Ranks:
    """torch.distributed plumbing of the bench: barrier, max over ranks, and a way to park the other
    ranks WITHOUT a spinning collective while rank 0 drives every GPU itself."""

    def __init__(self, torch, dist, world, rank, local):
        self.torch, self.dist, self.world, self.rank, self.local = torch, dist, world, rank, local

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def park_until(self, key, leader_work):
        """rank 0 runs `leader_work()`; the others sleep on the rendezvous store (no NCCL kernel spinning on
        their GPUs, no busy CPU) until it is done.  -> leader_work's result on rank 0, None elsewhere."""
        if self.world == 1:
            return leader_work()
        self.barrier()
        store = None
        try:
            store = self.dist.distributed_c10d._get_default_store()
        except Exception:
            pass
        res = None
        if self.rank == 0:
            try:
                res = leader_work()
            finally:
                if store is not None:
                    store.set(key, "done")
        elif store is not None:
            while True:
                try:
                    store.wait([key], __import__("datetime").timedelta(seconds=600))
                    break
                except Exception:
                    time.sleep(0.5)
        self.barrier()
        return res


def run_video(args, rk, nphusl, _cuda, hdt):
    """BASELINE.json configs[4]: 1000 frames of 4K video from a pool of 32 distinct pinned frames cycled
    to 1000, to_husl + to_rgb (no edit), every result frame copied back to the host; wall clock around
    the whole stream incl. H2D / D2H, max over ranks.  Method of the reference's
    tests/performance_test.py:66-103 (best of N passes of one call), N = 2 passes here."""
    from nphusl import _shard
    n, P, B = args.video_frames, args.video_pool, args.frames
    pool = _cuda.pinned_empty((P,) + FRAME, np.uint8)
    pool[...] = make_frames(P, 5)                            # one video: every rank holds the same pool
    out = {"frames": n, "pool": P, "batch": B, "husl_dtype": str(hdt),
           "bytes_per_frame_each_way": FRAME_PX * 3,
           "path": "nphusl._shard.stream_frames: per GPU one host thread, two CUDA streams, blocking=False calls on pinned "
                   "frames, ring of two pinned result batches"}

    def check(res):
        for dev, (i, p, m, view) in res["last"].items():
            if not np.array_equal(view, pool[p:p + m]):
                raise SystemExit("bench.py: video round trip of frames %d..%d on GPU %d is not exact" % (i, i + m, dev))

    # ---- one process per GPU: this rank converts its contiguous block of the stream on its own GPU
    a, b = _shard.shard_bounds(n, rk.world)[rk.rank]
    passes = []
    for _ in range(2):
        res = _shard.stream_frames(pool, b - a, devices=[rk.local], first_frame=a, batch=B, dtype=hdt, ready=rk.barrier)
        check(res)
        passes.append(rk.max(res["seconds"]))
    dt = min(passes)
    out["process_per_gpu"] = {"gpus": rk.world, "seconds": round(dt, 4), "fps_4k": round(n / dt, 1),
                              "mpix_s": round(n * FRAME_PX / dt / 1e6, 1),
                              "all_passes_fps_4k": [round(n / t, 1) for t in passes]}

    # ---- ONE process (rank 0) drives every GPU of the job through the same function
    def in_process():
        devs = list(range(rk.world)) if rk.world > 1 else [rk.local]
        rec = {"gpus": len(devs)}
        for schedule in ("static", "dynamic"):
            ps = []
            for _ in range(2):
                res = _shard.stream_frames(pool, n, devices=devs, batch=B, dtype=hdt, schedule=schedule)
                check(res)
                ps.append(res["seconds"])
            t = min(ps)
            one = {"seconds": round(t, 4), "fps_4k": round(n / t, 1), "mpix_s": round(n * FRAME_PX / t / 1e6, 1),
                   "all_passes_fps_4k": [round(n / x, 1) for x in ps],
                   "per_device": {str(d): [v["frames"], round(v["seconds"], 4)] for d, v in sorted(res["per_device"].items())}}
            if schedule == "static":
                rec.update(one)                      # same partition as process_per_gpu: contiguous blocks of the stream
            else:
                rec["dynamic"] = one                 # batches first come, first served (only one process can do this)
        return rec
    try:
        out["in_process"] = rk.park_until("nphusl_bench_inproc", in_process)
    except Exception as e:                                   # e.g. another rank's context leaves too little memory
        out["in_process"] = {"error": repr(e)}
    del pool
    return out


def jpeg_leg(args, rk, torch, _cuda, B):
    """Compressed frames in, compressed frames out: torchvision's nvJPEG decode (device="cuda") ->
    nphusl fused edit on the decoded (3, H, W) planes, in place -> nvJPEG encode -> bytes to the host.
    Frames are synthetic low-frequency images (uniform noise does not compress); compressed bytes per
    pixel are reported beside the rate."""
    try:
        from torchvision.io import decode_jpeg, encode_jpeg
    except Exception as e:
        return {"error": "torchvision unavailable: %r" % (e,)}
    try:
        rng = np.random.default_rng(77 + rk.rank)
        frames = []
        for _ in range(B):                                   # smooth colour fields + mild texture
            low = torch.from_numpy(rng.random((1, 3, 34, 60), dtype=np.float32)).cuda()
            img = torch.nn.functional.interpolate(low, size=FRAME[:2], mode="bicubic", align_corners=False)[0]
            img = img + 0.03 * torch.randn(img.shape, device="cuda")
            frames.append((img.clamp(0, 1) * 255).to(torch.uint8))
        blobs = [b.cpu() for b in encode_jpeg(frames, quality=90)]
        in_bytes = sum(int(b.numel()) for b in blobs)
        del frames

        def step():
            dec = decode_jpeg(blobs, device="cuda")
            for fr in dec:
                _cuda.edit_rgb(fr, out=fr, channels_first=True, stream=torch.cuda.current_stream(), **EDIT)
            enc = encode_jpeg(dec, quality=90)
            return [e.cpu() for e in enc]
        outs = step()
        torch.cuda.synchronize()
        rk.barrier()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            outs = step()
        torch.cuda.synchronize()
        dt = rk.max(time.perf_counter() - t0)
        out_bytes = sum(int(e.numel()) for e in outs)
        px = B * FRAME_PX
        return {"value": px * rk.world * reps / dt / 1e6, "unit": UNIT, "steps": reps,
                "h2d_bytes_per_step": in_bytes * rk.world, "d2h_bytes_per_step": out_bytes * rk.world,
                "compressed_bytes_per_pixel": {"in": round(in_bytes / px, 4), "out": round(out_bytes / px, 4)},
                "path": "torchvision.io.decode_jpeg(device='cuda') -> nphusl._cuda.edit_rgb(channels_first=True, in place) -> "
                        "torchvision.io.encode_jpeg (nvJPEG) -> .cpu(); quality 90; decode + encode are library code, "
                        "the colour work is this repository's kernel"}
    except Exception as e:
        return {"error": repr(e)}


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
    rk = Ranks(torch, dist, world, rank, local)
    barrier = rk.barrier

    B, K, W = args.frames, args.steps, max(args.warmup, 3)
    hdt = np.dtype(args.husl_dtype)
    tdt = torch.float64 if hdt == np.float64 else torch.float32
    px = B * FRAME_PX
    bpp = 3 + 3 * hdt.itemsize                       # algorithmic bytes per pixel per kernel
    edit = _cuda.Edit(**EDIT)

    if args.workload == "video":
        sampler = ClockSampler(physical_gpu_index(local))
        sampler.start()
        n0 = _cuda.launch_count()
        video = run_video(args, rk, nphusl, _cuda, hdt)
        launches = _cuda.launch_count() - n0
        clocks = sampler.stop()
        if rank == 0:
            v = video["process_per_gpu"]
            n = args.video_frames
            line = {"metric": METRIC, "value": v["mpix_s"], "unit": UNIT, "n_gpus": world, "steps": 2, "warmup": 1,
                    "ms_per_step": v["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic",
                    "config": {"workload": "configs[4]: %d-frame synthetic 4K video (pool of %d pinned frames) sharded over %d GPU(s), "
                                           "to_husl + to_rgb, end to end incl. H2D / D2H; one step = the whole stream"
                                           % (n, args.video_pool, world), "husl_dtype": args.husl_dtype},
                    "e2e": {"value": v["mpix_s"], "unit": UNIT, "h2d_bytes_per_step": n * FRAME_PX * 3, "d2h_bytes_per_step": n * FRAME_PX * 3},
                    "video": video, "gpu_launches": int(launches), "clocks": clocks}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # inputs: pinned host frames (e2e) and a device-resident copy (value)
    host_in = _cuda.pinned_empty((B,) + FRAME, np.uint8)
    host_in[...] = make_frames(B, 1000 + rank)
    host_out = _cuda.pinned_empty((B,) + FRAME, np.uint8)
    dev_in = torch.from_numpy(host_in).cuda()
    dev_hsl = torch.empty((B,) + FRAME, dtype=tdt, device="cuda")
    dev_out = torch.empty_like(dev_in)
    stream = torch.cuda.current_stream()

    def step_dev():
        nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
        nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream, edit=edit)

    # ---- what the step computes, checked before anything is timed ---------------------
    nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
    nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)
    if not torch.equal(dev_out, dev_in):
        raise SystemExit("bench.py: round trip is not exact -- refusing to report a number")
    expect = torch.empty_like(dev_in)                # the edited batch, from the one-kernel fused form
    _cuda.edit_rgb(dev_in, edit, out=expect, stream=stream)
    step_dev()
    if not torch.equal(dev_out, expect):
        raise SystemExit("bench.py: to_rgb(edit=...) and the fused edit kernel disagree")
    # the caller-side form: the README lines as torch ops on the HUSL array (the library's ranges are closed
    # intervals, so the comparison is written with >= / <= here; a hue of exactly 250 or 290 is the only difference)
    nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
    hue, light = dev_hsl[..., 0], dev_hsl[..., 2]
    light[~((hue >= 250) & (hue <= 290))] *= 0.5
    nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)
    if not torch.equal(dev_out, expect):
        raise SystemExit("bench.py: the three-call pipeline with a torch edit and the fused kernel disagree")
    del hue, light
    changed = int((expect != dev_in).any(dim=-1).sum().item())
    if not 0.5 * px < changed <= px:
        raise SystemExit("bench.py: the edit changed %d of %d pixels -- not the README edit" % (changed, px))

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
        nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream, edit=edit)
        if e:
            e[2].record(stream)
    t_end.record(stream)
    torch.cuda.synchronize()
    m1 = sampler.mark()
    launches = _cuda.launch_count() - n0
    total_ms = t_beg.elapsed_time(t_end)
    t_fwd = float(np.mean([e[0].elapsed_time(e[1]) for e in ev.values()]))
    t_inv = float(np.mean([e[1].elapsed_time(e[2]) for e in ev.values()]))
    if not torch.equal(dev_out, expect):
        raise SystemExit("bench.py: the timed steps did not produce the edited batch")
    total_ms = rk.max(total_ms)
    value = px * world * K / (total_ms * 1e-3) / 1e6

    # ---- host link of THIS box (diagnostic beside e2e): plain pinned copies -----------
    def link_gbs():
        nbytes = host_in.nbytes
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        flat_in = torch.from_numpy(host_in.reshape(-1))
        flat_out = torch.from_numpy(host_out.reshape(-1))

        def run(h2d, d2h, reps=3):
            best = 1e9
            for _ in range(reps):
                barrier()
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
                "duplex_each": round(run(True, True), 1),
                "note": "GB/s per GPU with all %d rank(s) copying at once" % world}
    link = link_gbs()

    # ---- e2e: public API, host buffers ------------------------------------------
    # step k runs on user stream k%2 with its own HUSL buffer; calls are
    # non-blocking (pinned memory), so the H2D of step k+1 overlaps the D2H of step k
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    hsl2 = [dev_hsl, torch.empty_like(dev_hsl)]
    Ke = max(4, min(K, 20))

    def timed_e2e(step, reps, warm=2):
        """wall clock around submit + drain of `reps` steps, max over ranks -> seconds"""
        for k in range(warm):
            step(k)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for k in range(reps):
            step(k)
        torch.cuda.synchronize()
        return rk.max(time.perf_counter() - t0)

    def step_e2e(k):
        s = streams[k & 1]
        nphusl.to_husl(host_in, out=hsl2[k & 1], stream=s, blocking=False)                 # H2D + kernel, chunk-pipelined
        nphusl.to_rgb(hsl2[k & 1], out=host_out, stream=s, blocking=False, edit=edit)      # edit + kernel + D2H

    # the plain way first: blocking calls, one after the other (reported beside the headline)
    def step_blocking(k):
        nphusl.to_husl(host_in, out=dev_hsl)
        nphusl.to_rgb(dev_hsl, out=host_out, edit=edit)
    e2e_blocking = px * world * 5 / timed_e2e(step_blocking, 5) / 1e6

    host_out[...] = 0
    e2e_s = timed_e2e(step_e2e, Ke, warm=4)
    e2e_val = px * world * Ke / e2e_s / 1e6
    if not np.array_equal(host_out, expect.cpu().numpy()):
        raise SystemExit("bench.py: e2e result is not the edited batch")

    # the reference's own data flow: the HUSL array comes back to the HOST (to_husl returns a
    # float64 ndarray, 24 B/px) and goes up again for to_rgb -- 27 B/px over the link each way
    e2e_extra = {}
    if not args.no_variants:
        try:
            host_hsl = _cuda.pinned_empty((B,) + FRAME, hdt)

            def step_host_husl(k):
                nphusl.to_husl(host_in, out=host_hsl)
                nphusl.to_rgb(host_hsl, out=host_out, edit=edit)
            t = timed_e2e(step_host_husl, 3, warm=1)
            e2e_extra["host_husl"] = {"value": px * world * 3 / t / 1e6, "unit": UNIT, "steps": 3,
                                      "h2d_bytes_per_step": px * (3 + 3 * hdt.itemsize) * world, "d2h_bytes_per_step": px * (3 + 3 * hdt.itemsize) * world,
                                      "path": "nphusl.to_husl(host u8 -> HOST %s HUSL), nphusl.to_rgb(HOST HUSL -> host u8, edit on load): the "
                                              "reference's data flow, pinned buffers, blocking calls" % hdt}
            if not np.array_equal(host_out, expect.cpu().numpy()):
                raise SystemExit("bench.py: host-HUSL e2e result is not the edited batch")
            # the same data flow with the two directions of the link used at once: frame i on stream i % 2,
            # to_husl (3 B/px up, 24 down) and to_rgb (24 up, 3 down) queued back to back, non-blocking --
            # one stream's 24 B/px download runs beside the other's 24 B/px upload
            host_out[...] = 0

            def step_host_husl_overlapped(k):
                for i in range(B):
                    s = streams[i & 1]
                    nphusl.to_husl(host_in[i], out=host_hsl[i], stream=s, blocking=False)
                    nphusl.to_rgb(host_hsl[i], out=host_out[i], stream=s, blocking=False, edit=edit)
            t = timed_e2e(step_host_husl_overlapped, 3, warm=1)
            e2e_extra["host_husl"]["overlapped"] = {
                "value": px * world * 3 / t / 1e6, "unit": UNIT, "steps": 3,
                "path": "the same calls per FRAME, blocking=False on two alternating streams: the HUSL download of one frame "
                        "beside the HUSL upload of the previous one (full-duplex link)"}
            if not np.array_equal(host_out, expect.cpu().numpy()):
                raise SystemExit("bench.py: overlapped host-HUSL e2e result is not the edited batch")
            del host_hsl
        except SystemExit:
            raise
        except Exception as e:                         # not enough pinned memory on this box
            e2e_extra["host_husl"] = {"error": repr(e)}
        # frames as NV12 surfaces (what decoders, capture cards and encoders exchange): 1.5 B/px each way
        try:
            from nphusl._cuda import rgb_to_nv12
            H8 = B * FRAME[0]
            y_in, uv_in = _cuda.pinned_empty((H8, FRAME[1]), np.uint8), _cuda.pinned_empty((H8 // 2, FRAME[1]), np.uint8)
            y_out, uv_out = _cuda.pinned_empty((H8, FRAME[1]), np.uint8), _cuda.pinned_empty((H8 // 2, FRAME[1]), np.uint8)
            dy, duv = rgb_to_nv12(dev_in.view(H8, FRAME[1], 3), stream=stream)        # the batch as one tall surface
            dy.copy_to_host(y_in, stream=stream)
            duv.copy_to_host(uv_in, stream=stream)
            ey, euv = _cuda.nv12_edit_nv12(dy, duv, edit, stream=stream)
            want_y, want_uv = ey.copy_to_host(stream=stream), euv.copy_to_host(stream=stream)
            torch.cuda.synchronize()
            del dy, duv, ey, euv

            def step_nv12(k):
                _cuda.nv12_edit_nv12(y_in, uv_in, edit, out_y=y_out, out_uv=uv_out, stream=streams[k & 1], blocking=False)
            t = timed_e2e(step_nv12, Ke, warm=3)
            if not (np.array_equal(y_out, want_y) and np.array_equal(uv_out, want_uv)):
                raise SystemExit("bench.py: NV12 e2e result differs from the device-resident result")
            e2e_extra["nv12"] = {"value": px * world * Ke / t / 1e6, "unit": UNIT, "steps": Ke,
                                 "h2d_bytes_per_step": px * 3 // 2 * world, "d2h_bytes_per_step": px * 3 // 2 * world,
                                 "path": "nphusl._cuda.nv12_edit_nv12(pinned host Y + UV planes in, pinned host planes out, blocking=False on two "
                                         "alternating streams): H2D, ONE fused kernel NV12 -> HUSL -> edit -> NV12, D2H; BT.709 video range"}
            del y_in, uv_in, y_out, uv_out
        except SystemExit:
            raise
        except Exception as e:
            e2e_extra["nv12"] = {"error": repr(e)}
        if not args.no_jpeg:
            e2e_extra["jpeg"] = jpeg_leg(args, rk, torch, _cuda, B)

    # ---- config 5: the 1000-frame video --------------------------------------------------
    video = None
    if not args.no_video:
        video = run_video(args, rk, nphusl, _cuda, hdt)
    clocks = sampler.stop()                           # sampled across the value and e2e timed regions
    clocks["value_region"] = sampler.stats(m0, m1)    # the device-resident K steps alone

    # ---- per-kernel variants (rank 0, informational) -------------------------------
    variants, config4 = {}, {}
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
        # configs[3] in the ways a caller can write it; all produce the same uint8 batch
        def no_edit():
            nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
            nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)

        def torch_edit():
            nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
            hue, light = dev_hsl[..., 0], dev_hsl[..., 2]
            light[~((hue > 250) & (hue < 290))] *= 0.5
            nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)

        def edit_pass():
            nphusl.to_husl(dev_in, out=dev_hsl, stream=stream)
            _cuda.edit_husl(dev_hsl, edit, stream=stream)
            nphusl.to_rgb(dev_hsl, out=dev_out, stream=stream)
        for name, fn, reps, what in (
                ("edit_in_to_rgb", step_dev, 10, "to_husl; to_rgb(hsl, edit=...) -- 2 kernels, the headline `value`"),
                ("edit_husl_pass", edit_pass, 10, "to_husl; nphusl._cuda.edit_husl(hsl) in place; to_rgb -- 3 kernels"),
                ("torch_edit", torch_edit, 5, "to_husl; the README lines as torch ops on the device array; to_rgb"),
                ("fused_kernel", lambda: _cuda.edit_rgb(dev_in, edit, out=dev_out, stream=stream), 10, "nphusl.edit(img, ...) -- 1 kernel, 6 B/px"),
                ("no_edit", no_edit, 10, "to_husl; to_rgb -- the conversions alone")):
            ms = timed(fn, reps)
            config4[name] = {"ms": round(ms, 4), "gpix_s": round(px / ms / 1e6, 2), "how": what}
        h32 = torch.empty((B,) + FRAME, dtype=torch.float32, device="cuda")
        hue64 = torch.empty((B,) + FRAME[:2], dtype=torch.float64, device="cuda")
        hue32 = torch.empty((B,) + FRAME[:2], dtype=torch.float32, device="cuda")
        _cuda._rgb_to_husl(dev_in, out=h32, stream=stream)
        # each kernel against the SLOWER of its two limits: HBM bytes at the measured copy peak, or
        # instruction issue (warp instructions of the ncu capture, profiles/traffic.json, scaled to this launch,
        # over SMs x 4 schedulers x max SM clock)
        issue_rate = _cuda.device_info()["sm_count"] * 4 * float(clocks.get("sm_max_mhz") or 1965) * 1e6

        def record(name, kern, fn, nbytes, npx, reps=20):
            ms = timed(fn, reps)
            variants[name] = {"ms": round(ms, 4), "gpix_s": round(npx / ms / 1e6, 2),
                              "gb_s": round(npx * nbytes / ms / 1e6, 1), "hbm_frac": round(npx * nbytes / ms / 1e6 / hbm, 3)}
            winst = warp_inst_for(kern)
            if winst:
                hbm_ms = npx * nbytes / hbm / 1e6
                issue_ms = winst * (npx / 66355200.0) / issue_rate * 1e3
                variants[name].update({"issue_frac": round(issue_ms / ms, 3), "bound": "issue" if issue_ms > hbm_ms else "hbm",
                                       "roofline_frac": round(max(hbm_ms, issue_ms) / ms, 3)})
        lut_ready = True
        try:
            _cuda.lut_generate()
        except Exception:
            lut_ready = False
        for name, kern, fn, nbytes in (
                ("to_husl_u8_f64", "k_fwd_tma<double>", lambda: _cuda._rgb_to_husl(dev_in, out=dev_hsl, stream=stream), 27),
                ("to_rgb_f64_u8", "k_inv_direct", lambda: _cuda._husl_to_rgb(dev_hsl, out=dev_out, stream=stream), 27),
                ("to_rgb_f64_u8_edit", "k_inv_direct<edit>", lambda: _cuda._husl_to_rgb(dev_hsl, out=dev_out, stream=stream, edit=edit), 27),
                ("to_husl_u8_f64_lut", "k_fwd_lut", lambda: _cuda._rgb_to_husl(dev_in, out=dev_hsl, stream=stream, mode="lut"), 27),
                ("to_husl_u8_f32", "k_fwd_u8<float>", lambda: _cuda._rgb_to_husl(dev_in, out=h32, stream=stream), 15),
                ("to_rgb_f32_u8", "k_inv_direct32", lambda: _cuda._husl_to_rgb(h32, out=dev_out, stream=stream), 15),
                ("to_rgb_f32_u8_edit", "k_inv_cp<float,edit>", lambda: _cuda._husl_to_rgb(h32, out=dev_out, stream=stream, edit=edit), 15),
                ("to_hue_u8_f64", "k_fwd_u8<double,hue>", lambda: _cuda._rgb_to_hue(dev_in, out=hue64, stream=stream), 11),
                ("to_hue_u8_f32", "k_fwd_u8<float,hue>", lambda: _cuda._rgb_to_hue(dev_in, out=hue32, stream=stream), 7),
                ("fused_edit_u8_u8", "k_edit", lambda: _cuda.edit_rgb(dev_in, edit, out=dev_out, stream=stream), 6)):
            if hdt != np.float64 and "f64" in name and "hue" not in name:
                continue
            if name.endswith("_lut") and not lut_ready:
                continue
            record(name, kern, fn, nbytes, px)
        # decoder surfaces and float images as inputs (8 frames as one tall NV12 surface; 2 frames of float RGB)
        ysurf = torch.randint(0, 256, (B * FRAME[0], FRAME[1]), dtype=torch.uint8, device="cuda")
        uvsurf = torch.randint(0, 256, (B * FRAME[0] // 2, FRAME[1]), dtype=torch.uint8, device="cuda")
        hv64 = dev_hsl.view(B * FRAME[0], FRAME[1], 3) if hdt == np.float64 else None
        hv32 = h32.view(B * FRAME[0], FRAME[1], 3)
        f64in = (dev_in[:2].to(torch.float64) / 255.0).contiguous()
        f64out = torch.empty(f64in.shape, dtype=torch.float64, device="cuda")
        f32in, f32out = f64in.to(torch.float32), torch.empty(f64in.shape, dtype=torch.float32, device="cuda")
        g64 = torch.rand((2 * FRAME[0], FRAME[1]), dtype=torch.float64, device="cuda")
        gh64 = torch.empty_like(g64)
        px2 = f64in.numel() // 3
        if hv64 is not None:
            record("nv12_to_husl_f64", "k_nv12<husl,double>", lambda: _cuda.nv12_to_husl(ysurf, uvsurf, out=hv64, stream=stream), 25.5, px)
        record("nv12_to_husl_f32", "k_nv12<husl,float>", lambda: _cuda.nv12_to_husl(ysurf, uvsurf, out=hv32, stream=stream), 13.5, px)
        record("edit_husl_f64_in_place", "k_husl_edit_wide<double>", lambda: _cuda.edit_husl(dev_hsl, edit, stream=stream), 48, px)
        record("nv12_edit_nv12", "k_to_nv12<nv12>", lambda: _cuda.nv12_edit_nv12(ysurf, uvsurf, edit, out_y=ysurf, out_uv=uvsurf, stream=stream), 3, px)
        record("frgb_f64_to_husl_f64", "k_fwd_flt<double,double>", lambda: _cuda._rgb_to_husl(f64in, out=f64out, stream=stream), 48, px2)
        record("husl_f64_to_frgb_f64", "k_inv_flt_wide",
               lambda: _cuda._husl_to_rgb(f64out, out=f64in, dtype=np.float64, stream=stream), 48, px2)
        record("frgb_f32_to_husl_f32", "k_fwd_flt<float,float>", lambda: _cuda._rgb_to_husl(f32in, out=f32out, dtype=np.float32, stream=stream), 24, px2)
        with __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            record("gray_f64_to_hue_f64", "k_gray_hue", lambda: nphusl.to_hue(g64, out=gh64, stream=stream), 16, px2)
        del h32, hue64, hue32, ysurf, uvsurf, f64in, f64out, f32in, f32out, g64, gh64

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
               else ("to_rgb", "k_inv_direct<edit>" if f64 else "k_inv_cp<float,edit>", t_inv))
        achieved = px * bpp / (dom[2] * 1e-3) / 1e9
        kname = dom[1]
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": px * 3 * world, "d2h_bytes_per_step": px * 3 * world,
               "bytes_scope": "whole job (all %d ranks), like `value`" % world,
               "steps": Ke, "link_gbs": link, "blocking_calls_mpix_s": e2e_blocking,
               "path": "nphusl.to_husl(pinned host u8 -> device HUSL) + nphusl.to_rgb(device HUSL -> pinned host u8, edit=...), "
                       "blocking=False on two alternating streams; wall clock incl. final synchronize, max over ranks"}
        e2e.update(e2e_extra)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(args, world, B),
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic_for(kname), "peak_source": which,
                         "bytes_per_pixel": bpp, "pixels_per_launch": px,
                         "ms_per_launch": {"to_husl": t_fwd, "to_rgb_with_edit": t_inv}},
            "config4": config4,
            "video": video,
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
