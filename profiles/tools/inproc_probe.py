#!/usr/bin/env python
"""Why is one process driving N GPUs slower than N processes driving one GPU each?
(round-1 record: config 5 ran at 703 frames/s on 8 GPUs in-process vs 1 381 on one GPU.)

Runs the config-5 inner loop (pinned uint8 4K frames -> to_husl -> to_rgb -> pinned uint8)
over N GPUs in several ways and prints one JSON object:

  threads      one Python thread per GPU, blocking=False calls (what _shard / configs_report did)
  one_thread   ONE Python thread enqueues round-robin onto every GPU (calls are pure enqueues)
  copies       threads, plain cudaMemcpyAsync H2D + D2H of the same bytes, no kernels
  procs        one process per GPU (multiprocessing spawn), same loop as `threads`
  threads_big  as `threads`, chunking off (one H2D / kernel / D2H per call)

    python profiles/tools/inproc_probe.py --gpus 2 --batches 24
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))

FRAME = (2160, 3840, 3)
B = 8
POOL = 16


def setup(dev):
    from nphusl import _cuda
    src = _cuda.pinned_empty((POOL,) + FRAME, np.uint8)
    src[...] = np.random.default_rng(100 + dev).integers(0, 256, src.shape, dtype=np.uint8)
    dst = _cuda.pinned_empty((2, B) + FRAME, np.uint8)
    streams = [_cuda.Stream(dev), _cuda.Stream(dev)]
    hs = [_cuda.DeviceArray((B,) + FRAME, np.float64, dev) for _ in range(2)]
    for k in range(2):
        _cuda._rgb_to_husl(src[:B], out=hs[k], stream=streams[k], device=dev, blocking=False)
        _cuda._husl_to_rgb(hs[k], out=dst[k], stream=streams[k], device=dev, blocking=False)
        streams[k].synchronize()
    return src, dst, streams, hs


def enqueue(state, dev, k):
    from nphusl import _cuda
    src, dst, streams, hs = state
    p0 = (k * B) % (POOL - B + 1)
    s = streams[k & 1]
    _cuda._rgb_to_husl(src[p0:p0 + B], out=hs[k & 1], stream=s, device=dev, blocking=False)
    _cuda._husl_to_rgb(hs[k & 1], out=dst[k & 1], stream=s, device=dev, blocking=False)


def loop(state, dev, n):
    t0 = time.perf_counter()
    for k in range(n):
        enqueue(state, dev, k)
    t1 = time.perf_counter()
    for s in state[2]:
        s.synchronize()
    return t1 - t0, time.perf_counter() - t0


def run_threads(devs, n, states):
    bar = threading.Barrier(len(devs) + 1)
    info = {}

    def work(d):
        bar.wait()
        info[d] = loop(states[d], d, n)
    ths = [threading.Thread(target=work, args=(d,)) for d in devs]
    for t in ths:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return dt, {str(d): [round(x, 4) for x in info[d]] for d in devs}


def run_one_thread(devs, n, states):
    t0 = time.perf_counter()
    for k in range(n):
        for d in devs:
            enqueue(states[d], d, k)
    t1 = time.perf_counter()
    for d in devs:
        for s in states[d][2]:
            s.synchronize()
    return time.perf_counter() - t0, {"enqueue_s": round(t1 - t0, 4)}


def run_copies(devs, n, states):
    from nphusl import _cuda
    lib = _cuda.load()
    bar = threading.Barrier(len(devs) + 1)
    scratch = {d: _cuda.DeviceArray((2, B) + FRAME, np.uint8, d) for d in devs}

    def work(d):
        src, dst, streams, _ = states[d]
        nb = B * FRAME[0] * FRAME[1] * 3
        bar.wait()
        for k in range(n):
            p0 = (k * B) % (POOL - B + 1)
            s = streams[k & 1]
            lib.nphusl_cuda_memcpy(scratch[d][k & 1].ptr, src[p0:p0 + B].ctypes.data, nb, 1, d, _cuda._stream_ptr(s))
            lib.nphusl_cuda_memcpy(dst[k & 1].ctypes.data, scratch[d][k & 1].ptr, nb, 2, d, _cuda._stream_ptr(s))
        for s in streams:
            s.synchronize()
    ths = [threading.Thread(target=work, args=(d,)) for d in devs]
    for t in ths:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, {}


def proc_main(dev, n, bar, q):
    from nphusl import _cuda  # noqa: F401
    state = setup(dev)
    bar.wait()
    t = loop(state, dev, n)
    bar.wait()
    q.put((dev, t))


def run_procs(devs, n):
    ctx = mp.get_context("spawn")
    bar = ctx.Barrier(len(devs) + 1)
    q = ctx.Queue()
    ps = [ctx.Process(target=proc_main, args=(d, n, bar, q)) for d in devs]
    for p in ps:
        p.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    info = dict(q.get() for _ in devs)
    for p in ps:
        p.join()
    return dt, {str(d): [round(x, 4) for x in info[d]] for d in devs}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--batches", type=int, default=24, help="batches of 8 frames per GPU")
    ap.add_argument("--modes", default="procs,threads,one_thread,copies,threads_big,threads")
    args = ap.parse_args()
    devs = list(range(args.gpus))
    n = args.batches
    px = n * B * FRAME[0] * FRAME[1]
    out = {"gpus": args.gpus, "batches_per_gpu": n, "cpus": len(os.sched_getaffinity(0))}
    modes = args.modes.split(",")

    def rec(name, dt, extra, ndev):
        out.setdefault(name, []).append({"gpus": ndev, "seconds": round(dt, 4), "fps_4k": round(ndev * n * B / dt, 1),
                                         "gbs_each_way": round(ndev * px * 3 / dt / 1e9, 2), **extra})

    if "procs" in modes:
        for nd in sorted({1, args.gpus}):
            dt, info = run_procs(devs[:nd], n)
            rec("procs", dt, {"per_dev_enqueue_total": info}, nd)
    from nphusl import _cuda
    states = {d: setup(d) for d in devs}
    for m in modes:
        if m == "procs":
            continue
        for nd in sorted({1, args.gpus}):
            if m == "threads":
                dt, info = run_threads(devs[:nd], n, states)
                rec(m, dt, {"per_dev_enqueue_total": info}, nd)
            elif m == "one_thread":
                dt, info = run_one_thread(devs[:nd], n, states)
                rec(m, dt, info, nd)
            elif m == "copies":
                dt, info = run_copies(devs[:nd], n, states)
                rec(m, dt, info, nd)
            elif m == "threads_big":
                _cuda.set_chunk_pixels(1 << 30)
                for d in devs[:nd]:
                    loop(states[d], d, 2)
                dt, info = run_threads(devs[:nd], n, states)
                _cuda.set_chunk_pixels(0)
                rec(m, dt, {"per_dev_enqueue_total": info}, nd)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
