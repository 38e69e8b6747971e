"""Multi-GPU host logic on CPU: the path shards by contiguous blocks with no
data-path collective.  Two ``gloo`` ranks each take ``shard_bounds(n, 2)[rank]``
of a seeded frame batch, convert their block (the oracle stands in for the
kernels: there is no GPU here), and the test checks that the blocks tile the
batch exactly, that the per-rank results equal the single-process result, and
that the benchmark's max-over-ranks timing reduction works."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from nphusl import _shard


def test_shard_bounds_tile_exactly():
    for n in (0, 1, 7, 8, 9, 1000):
        for parts in (1, 2, 3, 4, 8):
            b = _shard.shard_bounds(n, parts)
            assert len(b) == parts and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        _shard.shard_bounds(4, 0)


def test_in_process_fan_out_uses_disjoint_slices(monkeypatch):
    calls = []

    def fake(src, out=None, device=None, **kw):
        calls.append((device, src.shape[0]))
        out[...] = src.astype(out.dtype) + device
    monkeypatch.setattr(_shard._cuda, "_rgb_to_husl", fake)
    rgb = np.zeros((10, 4, 4, 3), np.uint8)
    out = _shard.to_husl(rgb, devices=[0, 1, 2])
    assert sorted(calls) == [(0, 4), (1, 3), (2, 3)]
    assert np.all(out[:4] == 0) and np.all(out[4:7] == 1) and np.all(out[7:] == 2)
    with pytest.raises(Exception):
        _shard.to_husl(rgb, devices=[])


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "husl-numpy_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle as orc
    from nphusl import _shard as sh
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = np.random.default_rng(99).integers(0, 256, (5, 16, 24, 3), dtype=np.uint8)
        a, b = sh.shard_bounds(frames.shape[0], world)[rank]
        mine = orc.rgb_to_husl(frames[a:b])
        back = orc.husl_to_rgb(mine)
        # "gather" = disjoint slices; here: exchange checksums only, as a test of coverage
        cs = torch.tensor([float(mine.sum()), float(b - a), float((back != frames[a:b]).sum())], dtype=torch.float64)
        allcs = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allcs, cs)
        t = torch.tensor([10.0 + rank], dtype=torch.float64)      # rank-dependent "elapsed ms"
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, [c.tolist() for c in allcs], float(t.item()), float(orc.rgb_to_husl(frames).sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, allcs, tmax, whole in res:
        assert tmax == 11.0                                        # max over ranks
        assert sum(c[1] for c in allcs) == 5 and [c[1] for c in allcs] == [3, 2]
        assert all(c[2] == 0 for c in allcs)                       # exact round trip on every shard
        assert abs(sum(c[0] for c in allcs) - whole) < 1e-6        # shards tile the batch


def test_frame_batches_cover_a_cyclic_stream():
    """frame i of a video stream is pool[i % P]: batches never straddle the pool's end"""
    for start, stop, P, B in ((0, 1000, 32, 8), (125, 250, 32, 8), (3, 40, 5, 4), (0, 7, 16, 8), (9, 9, 4, 2), (0, 10, 3, 8)):
        got = list(_shard.frame_batches(start, stop, P, B))
        frames = [i + j for i, p, n in got for j in range(n)]
        assert frames == list(range(start, stop))
        assert all(1 <= n <= B and p == i % P and p + n <= P for i, p, n in got)
    assert [n for _, _, n in _shard.frame_batches(0, 1000, 32, 8)] == [8] * 125
    with pytest.raises(ValueError):
        list(_shard.frame_batches(0, 4, 0, 2))


def test_stream_frames_schedules_with_fake_backend(monkeypatch):
    """stream_frames' host logic (block partition, dynamic hand-out, ring slots, `last` bookkeeping) with the
    CUDA calls replaced by numpy stand-ins -- the two world-size-2 style partitions of the video on CPU."""
    import threading

    class FakeStream:
        def __init__(self, dev):
            self.dev = dev

        def synchronize(self):
            pass

    class FakeDev:
        def __init__(self, shape, dtype, dev, stream=None):
            self.a = np.zeros(shape, dtype)

        def __getitem__(self, k):
            v = FakeDev.__new__(FakeDev)
            v.a = self.a[k]
            return v
    lock, log = threading.Lock(), []

    def fwd(src, out=None, stream=None, device=None, blocking=True):
        out.a[...] = src.astype(out.a.dtype) + 1000 * (device + 1)
        with lock:
            log.append((device, src.shape[0]))

    def inv(h, out=None, stream=None, device=None, blocking=True, edit=None):
        out[...] = (h.a - 1000 * (device + 1)).astype(np.uint8)
    for name, obj in (("Stream", FakeStream), ("DeviceArray", FakeDev), ("pinned_empty", lambda s, d: np.zeros(s, d)),
                      ("_rgb_to_husl", fwd), ("_husl_to_rgb", inv)):
        monkeypatch.setattr(_shard._cuda, name, obj)
    pool = np.random.default_rng(1).integers(0, 256, (6, 4, 5, 3), dtype=np.uint8)
    for schedule in ("static", "dynamic"):
        seen = {}
        log.clear()
        res = _shard.stream_frames(pool, 17, devices=[0, 1], first_frame=2, batch=4, dtype=np.float64, schedule=schedule,
                                   on_batch=lambda dev, i, fr: seen.update({i + j: (dev, f.copy()) for j, f in enumerate(fr)}))
        assert sorted(seen) == list(range(2, 19)) and res["frames"] == 17
        assert all(np.array_equal(f, pool[i % 6]) for i, (d, f) in seen.items())
        assert sum(v["frames"] for v in res["per_device"].values()) == 17
        if schedule == "static":                               # contiguous blocks: frames 2..10 on device 0, 11..18 on device 1
            assert {d for i, (d, f) in seen.items() if i <= 10} == {0} and {d for i, (d, f) in seen.items() if i > 10} == {1}
            assert res["per_device"][0]["frames"] == 9 and res["per_device"][1]["frames"] == 8
        for dev, (i, p, n, view) in res["last"].items():
            assert p == i % 6 and np.array_equal(view, pool[p:p + n])
    with pytest.raises(ValueError):
        _shard.stream_frames(pool, 4, devices=[0], schedule="round-robin")
