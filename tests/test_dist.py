"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo process groups, a content-digest engine in
place of the CUDA engine.  The job result (frames in reference order) must be identical for every
world size and equal to the single-process result; the all-gather carries ragged counts."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from snout_b200 import _abi, dist as sdist, stream


def _captures(c):
    rng = np.random.default_rng(100 + c)
    n = 8192 * 21 + 1234
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def _job(rank, world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import ContentEngine
    eng = ContentEngine("ble_nb", channel=37, max_samples=8192 * 4 + 128 + 2048)
    units = sdist.plan_job(3, len(_captures(0)), eng, units_per_shard=4)
    return sdist.run_job(eng, _captures, units, rank, world), units, eng


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    r, w, _ = sdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    frames, units, eng = _job(rank, world)
    q.put((rank, frames.tobytes(), len(eng.calls)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_job_result_is_identical_for_any_world_size(world):
    import torch.multiprocessing as mp
    single, units, eng1 = _job(0, 1)
    assert len(units) == 3 * 6 and len(eng1.calls) == len(units)
    # every window of every capture exactly once, digest of the right samples
    assert len(single) == 3 * 22
    for c in range(3):
        x = _captures(c)
        fc = single[single["capture_id"] == c]
        assert list(fc["window"]) == list(range(22))
        for f in fc:
            s = int(f["window"]) * 8192
            assert bytes(f["bytes"][: f["len"]]) == x[s: s + 2].tobytes()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    calls = [r[2] for r in res]
    assert sum(calls) == len(units) and max(calls) - min(calls) <= 1          # round robin, no duplicates
    for _, blob, _ in res:
        assert blob == single.tobytes()                                       # identical on every rank, = world 1


def test_allgather_single_process_is_identity():
    f = np.zeros(3, _abi.FRAME_DTYPE)
    assert sdist.allgather_frames(f) is f
    assert sdist.assign_round_robin(7, 1, 3) == [1, 4]


def test_sort_reference_order():
    f = np.zeros(5, _abi.FRAME_DTYPE)
    f["capture_id"] = [1, 0, 0, 0, 0]
    f["proto"] = [3, 2, 3, 3, 3]
    f["channel"] = [0, 11, 5, 5, 4]
    f["window"] = [0, 0, 2, 1, 9]
    s = sdist.sort_reference_order(f)
    assert list(zip(s["capture_id"], s["proto"], s["channel"], s["window"])) == \
           [(0, 3, 4, 9), (0, 3, 5, 1), (0, 3, 5, 2), (0, 2, 11, 0), (1, 3, 0, 0)]


def test_plan_job_does_not_depend_on_world():
    from fake_engine import ContentEngine
    from snout_b200 import stream
    unit, pre, post = stream.shard_geometry(40, 16, 65536, 4096)
    eng = ContentEngine("mixed_wb56", max_samples=(2 * unit + pre + post) * 24, zb_segment=65536, zb_prehalo=4096)
    units = sdist.plan_job(2, 24 * 65536 * 5 + 240, eng)
    assert [u["first_window"] for u in units if u["capture"] == 0] == [0, 16, 32]
    assert all(u["pre_samples"] in (0, pre * 24) for u in units)
    # a capture whose length is not a multiple of the decimation: every shard still is
    odd = sdist.plan_job(1, 24 * 65536 * 3 + 7, eng)
    assert all((u["hi"] - u["lo"]) % 24 == 0 for u in odd) and odd[-1]["hi"] == 24 * 65536 * 3
    cover = sorted(i for w in (4,) for r in range(w) for i in sdist.assign_round_robin(len(units), r, w))
    assert cover == list(range(len(units)))


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    sdist.init_from_env(backend="gloo")
    g = sdist.FrameGather(cap=8, record_bytes=int(os.environ.get('SNRX_TEST_RECORD_BYTES', '160')))

    def frames_of(r, step):
        n = [3, 7, 0, 20][(r + step) % 4]                    # ragged, one empty, one beyond the capacity of 8
        f = np.zeros(n, _abi.FRAME_DTYPE)
        f["capture_id"], f["sample_index"], f["channel"] = r, np.arange(n) + 1000 * step, step
        f["bytes"][:, 0] = r + 1
        return f

    out, pend = [], None
    for step in range(5):                                    # two steps in flight: start step i, collect step i-1
        h = g.start(frames_of(rank, step), defer=(step % 2 == 1))     # deferred: copy now, collective on launch()
        h.launch()
        if pend is not None:
            out.append((pend.counts(), pend.frames().tobytes()))
        pend = h
    out.append((pend.counts(), pend.frames().tobytes()))
    want = [np.concatenate([frames_of(r, s) for r in range(world)]).tobytes() for s in range(5)]
    q.put((rank, [o[1] for o in out] == want, [o[0] for o in out], g.fallbacks))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("record_bytes", [160, 80])
def test_pipelined_frame_gather_world2(record_bytes, monkeypatch):
    import torch.multiprocessing as mp
    monkeypatch.setenv("SNRX_TEST_RECORD_BYTES", str(record_bytes))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, counts, fallbacks in res:
        assert same, rank
        assert counts[0] == [3, 7] and counts[3] == [20, 3]
        assert fallbacks >= 1                                 # the 20-frame step went through the exact two-phase path
    single = sdist.FrameGather()
    f = np.zeros(2, _abi.FRAME_DTYPE)
    assert single.start(f).frames() is f and single.start(f).counts() == [2]
