#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun on N GPUs of one box):
  1. every rank decodes its own small wideband capture; the pipelined device-side gather (dist.FrameGather, fed from
     the engine's HBM frame list) must return exactly what the two-phase host gather returns, for several steps;
  2. dist.run_job over time shards of two captures: the frames of the whole job are identical on every rank and
     equal to what rank 0 computes alone (world 1)."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from snout_b200 import dist as sdist, synth
    from snout_b200.engine import RxEngine
    rank, world, local = sdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cap = synth.wideband_capture(seconds=0.02, kind="ble", seed=7000 + rank, esn0_db=25.0, gap=(300, 3000))
    x = cap.iq[: len(cap.iq) - 24 * 1000 * rank]                 # ragged frame counts across ranks
    g = sdist.FrameGather(dev, cap=64, record_bytes=80)                             # small: the first step overflows and falls back
    with RxEngine("ble_wb40", max_samples=len(cap.iq), device=local) as eng:
        pend, got = None, []
        for step in range(4):
            fr = eng.process(x).poll(copy=True)
            h = g.start(fr, *eng.polled_frames_device()[:2])
            if pend is not None:
                got.append(pend.frames())
            pend = h
            want = sdist.allgather_frames(fr, dev)
        got.append(pend.frames())
        for k, f in enumerate(got):
            assert f.tobytes() == want.tobytes(), (rank, k, len(f), len(want))
        counts = pend.counts()
        assert len(counts) == world and sum(counts) == len(want)

        # the same exchange as peer-to-peer pushes on the copy engines (when symmetric memory is available here)
        pg = sdist.PeerGather.available(dev)
        peer_note = "PeerGather unavailable (NCCL path only)"
        if pg is not None:
            pend, got = None, []
            for step in range(5):
                fr = eng.process(x).poll(copy=True)
                h = pg.start(fr, *eng.polled_frames_device()[:2], defer=(step % 2 == 1)).launch()
                if pend is not None:
                    got.append(pend.frames())
                pend = h
            got.append(pend.frames())
            for k, f in enumerate(got):
                assert f.tobytes() == want.tobytes(), ("peer", rank, k, len(f), len(want))
            assert pend.counts() == counts
            peer_note = f"PeerGather == exact gather over {len(got)} steps"

        # the exchange of the C ABI: records stored into every rank's HBM by the export kernel (snrx_exchange_* / snrx_allgather)
        ag = sdist.AbiGather.available(eng, dev)
        abi_note = "AbiGather unavailable (CUDA IPC)"
        if ag is not None:
            pend, got = None, []
            for step in range(6):
                xs = x if step % 2 == 0 else x[: len(x) - 24 * 4000]         # the record count changes from step to step
                fr = eng.process(xs).poll(copy=True)
                h = ag.start(fr)
                w = sdist.allgather_frames(fr, dev)
                if pend is not None:
                    got.append((pend[0].frames(), pend[1]))
                pend = (h, w)
            got.append((pend[0].frames(), pend[1]))
            for k, (f, w) in enumerate(got):
                assert f.tobytes() == w.tobytes(), ("abi", rank, k, len(f), len(w))
            assert sum(pend[0].counts()) == len(pend[1])
            abi_note = f"AbiGather == exact gather over {len(got)} steps"
        peer_note = peer_note + "; " + abi_note

        # time-sharded job over two captures
        caps = [synth.wideband_capture(seconds=0.03, kind="ble", seed=7100 + c, esn0_db=25.0, gap=(300, 3000)).iq for c in range(2)]
    with RxEngine("ble_wb40", max_samples=24 * (8192 * 3 + 128 + 2048), device=local) as eng:
        units = sdist.plan_job(2, len(caps[0]), eng, units_per_shard=3)
        job = sdist.run_job(eng, lambda c: caps[c], units, rank, world, device=dev)
        blob = torch.tensor([zlib.crc32(job.tobytes())], dtype=torch.int64, device=dev)
        all_blobs = [torch.zeros_like(blob) for _ in range(world)]
        dist.all_gather(all_blobs, blob)
        assert len({int(b.item()) for b in all_blobs}) == 1, "ranks disagree on the job result"
        if rank == 0:
            alone = sdist.run_job(eng, lambda c: caps[c], units, 0, 1, gather=False)
            assert alone.tobytes() == job.tobytes(), (len(alone), len(job))
            print(f"multi-gpu check ok: world {world}, {peer_note}; gather steps {len(got)} (fallbacks {g.fallbacks}), "
                  f"job frames {len(job)} identical on every rank and to world 1")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
