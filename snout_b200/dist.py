"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The receive path shards by independent units -- captures, or time shards of a capture with a halo
(include/snoutrx.h snrx_shard_t) -- so there is NO collective inside the DSP.  The only exchange is
the all-gather of the fixed-size frame records at the end of a batch (BASELINE north_star: "NCCL
over NVLink is used only to allgather the decoded-frame records")."""
from __future__ import annotations

import os

import numpy as np

from ._abi import FRAME_DTYPE
from . import chanplan


def init_from_env(backend: str | None = None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
    (rank, world, local_rank); a single process without those variables is (0, 1, 0)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def plan_job(n_captures: int, n_samples: int, engine=None, units_per_shard: int = 0, geometry=None, decim: int | None = None):
    """Work units of a batch job (BASELINE config 5: many long captures, sharded by time segment
    with halo): one unit = (capture_id, time shard).  Returns a list of dicts
    {capture, lo, hi, pre_samples, body_samples, first_window} in input-rate samples, ordered by
    (capture, time).  The cut depends only on the engine geometry, never on the number of ranks,
    so the union of the results is the same for any world size."""
    from . import stream
    if geometry is None:
        geometry = stream.shard_geometry(engine.n_ble, engine.n_zb, engine.cfg.zb_segment or 8192, engine.cfg.zb_prehalo or 4096)
    if decim is None:
        decim = engine.decim
    unit, pre, post = geometry
    if not units_per_shard:
        cap_in = int(engine.cfg.max_samples) or (96_000_000 if engine.wideband else 10_000_000)
        units_per_shard = max(1, (cap_in // decim - pre - post) // unit)
    units = []
    for c in range(n_captures):
        for sh in stream.plan_shards(n_samples, decim, unit, pre, post, units_per_shard):
            units.append(dict(capture=c, **sh))
    return units


def assign_round_robin(n_units: int, rank: int, world: int):
    """Unit u belongs to rank u mod world."""
    return list(range(rank, n_units, world))


def run_job(engine, captures, units, rank: int = 0, world: int = 1, gather: bool = True, device=None) -> np.ndarray:
    """Process this rank's share of `units` and return the frames of the WHOLE job in reference order
    (identical on every rank and for every world size).  `captures(c)` returns capture c as a
    complex64 array (host) -- only called for captures this rank needs.  Two shards are kept in flight."""
    mine = [units[i] for i in assign_round_robin(len(units), rank, world)]
    out, pending = [], 0
    cache, alive = {}, []                # `alive`: buffers of the batches still in flight (host->device copies are async)
    for u in mine:
        c = u["capture"]
        if c not in cache:
            cache.clear()
            cache[c] = captures(c)
        x = cache[c][u["lo"]: u["hi"]]
        if pending == 2:
            out.append(engine.poll())
            alive.pop(0)
            pending -= 1
        alive.append(x)
        engine.process(x, shard=dict(pre_samples=u["pre_samples"], body_samples=u["body_samples"],
                                     first_window=u["first_window"], first_capture_id=c))
        pending += 1
    while pending:
        out.append(engine.poll())
        pending -= 1
    frames = np.concatenate(out) if out else np.zeros(0, dtype=FRAME_DTYPE)
    if gather and world > 1:
        frames = allgather_frames(frames, device)
    return sort_reference_order(frames)


def allgather_frames(frames: np.ndarray, device=None) -> np.ndarray:
    """All-gather frame records over the default process group: counts first, then records padded
    to the maximum count.  Returns every rank's frames concatenated in rank order (callers sort by
    (capture_id, channel, window, sample_index) to obtain reference order)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return frames
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    cnt = torch.tensor([len(frames)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    buf = np.zeros(mx, dtype=FRAME_DTYPE)
    buf[: len(frames)] = frames
    send = torch.from_numpy(buf.view(np.uint8).reshape(mx, FRAME_DTYPE.itemsize)).to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    out = [r.cpu().numpy().reshape(-1).view(FRAME_DTYPE)[:c] for r, c in zip(recv, counts)]
    return np.concatenate(out) if out else frames[:0]


def sort_reference_order(frames: np.ndarray) -> np.ndarray:
    order = np.lexsort((frames["sample_index"], frames["window"], frames["channel"],
                        255 - frames["proto"].astype(np.int32), frames["capture_id"]))
    return frames[order]
