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


def plan_time_shards(n_samples: int, decim: int, windows_per_shard: int, post_channel_samples: int = 2048,
                     pre_channel_samples: int = 128):
    """Cut one capture of n_samples input samples into time shards aligned to the 8192-sample BLE
    window grid (= Zigbee segment grid when zb_segment divides 8192*k).  Returns a list of dicts
    {lo, hi, pre_samples, body_samples, first_window} in input-rate samples."""
    n_ch = n_samples // decim
    n_win = (n_ch + chanplan.BLE_WINDOW - 1) // chanplan.BLE_WINDOW
    shards = []
    for w0 in range(0, n_win, windows_per_shard):
        w1 = min(n_win, w0 + windows_per_shard)
        lo_ch = max(0, w0 * chanplan.BLE_WINDOW - pre_channel_samples)
        hi_ch = min(n_ch, w1 * chanplan.BLE_WINDOW + post_channel_samples)
        body_ch = min(n_ch, w1 * chanplan.BLE_WINDOW) - w0 * chanplan.BLE_WINDOW
        shards.append(dict(lo=lo_ch * decim, hi=hi_ch * decim, pre_samples=(w0 * chanplan.BLE_WINDOW - lo_ch) * decim,
                           body_samples=body_ch * decim, first_window=w0))
    return shards


def assign_round_robin(n_units: int, rank: int, world: int):
    return list(range(rank, n_units, world))


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
