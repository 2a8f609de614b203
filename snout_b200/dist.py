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


def bind_to_gpu_numa(local_rank: int) -> int | None:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the page-locked staging buffers it
    allocates afterwards (first touch) and its copy threads sit next to the GPU's PCIe root.  With several ranks
    streaming host IQ at PCIe rate the cross-socket hop is the first thing to saturate.  Returns the node or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def plan_job(n_captures: int, n_samples: int, engine=None, units_per_shard: int = 0, geometry=None, decim: int | None = None):
    """Work units of a batch job (BASELINE config 5: many long captures, sharded by time segment
    with halo): one unit = (capture_id, time shard).  Returns a list of dicts
    {capture, lo, hi, pre_samples, body_samples, first_window} in input-rate samples, ordered by
    (capture, time).  The cut depends only on the engine geometry, never on the number of ranks,
    so the union of the results is the same for any world size."""
    from . import stream
    if geometry is None:
        geometry = stream.shard_geometry(engine.n_ble, engine.n_zb, engine.cfg.zb_segment, engine.cfg.zb_prehalo)
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
    from . import stream
    return stream.zb_span_filter(sort_reference_order(frames))      # the span rule across shard (and rank) boundaries


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


class _DevView:
    """Raw device memory as a CUDA-array-interface object (uint8 [n])."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class FrameGather:
    """Pipelined all-gather of the per-step frame records (the one exchange of the path): a count all-gather and
    ONE fixed-capacity record all-gather per step, launched asynchronously on a high-priority side stream, so that
    they overlap the next step's kernels and the host never waits inside the step.  On GPUs the send buffer is the
    engine's own HBM frame list (RxEngine.polled_frames_device(): no copy, the records never cross PCIe).  The
    gathered records stay in device memory on every rank; `Pending.counts()` reads back only the counts,
    `Pending.frames()` the records.  If a rank ever holds more than `cap` records the step falls back to the exact
    two-phase `allgather_frames` and the capacity grows.

        g = FrameGather(device); h = g.start(frames, ptr, records)   ...next step...   all_frames = h.frames()

    A Pending must be collected before `depth` further steps have been started (the engine keeps a polled frame
    list valid for two further process() calls, include/snoutrx.h).  The overflow fallback re-sends `frames` from host
    memory: pass a copy (poll(copy=True)) if a step may exceed `cap` and the engine's pinned view will be reused."""

    def __init__(self, device=None, cap: int = 4096, depth: int = 2, record_bytes: int | None = None):
        """record_bytes: leading bytes of every 160-byte record that are exchanged (multiple of 16).  BLE records carry
        at most 42 payload bytes after the 28-byte header, so 80 is enough for a BLE-only engine and halves the NVLink
        traffic; the receiving side pads with zeros.  Default: whole records."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rec = int(record_bytes or FRAME_DTYPE.itemsize)
        assert self.rec % 16 == 0 and 32 <= self.rec <= FRAME_DTYPE.itemsize
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        nccl = dist.is_initialized() and dist.get_backend() == "nccl"
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
        self.device = device
        self.cuda = device.type == "cuda"
        self.cap, self.depth, self.slot = int(cap), depth, 0
        # high priority: the collective must not queue behind the ~10^5 CTAs of the next step's channelizer
        self.stream = torch.cuda.Stream(device, priority=-1) if self.cuda else None
        self.bufs = []
        self.src_cache = {}
        self.fallbacks = 0

    def _alloc(self):
        t, rec = self.torch, self.rec
        self.bufs = []
        for _ in range(self.depth):
            b = {"send": t.zeros((self.cap, rec), dtype=t.uint8, device=self.device),
                 "recv": t.zeros((self.world, self.cap, rec), dtype=t.uint8, device=self.device),
                 "host": t.zeros((self.cap, rec), dtype=t.uint8),
                 "cnt_host": t.zeros(1, dtype=t.int64), "cnts_host": t.zeros(self.world, dtype=t.int64),
                 "cnt": t.zeros(1, dtype=t.int64, device=self.device),
                 "cnts": t.zeros(self.world, dtype=t.int64, device=self.device)}
            b["send_flat"], b["recv_flat"] = b["send"].view(-1), b["recv"].view(-1)
            if self.cuda:
                for k in ("host", "cnt_host", "cnts_host"):
                    b[k] = b[k].pin_memory()
                b["ev_copy"], b["ev_done"] = t.cuda.Event(), t.cuda.Event()
            self.bufs.append(b)

    class Pending:
        def __init__(self, g, frames, buf, src):
            self.g, self.local, self.buf, self.src = g, frames, buf, src
            self.cap = g.cap                                   # capacity of the buffers this step was sent in
            self._counts = None
            self.launched = False
            self.work = None

        def launch(self):
            """Queue the collectives (and the read-back of the counts) of a deferred start()."""
            g, b = self.g, self.buf
            if b is None or self.launched:
                return self
            self.launched = True
            if g.cuda:
                with g.torch.cuda.stream(g.stream):
                    b["cnt"].copy_(b["cnt_host"], non_blocking=True)
                    g.dist.all_gather_into_tensor(b["cnts"], b["cnt"], async_op=True).wait()      # stream-ordered: the host
                    g.dist.all_gather_into_tensor(b["recv_flat"], self.src, async_op=True).wait()  # does not block
                    b["cnts_host"].copy_(b["cnts"], non_blocking=True)
                    b["ev_done"].record(g.stream)
            else:
                b["cnt"][0] = len(self.local)
                self.work = [g.dist.all_gather_into_tensor(b["cnts"], b["cnt"], async_op=True),
                             g.dist.all_gather_into_tensor(b["recv_flat"], self.src, async_op=True)]
            return self

        def counts(self):
            """Frames per rank (reads back the counts only)."""
            self.launch()
            if self._counts is None:
                g = self.g
                if g.world == 1:
                    self._counts = [len(self.local)]
                elif g.cuda:
                    self.buf["ev_done"].synchronize()          # the collectives were queued a step ago
                    self._counts = self.buf["cnts_host"].tolist()
                else:
                    for w in self.work:
                        w.wait()
                    self._counts = self.buf["cnts"].tolist()
            return self._counts

        def frames(self):
            """All ranks' frames concatenated in rank order."""
            g = self.g
            if g.world == 1:
                return self.local
            c = self.counts()
            if max(c) > self.cap:                                  # some rank overflowed the fixed buffers (every rank sees
                g.fallbacks += 1                                   # the same counts): exact two-phase path, larger buffers
                if 2 * max(c) > g.cap:
                    g.cap = 2 * max(c)
                    g.bufs = []
                return allgather_frames(self.local, g.device)
            recv = self.buf["recv"]
            full = FRAME_DTYPE.itemsize
            out = []
            for r in range(g.world):
                a = recv[r, :c[r]].cpu().numpy()
                if g.rec != full:                                  # packed exchange: the tail of every record is zero
                    z = np.zeros((c[r], full), dtype=np.uint8)
                    z[:, :g.rec] = a
                    a = z
                out.append(a.reshape(-1).view(FRAME_DTYPE))
            return np.concatenate(out) if out else self.local[:0]

    def start(self, frames: np.ndarray, device_ptr: int = 0, device_records: int = 0, defer: bool = False) -> "FrameGather.Pending":
        """Launch the gather of this step's frames.  `device_ptr` / `device_records`: device address and capacity (in
        records) of a device buffer that starts with the same records (RxEngine.polled_frames_device()); when it holds
        at least `cap` records it is the send buffer itself.  Otherwise the records are copied device-to-device, or
        staged from host memory when there is no device copy.  With defer=True the collectives are queued by
        Pending.launch() -- the caller can queue its next batch first."""
        t = self.torch
        if self.world == 1:
            return FrameGather.Pending(self, frames, None, None)
        n = len(frames)
        if not self.bufs:
            self._alloc()
        b = self.bufs[self.slot]
        self.slot = (self.slot + 1) % self.depth
        m = min(n, self.cap)
        rec, full = self.rec, FRAME_DTYPE.itemsize
        src = b["send_flat"]

        def packed(fr):                                             # host records -> [m, rec] uint8
            return t.from_numpy(np.ascontiguousarray(fr[:m]).view(np.uint8).reshape(m, full)[:, :rec].copy())
        if self.cuda:
            b["cnt_host"][0] = n
            dev = None
            if device_ptr:
                key = (device_ptr, device_records)
                dev = self.src_cache.get(key)
                if dev is None:
                    dev = t.as_tensor(_DevView(device_ptr, max(device_records, m) * full), device=self.device).view(-1, full)
                    self.src_cache[key] = dev
            if dev is not None and device_records >= self.cap and rec == full:
                src = dev[: self.cap].view(-1)                      # zero copy: the engine's frame list is the send buffer
            elif m:
                with t.cuda.stream(self.stream):
                    if dev is not None:                             # (packing) device-to-device copy
                        b["send"][:m].copy_(dev[:m, :rec], non_blocking=True)
                    else:
                        b["host"][:m] = packed(frames)
                        b["send"][:m].copy_(b["host"][:m], non_blocking=True)
                    b["ev_copy"].record(self.stream)
                if dev is None:
                    b["ev_copy"].synchronize()    # the staging rows are free again from here on
        elif m:
            b["send"][:m] = packed(frames)
        p = FrameGather.Pending(self, frames, b, src)
        if not defer:
            p.launch()
        return p


class PeerGather:
    """The same exchange as FrameGather without a collective kernel: every rank pushes its [header | records] block (the
    engine writes the header in front of its HBM frame list, include/snoutrx.h) straight into a receive slot in every
    peer's HBM -- peer-to-peer copies over NVLink / NVSwitch on the copy engines, through buffers mapped with
    torch.distributed._symmetric_memory -- followed by a stream-ordered signal per peer.  No SM-resident kernel waits for
    the slowest rank beside the channelizer (an NCCL all-gather does: DESIGN.md 9).  Same interface as FrameGather
    (start(frames, device_ptr, device_records, defer) -> Pending with launch() / counts() / frames()); `available()` tells
    whether the symmetric-memory rendezvous works on this system -- callers fall back to FrameGather otherwise."""

    SLOTS = 4

    def __init__(self, device, cap: int = 1 << 15):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch, self.dist = torch, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.device, self.cap, self.step = device, int(cap), 0
        self.rec = FRAME_DTYPE.itemsize
        self.slot_bytes = (self.cap + 1) * self.rec
        self.buf = symm.empty(self.world * self.SLOTS * self.slot_bytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        shape = (self.world, self.SLOTS, self.slot_bytes)
        self.views = [self.hdl.get_buffer(r, shape, torch.uint8) for r in range(self.world)]     # views[p][src, slot, byte]
        self.stream = torch.cuda.Stream(device, priority=-1)
        self.hdr_host = [torch.zeros((self.world, 16), dtype=torch.uint8).pin_memory() for _ in range(self.SLOTS)]
        self.ev_done = [torch.cuda.Event() for _ in range(self.SLOTS)]
        self.src_cache = {}
        self.fallbacks = 0
        torch.cuda.synchronize(device)
        dist.barrier()

    @staticmethod
    def available(device) -> "PeerGather | None":
        """Try to set the exchange up on every rank; all ranks agree on the outcome."""
        import torch
        import torch.distributed as dist
        g, ok = None, 1
        try:
            g = PeerGather(device)
        except Exception:                                         # no symmetric memory on this system / torch build
            ok = 0
        t = torch.tensor([ok], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return g if int(t.item()) == 1 else None

    class Pending:
        def __init__(self, g, frames, slot, step, src, nbytes):
            self.g, self.local, self.slot, self.step, self.src, self.nbytes = g, frames, slot, step, src, nbytes
            self.launched, self._counts = False, None

        def launch(self):
            g = self.g
            if self.launched:
                return self
            self.launched = True
            t = g.torch
            with t.cuda.stream(g.stream):
                for p in range(g.world):                                   # [header | records] into slot (me, step) of every rank
                    g.views[p][g.rank, self.slot, : self.nbytes].copy_(self.src[: self.nbytes], non_blocking=True)
                for p in range(g.world):
                    if p != g.rank:
                        g.hdl.put_signal(p)                                # my block of this step has landed at p
                for p in range(g.world):
                    if p != g.rank:
                        g.hdl.wait_signal(p)                               # p's block of this step has landed here
                g.hdr_host[self.slot].copy_(g.views[g.rank][:, self.slot, :16], non_blocking=True)
                g.ev_done[self.slot].record(g.stream)
            return self

        def counts(self):
            self.launch()
            if self._counts is None:
                g = self.g
                g.ev_done[self.slot].synchronize()
                h = g.hdr_host[self.slot].numpy().view(np.uint64).reshape(g.world, 2)
                self._counts = [int(c) for c in h[:, 0]]
            return self._counts

        def frames(self):
            g = self.g
            c = self.counts()
            if max(c) > g.cap:                                             # a rank overflowed its slot: exact two-phase gather
                g.fallbacks += 1
                return allgather_frames(self.local, g.device)
            own = g.views[g.rank]
            out = [own[r, self.slot, g.rec: g.rec * (1 + c[r])].cpu().numpy().view(FRAME_DTYPE) for r in range(g.world)]
            return np.concatenate(out) if out else self.local[:0]

    def start(self, frames: np.ndarray, device_ptr: int = 0, device_records: int = 0, defer: bool = False) -> "PeerGather.Pending":
        """`device_ptr` / `device_records`: RxEngine.polled_frames_device() -- the engine's HBM list, whose header record
        sits 160 bytes in front of it.  Collect a Pending before SLOTS - 1 further steps have been started."""
        t = self.torch
        if not device_ptr:
            raise ValueError("PeerGather sends the engine's device frame list: pass RxEngine.polled_frames_device()")
        n = len(frames)
        m = min(n, self.cap)
        key = (device_ptr, device_records)
        src = self.src_cache.get(key)
        if src is None:
            src = t.as_tensor(_DevView(device_ptr - self.rec, (device_records + 1) * self.rec), device=self.device)
            self.src_cache[key] = src
        slot, step = self.step % self.SLOTS, self.step
        self.step += 1
        p = PeerGather.Pending(self, frames, slot, step, src, (m + 1) * self.rec)
        if not defer:
            p.launch()
        return p


class AbiGather:
    """The frame exchange of the C ABI (include/snoutrx.h snrx_exchange_create / _connect / snrx_allgather): once the
    engines of the node are connected, the kernel that exports a batch also stores its records and a {count, batch} header
    into every rank's HBM over NVLink -- no collective kernel, nothing launched per step on the host.  torch.distributed is
    only used once, to pass the 64-byte CUDA IPC handles around.  Same interface as FrameGather / PeerGather: start() after
    a poll() -> Pending with counts() / frames().  Every rank must start() every batch, in order, and collect a Pending
    before it queues four further batches."""

    def __init__(self, engine, device, cap: int = 1 << 15):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.eng, self.device = torch, dist, engine, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.cap, self.fallbacks = int(cap), 0
        mine = engine.exchange_create(self.rank, self.world, self.cap)
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
        allh = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(allh, t)
        engine.exchange_connect(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
        dist.barrier()

    @staticmethod
    def available(engine, device) -> "AbiGather | None":
        """Try to connect the engines of all ranks; every rank agrees on the outcome (CUDA IPC may be unavailable)."""
        import torch
        import torch.distributed as dist
        g, ok = None, 1
        try:
            g = AbiGather(engine, device)
        except Exception:
            ok = 0
        t = torch.tensor([ok], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return g if int(t.item()) == 1 else None

    class Pending:
        def __init__(self, g, frames, batch_no):
            self.g, self.local, self.batch_no, self._counts = g, frames, batch_no, None

        def launch(self):
            return self

        def counts(self):
            if self._counts is None:
                self._counts, _ = self.g.eng.allgather(self.batch_no)
            return self._counts

        def frames(self):
            self._counts, fr = self.g.eng.allgather(self.batch_no, want_frames=True)
            if fr is None:                                         # a rank exceeded the slot capacity (all ranks see it)
                self.g.fallbacks += 1
                fr = allgather_frames(self.local, self.g.device)
            return fr

    def start(self, frames: np.ndarray, device_ptr: int = 0, device_records: int = 0, defer: bool = False) -> "AbiGather.Pending":
        return AbiGather.Pending(self, frames, self.eng.polled_batch_no)


def sort_reference_order(frames: np.ndarray) -> np.ndarray:
    order = np.lexsort((frames["sample_index"], frames["window"], frames["channel"],
                        255 - frames["proto"].astype(np.int32), frames["capture_id"]))
    return frames[order]
