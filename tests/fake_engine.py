"""Stand-ins for RxEngine used by the CPU tests of the HOST logic (streaming, sharding, multi-rank
gather).  They contain no DSP: RecordingEngine records calls; ContentEngine "decodes" one frame per
8192-sample window of a shard body whose bytes are a digest of the samples at that window, so a test
can prove that every window of every capture was processed exactly once, from the right samples, no
matter how the job was cut or how many ranks shared it."""
import numpy as np

from snout_b200 import _abi

MARK = [0x40, 6, 1, 2, 3, 4, 5, 6, 0xAA, 0xBB, 0xCC]       # ADV_IND, TxAdd=1, PloadL6, AdvA, CRC


class _Buf:
    def __init__(self, n):
        self.array = np.zeros(n, np.complex64)

    def free(self):
        self.array = None


class RecordingEngine:
    """Quacks like RxEngine for ShardStreamer: records every process() call, checks the queue
    discipline of snrx_process / snrx_poll, returns one marker frame per shard."""

    def __init__(self, mode="ble_nb", channel=37, max_samples=0, zb_segment=0, zb_prehalo=0, **kw):
        self.mode, self.channel, self.kw = mode, channel, kw
        self.wideband = mode in ("ble_wb40", "zb_wb16", "mixed_wb56")
        self.decim = 24 if self.wideband else 1
        self.n_ble = {"ble_nb": 1, "ble_wb40": 40, "mixed_wb56": 40}.get(mode, 0)
        self.n_zb = {"zb_nb": 1, "zb_wb16": 16, "mixed_wb56": 16}.get(mode, 0)
        self.cfg = _abi.Config()
        self.cfg.max_samples, self.cfg.zb_segment, self.cfg.zb_prehalo = max_samples, zb_segment, zb_prehalo
        self.calls, self.queue, self.max_queue, self.closed = [], [], 0, False

    def frames_for(self, iq, shard):
        f = np.zeros(1, _abi.FRAME_DTYPE)
        body0 = shard["first_window"] * 8192
        f["sample_index"], f["window"], f["channel"] = body0, shard["first_window"], self.channel
        f["proto"] = 3 if self.n_ble else 2
        f["len"], f["crc_ok"], f["access_addr"], f["lqi"] = 11, 1, 0x8E89BED6, 255
        f["bytes"][0, :11] = MARK
        return f

    def process(self, iq, shard=None):
        assert len(self.queue) < 2, "third batch queued"
        assert len(iq) <= self.cfg.max_samples and len(iq) % self.decim == 0
        self.calls.append((np.array(iq, copy=True), dict(shard)))
        self.queue.append(self.frames_for(iq, shard))
        self.max_queue = max(self.max_queue, len(self.queue))
        return self

    def poll(self, copy=True):
        return self.queue.pop(0)

    def alloc_host(self, n):
        return _Buf(n)

    def close(self):
        self.closed = True


class ContentEngine(RecordingEngine):
    """One frame per 8192-sample window of the body; bytes = digest of the window's first samples."""

    def frames_for(self, iq, shard):
        d = self.decim
        pre = shard["pre_samples"] // d
        n_ch = len(iq) // d
        body = shard["body_samples"] // d if shard["body_samples"] else n_ch - pre
        nw = -(-body // 8192)
        f = np.zeros(nw, _abi.FRAME_DTYPE)
        for w in range(nw):
            s = (pre + w * 8192) * d
            dig = np.frombuffer(np.ascontiguousarray(iq[s: s + 2]).tobytes(), np.uint8)
            f[w]["sample_index"] = (shard["first_window"] + w) * 8192
            f[w]["window"] = shard["first_window"] + w
            f[w]["capture_id"] = shard.get("first_capture_id", 0)
            f[w]["channel"], f[w]["proto"], f[w]["crc_ok"] = self.channel, 3, 1
            f[w]["len"] = len(dig)
            f[w]["bytes"][: len(dig)] = dig
        return f
