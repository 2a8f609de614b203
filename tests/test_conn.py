"""SURVEY 8(f) row N3 -- BLE connections (btle_rx -o: CONNECT_REQ fields btle_rx.c:1476-1557, receiver_controller :2167-2282).

CPU part: the kernel's per-thread parser (csrc/ble_conn.cuh, stepped on the host through tests/emu) is pinned against the
UNMODIFIED reference's own field extraction on committed vectors (tests/golden/btle_connreq_ref.json, made by
make_golden_conn.py through oracle/_ref); the `-o` logic of the btle_rx drop-in is driven with a scripted engine.
GPU part: on a 96 Msps capture of a connection being opened the engine finds the request, and the second search of the
batch's bit streams with the learned access address / CRC init equals the oracle run with those `-a` / `-k` values on the
engine's own quantised channel streams, channel by channel -- bit exact."""
import ctypes
import io
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_frames_equal
from snout_b200 import _abi, btle_cli, chanplan, synth


def _golden():
    return json.load(open(os.path.join(GOLDEN, "btle_connreq_ref.json")))


def _parse(emu, pdu: bytes):
    out = np.zeros(1, dtype=_abi.CONN_DTYPE)
    buf = np.frombuffer(bytes(pdu) + bytes(64), dtype=np.uint8).copy()
    ok = emu.emu_ble_conn_parse(buf.ctypes.data_as(ctypes.c_void_p), len(pdu), out.ctypes.data_as(ctypes.c_void_p))
    return ok, out[0]


def test_conn_parse_equals_reference_fields(emu):
    g = _golden()
    assert len(g["rows"]) >= 24 and g["rc_for_33_byte_payload"] == -1
    full = 0
    for r in g["rows"]:
        payload = bytes.fromhex(r["payload"])
        ok, c = _parse(emu, bytes([0x45, 34]) + payload + b"\0\0\0")
        assert ok == 1
        for k in ("access_addr", "crc_init", "win_size", "win_offset", "interval", "latency", "timeout", "hop", "sca", "chm_full"):
            assert int(c[k]) == r[k], (k, r["payload"])
        assert int(c["hop"]) == r["status_hop"] and int(c["interval"]) == r["status_interval"]
        # the reference keeps the addresses and the map byte-reversed (for printing); the record holds them as transmitted
        assert bytes(c["init_a"])[::-1].hex() == r["init_a_reversed"]
        assert bytes(c["adv_a"])[::-1].hex() == r["adv_a_reversed"]
        assert bytes(c["chm"])[::-1].hex() == r["chm_reversed"]
        full += r["chm_full"]
    assert 0 < full < len(g["rows"])


def test_conn_parse_rejects_what_the_reference_rejects(emu):
    p = bytes.fromhex(_golden()["rows"][0]["payload"])
    assert _parse(emu, bytes([0x45, 33]) + p[:33] + b"\0\0\0")[0] == 0          # wrong payload length (btle_rx.c:1477-1480)
    assert _parse(emu, bytes([0x40, 34]) + p + b"\0\0\0")[0] == 0               # ADV_IND of the same length
    assert _parse(emu, bytes([0x45, 34]) + p + b"\0\0\0")[0] == 1
    assert _parse(emu, bytes([0xC5, 34]) + p + b"\0\0\0")[0] == 1               # TxAdd / RxAdd do not matter


def test_synth_connect_req_is_what_the_reference_parses(emu):
    pdu = synth.ble_connect_req(bytes(range(6)), bytes(range(6, 12)), 0x50655A3B, 0x1A2B3C, hop=7, interval=6)
    ok, c = _parse(emu, pdu + b"\0\0\0")
    assert ok and int(c["access_addr"]) == 0x50655A3B and int(c["crc_init"]) == 0x1A2B3C and int(c["hop"]) == 7
    assert int(c["interval"]) == 6 and int(c["chm_full"]) == 1 and bytes(c["init_a"]) == bytes(range(6))


def test_hop_sequence():
    assert chanplan.ble_hop_channels(7, 6) == [7, 14, 21, 28, 35, 5]            # (0 + 7k) % 37, btle_rx.c:2194,2227
    assert chanplan.ble_hop_channels(16, 3, last=30) == [9, 25, 4]


# ------------------------------------------------------------------ the -o logic of the drop-in, scripted engine
class _ConnEngine:
    """Quacks like RxEngine for btle_cli.run: one shard, scripted records."""

    def __init__(self, *a, **kw):
        from fake_engine import RecordingEngine
        self._e = RecordingEngine("ble_wb40", channel=kw.get("channel", 37), max_samples=kw.get("max_samples", 0))
        self.__dict__.update({k: getattr(self._e, k) for k in ("wideband", "decim", "n_ble", "n_zb", "cfg")})
        self.followed = []

    @staticmethod
    def _frame(ch, idx, pdu, aa):
        f = np.zeros(1, _abi.FRAME_DTYPE)
        f["sample_index"], f["channel"], f["proto"], f["crc_ok"], f["access_addr"] = idx, ch, 3, 1, aa
        b = bytes(pdu) + b"\x11\x22\x33"
        f["len"] = len(b)
        f["bytes"][0, : len(b)] = np.frombuffer(b, np.uint8)
        return f

    def process(self, iq, shard=None):
        return self._e.process(iq, shard)

    def poll(self, copy=True):
        self._e.poll()
        self.polls = getattr(self, "polls", 0) + 1
        if self.polls > 1:
            return np.zeros(0, _abi.FRAME_DTYPE)
        adv = bytes([0x00, 9]) + bytes(range(6)) + b"\x02\x01\x06"
        self.partial = synth.ble_connect_req(bytes(6), bytes(range(6)), 0x11111111, 0x000001, 3, chm=b"\xff\xff\xff\xff\x0f")
        self.req = synth.ble_connect_req(bytes(6), bytes(range(6)), 0x50655A3B, 0x1A2B3C, 7)
        return np.concatenate([self._frame(37, 1000, adv, 0x8E89BED6), self._frame(37, 3000, self.partial, 0x8E89BED6),
                               self._frame(37, 6000, self.req, 0x8E89BED6), self._frame(37, 9000, adv, 0x8E89BED6),
                               self._frame(38, 9500, adv, 0x8E89BED6)])

    def connections(self):
        out = np.zeros(2, _abi.CONN_DTYPE)
        for o, (idx, aa, ci, hop, chm, full, frame) in zip(out, [(3000, 0x11111111, 1, 3, b"\xff\xff\xff\xff\x0f", 0, 1),
                                                                 (6000, 0x50655A3B, 0x1A2B3C, 7, b"\xff\xff\xff\xff\x1f", 1, 2)]):
            o["sample_index"], o["access_addr"], o["crc_init"], o["hop"], o["chm_full"], o["channel"], o["frame"] = idx, aa, ci, hop, full, 37, frame
            o["chm"] = np.frombuffer(chm, np.uint8)
        return out

    def follow(self, aa, ci):
        self.followed.append((aa, ci))
        if len(self.followed) > 1:
            return np.zeros(0, _abi.FRAME_DTYPE)
        data = bytes([0x02, 3, 1, 2, 3])
        return np.concatenate([self._frame(7, 5000, data, aa),            # before the request: not part of the connection
                               self._frame(7, 12000, data, aa), self._frame(14, 42000, bytes([0x01, 0]), aa)])

    def alloc_host(self, n):
        return self._e.alloc_host(n)

    def close(self):
        pass


def _run_cli(argv, n=24 * 8192 * 2):
    o = btle_cli.parse_commandline(argv)
    eng = {}

    def factory(*a, **kw):
        eng["e"] = _ConnEngine(*a, **kw)
        return eng["e"]

    out = io.StringIO()
    rc = btle_cli.run(o, out=out, engine_factory=factory, blocks=[np.zeros(n, np.complex64)])
    return rc, out.getvalue().splitlines(), eng.get("e")


def test_cli_hop_follows_the_connection():
    rc, lines, eng = _run_cli(["-c", "37", "-o", "--wideband", "--shard-windows", "2"])
    assert rc == 0
    body = [" ".join(l.split(" ")[1:]) if l[0].isdigit() else l for l in lines[1:-1]]
    assert eng.followed and set(eng.followed) == {(0x50655A3B, 0x1A2B3C)}     # every later shard is searched too
    assert body[0].startswith("Pkt1 Ch37 AA:8e89bed6 ADV_PDU_t0:ADV_IND")
    assert body[1].startswith("Pkt2 Ch37 AA:8e89bed6 ADV_PDU_t5:CONNECT_REQ")
    assert body[2] == "Hop: Not full ChnMap 1FFFFFFFFF! (0fffffffff) Stay in ADV Chn"       # btle_rx.c:2181
    assert body[3].startswith("Pkt3 Ch37 AA:8e89bed6 ADV_PDU_t5:CONNECT_REQ")
    assert body[4:7] == ["Hop: track start ...", "Hop: next ch 7 freq 2418MHz access 50655a3b crcInit 1a2b3c", "Hop: next state 1"]
    assert body[7].startswith("Pkt4 Ch7 AA:50655a3b LL_PDU_t2:LL_DATA2") and "LL_Data:010203 CRC0" in body[7]
    assert body[8:10] == ["Hop: 1st data pdu", "Hop: next state 2"]
    assert body[10].startswith("Pkt5 Ch14 AA:50655a3b LL_PDU_t1:LL_DATA1")
    assert len(body) == 11                                                # nothing after the request from the advertising channels


def test_cli_hop_needs_wideband():
    rc, lines, eng = _run_cli(["-c", "37", "-o"])
    assert rc == 1 and eng is None and "needs --wideband" in lines[-1]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_connections_and_follow_on_a_wideband_capture(oracle_mod):
    from snout_b200.engine import RxEngine
    cap = synth.connection_capture(seconds=0.06, seed=7000)
    m = cap.meta
    with RxEngine("ble_wb40", max_samples=len(cap.iq), keep_streams=True) as e:
        adv = e.run(cap.iq)
        conns = e.connections()
        data = e.follow(m["access_addr"], m["crc_init"])
        again = e.follow(chanplan.BLE_ADV_AA, chanplan.BLE_ADV_CRC_INIT)
        q8 = e.debug_stage(_abi.STAGE_BLE_Q8)[0]
    # the advertising pass is unchanged by what follows it, and the second search with the advertising parameters repeats it
    assert_frames_equal(again, adv, what="follow(advertising AA) == the batch's own records")
    # both requests, in record order, with the reference's field values
    assert len(conns) == 2 and list(conns["chm_full"]) == [0, 1] and list(conns["channel"]) == [37, 37]
    c = conns[1]
    assert int(c["access_addr"]) == m["access_addr"] and int(c["crc_init"]) == m["crc_init"] and int(c["hop"]) == m["hop"]
    assert int(c["interval"]) == m["interval"] and bytes(c["init_a"]) == m["init_a"] and bytes(c["adv_a"]) == m["adv_a"]
    rec = adv[int(c["frame"])]
    assert int(rec["sample_index"]) == int(c["sample_index"]) and rec["bytes"][0] & 0x0F == 5 and rec["crc_ok"]
    # the connection's PDUs: bit exact against the oracle run with -a / -k of the connection on the engine's own streams
    want = np.concatenate([oracle_mod.ble_decode(q8[ch], ch, aa=m["access_addr"], crc_init=m["crc_init"]) for ch in range(40)])
    assert_frames_equal(data, want, what="follow vs oracle with the connection's AA / CRC init")
    if oracle_mod.have_ref("btle_ref"):
        ref = np.concatenate([oracle_mod.ble_decode(q8[ch], ch, aa=m["access_addr"], crc_init=m["crc_init"], impl="reference") for ch in range(40)])
        assert_frames_equal(data, ref, what="follow vs the unmodified btle_rx.c")
    # every transmitted LL PDU is there, CRC ok, on the hop sequence of the request
    sent = [(t.channel, bytes(t.data)) for t in cap.truth if t.channel < 37]
    data = data[np.argsort(data["sample_index"], kind="stable")]          # records come per channel; the air order is by time
    got = [(int(f["channel"]), bytes(f["bytes"][: f["len"]])) for f in data if f["crc_ok"]]
    assert len(sent) >= 10 and got == sent
    assert [ch for ch, _ in got][::2] == chanplan.ble_hop_channels(int(c["hop"]), len(got) // 2) == m["event_channels"]


@pytest.mark.gpu
def test_btle_rx_cli_follows_a_connection(tmp_path):
    cap = synth.connection_capture(seconds=0.06, seed=7001, hop=11)
    path = tmp_path / "conn.cf32"
    cap.iq.tofile(path)
    out = io.StringIO()
    o = btle_cli.parse_commandline(["-c", "37", "-o", "--wideband", "--iq", str(path), "--shard-windows", "8"])
    assert btle_cli.run(o, out=out) == 0
    lines = out.getvalue().splitlines()
    hop_lines = [l for l in lines if l.startswith("Hop:")]
    assert hop_lines[0].startswith("Hop: Not full ChnMap 1FFFFFFFFF! (1fff0fffff)")
    assert hop_lines[1:4] == ["Hop: track start ...", f"Hop: next ch 11 freq {chanplan.ble_channel_mhz(11)}MHz access 50655a3b crcInit 1a2b3c",
                              "Hop: next state 1"]
    ll = [l for l in lines if "LL_PDU_t" in l]
    sent = [t for t in cap.truth if t.channel < 37]
    assert len(ll) == len(sent) and all(f"Ch{t.channel} AA:50655a3b" in l and l.endswith("CRC0") for l, t in zip(ll, sent))
    first_ll = lines.index(ll[0])
    assert not any("ADV_PDU" in l for l in lines[first_ll:])             # the reference has left the advertising channel
