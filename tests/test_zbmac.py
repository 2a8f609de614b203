"""MAC-header summaries of the decoded 802.15.4 records.  The Python restatement is pinned against the
imported reference dissector (tests/golden/zbmac_ref.json, made by make_golden_zbmac.py from the vendored scapy); the
kernel's per-thread code equals the restatement on the host; on the GPU the summaries of received frames equal it too."""
import ctypes
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from snout_b200 import _abi, messages, synth

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zbmac_oracle as zo  # noqa: E402


def _golden():
    return json.load(open(os.path.join(GOLDEN, "zbmac_ref.json")))["frames"]


def test_oracle_equals_reference_dissector():
    rows = _golden()
    assert len(rows) >= 100 and sum(r["zll_scan_response"] for r in rows) >= 4
    n_short = n_raw = 0
    for r in rows:
        o = zo.parse(bytes.fromhex(r["hex"]))
        # a frame that ends inside a field its header announces (one capture of the reference does: a security header cut
        # short) is flagged; the reference dissector silently fills what is left -- the fields read before that point agree
        short = bool(o["present"] & zo.MALFORMED)
        n_short += short
        assert (o["frame_type"], o["seqnum"], o["dest_mode"], o["src_mode"]) == (r["frame_type"], r["seqnum"], r["dest_mode"], r["src_mode"])
        assert bool(o["present"] & zo.SECURITY) == bool(r["security"]) and bool(o["present"] & zo.ACKREQ) == bool(r["ackreq"])
        assert bool(o["present"] & zo.PENDING) == bool(r["pending"]) and bool(o["present"] & zo.PANID_COMPRESS) == bool(r["panid_compress"])
        for name, flag in (("dest_panid", zo.DEST_PANID), ("src_panid", zo.SRC_PANID)):
            assert (o[name] if o["present"] & flag else None) == r[name], (name, r["hex"])
        # scapy holds an address field of length 0 (mode 0 / reserved) as 0 or None; compare where a value was read
        for name, flag in (("dest_addr", zo.DEST_ADDR), ("src_addr", zo.SRC_ADDR)):
            if o["present"] & flag:
                assert o[name] == r[name], (name, r["hex"])
            else:
                assert not r[name], (name, r["hex"])
        if short:
            continue
        if o["present"] & zo.NO_ADDRESSING:                  # the reference kept the MAC payload raw: no field at all
            assert all(r[k] is None for k in ("dest_panid", "dest_addr", "src_panid", "src_addr", "cmd_id")) and not r["interpan"]
            n_raw += 1
            continue
        assert (o["cmd_id"] if o["frame_type"] == 3 else None) == r["cmd_id"]
        assert bool(o["present"] & zo.ZLL_SCAN_RESPONSE) == bool(r["zll_scan_response"])
        assert bool(o["present"] & zo.INTERPAN) == bool(r["interpan"])
        assert (o["zll_command"] if o["present"] & zo.ZLL else None) == r["zll_command"]
    assert n_short == 0 and n_raw >= 8


def _emu_parse(emu, psdu: bytes):
    out = np.zeros(1, _abi.ZBMAC_DTYPE)
    buf = (ctypes.c_uint8 * max(len(psdu), 1))(*psdu)
    emu.emu_zb_mac_parse(buf, len(psdu), out.ctypes.data_as(ctypes.c_void_p))
    return out[0]


def _same(row, o):
    for k in ("dest_addr", "src_addr", "dest_panid", "src_panid", "fcf", "present", "seqnum", "frame_type", "dest_mode", "src_mode",
              "cmd_id", "payload_off", "zll_command", "cluster", "profile"):
        assert int(row[k]) == int(o[k]), (k, int(row[k]), o[k])


def test_kernel_code_equals_oracle_on_host(emu):
    rng = np.random.default_rng(4)
    frames = [bytes.fromhex(r["hex"]) for r in _golden()]
    for f in list(frames):                                   # every truncation of every golden frame + random bytes
        for cut in range(0, len(f), 3):
            frames.append(f[:cut])
    frames += [bytes(rng.integers(0, 256, int(rng.integers(0, 128)), dtype=np.uint8)) for _ in range(3000)]
    n_mal = 0
    for f in frames:
        o = zo.parse(f)
        _same(_emu_parse(emu, f), o)
        n_mal += bool(o["present"] & zo.MALFORMED)
    assert n_mal > 500 and len(frames) - n_mal > 500


def test_message_builder_keys():
    fr = np.zeros(2, _abi.FRAME_DTYPE)
    fr["proto"], fr["channel"], fr["lqi"], fr["crc_ok"], fr["len"] = 2, 15, [255, 128], 1, 20
    mac = np.zeros(2, _abi.ZBMAC_DTYPE)
    mac["present"] = [_abi.ZBMAC_SRC_ADDR | _abi.ZBMAC_DEST_ADDR | _abi.ZBMAC_DEST_PANID | _abi.ZBMAC_ZLL_SCAN_RESPONSE, _abi.ZBMAC_DEST_PANID]
    mac["src_addr"], mac["dest_addr"], mac["src_mode"], mac["dest_mode"], mac["seqnum"], mac["frame"] = [0x1122, 0], [0xFFFF, 0], [2, 0], [2, 0], [7, 8], [0, 1]
    ms = messages.zigbee_messages(fr, mac, timestamp=1.5)
    assert set(ms[0]) >= {"sender", "receiver", "seq_number", "timestamp", "pan", "rftap", "vuln"}      # message.py:286-303
    assert ms[0]["sender"] == 0x1122 and ms[0]["receiver"] == 0xFFFF and ms[0]["vuln"] == {"zll": True} and ms[0]["rftap"]["qual"] == 1.0
    assert ms[1]["sender"] is None and ms[1]["pan"] == {"src_panid": None, "dest_panid": 0}
    assert len(messages.zll_scan_responses(fr, mac)) == 1


@pytest.mark.gpu
def test_gpu_mac_summary_of_received_frames():
    """The golden frames (reference pcaps + scapy-built frames, FCS valid) transmitted, received by the engine and
    summarised on the GPU: every summary equals the restatement of the received PSDU; the ZLL scan responses are found."""
    from snout_b200.engine import RxEngine
    psdus = [bytes.fromhex(r["hex"]) for r in _golden() if 5 <= len(bytes.fromhex(r["hex"])) <= 127]
    rng = np.random.default_rng(8)
    sig, truth = synth.zb_baseband(2_400_000, 20, rng, gap=(1500, 5000), psdus=psdus)
    x = (sig + synth._awgn(len(sig), rng, 2.0 / 10 ** 2.5)).astype(np.complex64)
    with RxEngine("zb_nb", channel=20, max_samples=len(x)) as e:
        fr = e.run(x)
        mac = e.zb_mac_summary()
    assert len(fr) == len(mac) >= len(psdus) and fr["crc_ok"].all()
    assert [bytes(f["bytes"][: f["len"]]) for f in fr[: len(psdus)]] == psdus
    for f, m in zip(fr, mac):
        _same(m, zo.parse(bytes(f["bytes"][: f["len"]])))
    assert np.array_equal(mac["frame"], np.arange(len(fr)))
    want = sum(r["zll_scan_response"] for r in _golden() if 5 <= len(bytes.fromhex(r["hex"])) <= 127)
    got = messages.zll_scan_responses(fr[: len(psdus)], mac[: len(psdus)])
    assert len(got) == want >= 4 and all(g["channel"] == 20 for g in got)
    ms = messages.zigbee_messages(fr, mac)
    assert len(ms) == len(fr) and sum(m["vuln"]["zll"] for m in ms[: len(psdus)]) == want
