"""Output formats against the reference's own outputs.

* btle_rx stdout line: golden text printed by the UNMODIFIED reference receiver()
  (tests/golden/btle_synth_ref.npz stdout_*, btle_welcome.npz stdout; made by make_golden.py);
* RFtap datagram: the reference's test/rftap.pcap and datagrams that passed the vendored scapy
  dissector (formats_ref.json, made by make_golden_formats.py);
* GnuradioPacket / wrpcap DLT 195: bytes built by the vendored scapy;
* btle_rx -s pcap: layout of write_packet_to_file(), btle_rx.c:126-163.
"""
import io
import json
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN
from snout_b200 import _abi, formats


def _frame(psdu: bytes, proto=2, lqi=0, channel=11, aa=0, crc_ok=1, sample_index=0):
    f = np.zeros(1, _abi.FRAME_DTYPE)[0]
    f["bytes"][: len(psdu)] = np.frombuffer(psdu, np.uint8)
    f["len"], f["lqi"], f["proto"], f["channel"], f["access_addr"], f["crc_ok"] = len(psdu), lqi, proto, channel, aa, crc_ok
    f["sample_index"] = sample_index
    return f


@pytest.fixture(scope="module")
def ref():
    return json.load(open(os.path.join(GOLDEN, "formats_ref.json")))


@pytest.mark.parametrize("seed", [1001, 1002, 1003, 1004])
def test_btle_rx_line_equals_reference_stdout(golden, seed):
    g = golden("btle_synth_ref.npz")
    want = str(g[f"stdout_{seed}"]).splitlines(keepends=True)
    got = formats.btle_rx_lines(g[f"frames_{seed}"], first_pkt=1, timestamp=(1, 2))
    assert len(got) == len(want) > 40
    for a, b in zip(got, want):
        assert formats.strip_timestamp(a) == formats.strip_timestamp(b)
    assert got[0].startswith("1.000002 Pkt1 Ch")


def test_btle_rx_line_welcome_and_snout_parser(golden):
    g = golden("btle_welcome.npz")
    line = formats.btle_rx_line(g["frames"][0], 1, timestamp=1567108496.651985)
    assert formats.strip_timestamp(line) == formats.strip_timestamp(str(g["stdout"]))
    # what BtleMessage.fromraw does with it (snout/core/message.py:226-235)
    tok = line.encode().decode().split(" ")
    assert len(tok) == 11 and tok[-1] == "CRC0\n"
    assert tok[0] == "1567108496.651985" and tok[1][3:] == "1" and tok[2][2:] == "37"
    assert tok[4][11:] == "ADV_NONCONN_IND" and tok[8][5:] == "010203040506"
    assert bytes.fromhex(tok[9][5:]).endswith(b"welcome u!")


def test_btle_rx_line_pdu_types():
    def ble(hdr0, payload, ch=37, crc_ok=1):
        b = bytes([hdr0, len(payload)]) + payload + b"\x00\x00\x00"
        return _frame(b, proto=3, channel=ch, aa=0x8E89BED6, crc_ok=crc_ok)
    t = (0, 0)
    a, i = bytes(range(1, 7)), bytes(range(0x11, 0x17))
    # SCAN_REQ: A0 / A1 both printed MSB first (btle_rx.c:1979-1988)
    assert formats.strip_timestamp(formats.btle_rx_line(ble(0x43, a + i), 7, t)) == \
        "Pkt7 Ch37 AA:8e89bed6 ADV_PDU_t3:SCAN_REQ T1 R0 PloadL12 A0:060504030201 A1:161514131211 CRC0\n"
    # wrong length for type 1: the Error text follows the prefix, no CRC token (1453-1455, 2142-2144)
    assert formats.strip_timestamp(formats.btle_rx_line(ble(0x01, a + i + b"\x00"), 1, t)) == \
        "Pkt1 Ch37 AA:8e89bed6 ADV_PDU_t1:ADV_DIRECT_IND T0 R0 PloadL13 Error: Payload length 13 bytes. Need to be 12 for PDU Type ADV_DIRECT_IND!\n"
    # CONNECT_REQ field order and byte orders (1476-1557, 1989-2008)
    ll = bytes.fromhex("d6be898e") + bytes.fromhex("123456") + bytes([3]) + struct.pack("<HHHH", 5, 0x18, 0, 0x48) + \
        bytes.fromhex("ffffffff1f") + bytes([0x2A])
    got = formats.strip_timestamp(formats.btle_rx_line(ble(0x85, a + i + ll, ch=39, crc_ok=0), 2, t))
    assert got == ("Pkt2 Ch39 AA:8e89bed6 ADV_PDU_t5:CONNECT_REQ T0 R1 PloadL34 InitA:060504030201 AdvA:161514131211 AA:8e89bed6 "
                   "CRCInit:123456 WSize:03 WOffset:0005 Itrvl:0018 Ltncy:0000 Timot:0048 ChM:1fffffffff Hop:10 SCA:1 CRC1\n")
    # reserved advertising type: raw bytes
    assert formats.strip_timestamp(formats.btle_rx_line(ble(0x0A, a), 3, t)).endswith("ADV_PDU_t10:RESERVED3 T0 R0 PloadL6 Byte:010203040506 CRC0\n")
    # data channel: empty PDU, data PDU, control PDU (1850-1958)
    assert formats.strip_timestamp(formats.btle_rx_line(ble(0x01, b"", ch=5), 1, t)) == \
        "Pkt1 Ch5 AA:8e89bed6 LL_PDU_t1:LL_DATA1 NESN0 SN0 MD0 PloadL0 CRC0\n"
    assert formats.strip_timestamp(formats.btle_rx_line(ble(0x1E, b"\xde\xad", ch=5), 1, t)) == \
        "Pkt1 Ch5 AA:8e89bed6 LL_PDU_t2:LL_DATA2 NESN1 SN1 MD1 PloadL2 LL_Data:dead CRC0\n"
    ver = bytes([0x0C, 0x09, 0x0F, 0x00, 0x34, 0x12])
    assert formats.strip_timestamp(formats.btle_rx_line(ble(0x03, ver, ch=5), 1, t)) == \
        "Pkt1 Ch5 AA:8e89bed6 LL_PDU_t3:LL_CTRL NESN0 SN0 MD0 PloadL6 Op0c(LL_VERSION_IND) Ver:09 CompId:000f SubVer:1234 CRC0\n"
    assert "Op63(LL_RESERVED) Byte:0102 CRC0" in formats.btle_rx_line(ble(0x03, bytes([0x63, 1, 2]), ch=5), 1, t)


def test_ble_pcap_record_layout(golden):
    g = golden("btle_sample_iq_4msps.npz")
    buf = io.BytesIO()
    assert formats.write_ble_pcap(buf, g["frames"]) == 3
    data = buf.getvalue()
    assert data[:24] == bytes.fromhex("a1b2c3d4000200040000000000000000000005dc00000100")      # btle_rx.c:129, linktype 256
    off = 24
    for f in g["frames"]:
        n = int(f["len"]) - 3                                   # header + payload, CRC not stored (btle_rx.c:2134)
        caplen, plen = struct.unpack_from(">ii", data, off + 8)  # htonl()
        assert caplen == plen == 10 + 4 + n
        assert data[off + 16: off + 26] == bytes([37, 0, 0, 0, 0, 0, 0, 0, 1, 0])
        assert data[off + 26: off + 30] == bytes.fromhex("d6be898e")
        assert data[off + 30: off + 30 + n] == bytes(f["bytes"][:n])
        off += 16 + caplen
    assert off == len(data)


def test_rftap_datagram_equals_reference_fixture(ref):
    lt, recs = formats.read_pcap(open(os.path.join(GOLDEN, "rftap.pcap"), "rb").read())
    assert len(recs) == 10 == len(ref["rftap_pcap"])
    for (_, raw), r in zip(recs, ref["rftap_pcap"]):
        assert raw.hex() == r["datagram_hex"]
        f = _frame(bytes.fromhex(r["psdu_hex"]), lqi=int(round(r["qual"] * 255)))
        assert formats.rftap_datagram(f) == raw
        p = formats.parse_rftap(raw)
        assert p["dlt"] == 195 and p["qual"] == np.float32(r["qual"]) and p["payload"].hex() == r["psdu_hex"]
    assert recs[0][1][:16] == bytes.fromhex("5246746104000101c30000000000803f")
    for b in ref["rftap_built"]:                                 # datagrams accepted by the reference dissector
        f = _frame(bytes.fromhex(b["psdu_hex"]), lqi=b["lqi"])
        assert formats.rftap_datagram(f).hex() == b["datagram_hex"]


def test_gnuradio_packet_equals_scapy_build(ref):
    for r in ref["gnuradio_zigbee"]:
        assert formats.gnuradio_packet(_frame(bytes.fromhex(r["psdu_hex"]))).hex() == r["packet_hex"]
    for r in ref["gnuradio_ble"]:
        f = _frame(bytes.fromhex(r["pdu_crc_hex"]), proto=3, channel=37, aa=r["access_addr"])
        assert formats.gnuradio_packet(f).hex() == r["packet_hex"]
        assert r["scapy_crc_hex"] == r["pdu_crc_hex"][-6:]


def test_zigbee_pcap_equals_scapy_wrpcap(ref):
    w = ref["wrpcap195"]
    buf = io.BytesIO()
    buf.write(formats.pcap_global_header(195))
    for p, t in zip(w["psdus_hex"], w["times"]):
        buf.write(formats.pcap_record(bytes.fromhex(p), t))
    assert buf.getvalue().hex() == w["file_hex"]
    # the record-driven writer: time stamp = base + sample_index / 4 Msps
    frames = [_frame(bytes.fromhex(p), sample_index=1_000_000 * i) for i, p in enumerate(w["psdus_hex"])]
    buf2 = io.BytesIO()
    assert formats.write_zigbee_pcap(buf2, frames, ts=1000.0) == 4
    assert buf2.getvalue().hex() == w["file_hex"]
    lt, recs = formats.read_pcap(buf2.getvalue())
    assert lt == 195 and [r[1].hex() for r in recs] == w["psdus_hex"]


def test_fcs_known_answers_from_scapy(ref, oracle_mod):
    for r in ref["fcs"]:
        want = int.from_bytes(bytes.fromhex(r["fcs_le_hex"]), "little")
        assert oracle_mod.zb_fcs16(bytes.fromhex(r["frame_hex"])) == want


def test_vectorised_block_writers_equal_the_record_writers():
    """the numpy batch writers produce exactly the bytes of the per-record writers (which the tests above pin
    to the reference's own files), for ragged lengths, empty batches and mixed-protocol lists."""
    import time
    from snout_b200 import _abi
    rng = np.random.default_rng(3)
    n = 5000
    f = np.zeros(n, _abi.FRAME_DTYPE)
    f["proto"] = rng.choice([2, 3], n)
    f["channel"] = np.where(f["proto"] == 3, rng.integers(0, 40, n), rng.integers(11, 27, n))
    f["len"] = np.where(f["proto"] == 3, rng.integers(5, 45, n), rng.integers(2, 128, n))
    f["bytes"] = rng.integers(0, 256, (n, 132), dtype=np.uint8)
    f["access_addr"] = 0x8E89BED6
    f["lqi"] = rng.integers(0, 256, n)
    f["sample_index"] = np.sort(rng.integers(-4, 40_000_000, n))
    t0 = time.perf_counter()
    ble = formats.ble_pcap_block(f, ts=(7, 9))
    zb = formats.zigbee_pcap_block(f, ts=1000.0)
    rf, off = formats.rftap_block(f)
    t_vec = time.perf_counter() - t0
    t0 = time.perf_counter()
    ble_ref = b"".join(formats.ble_pcap_record(x, (7, 9)) for x in f if x["proto"] == 3)
    zb_ref = b"".join(formats.pcap_record(bytes(x["bytes"][: int(x["len"])]), 1000.0 + max(0, int(x["sample_index"])) / 4e6) for x in f if x["proto"] == 2)
    rf_ref = [formats.rftap_datagram(x) for x in f if x["proto"] == 2]
    t_loop = time.perf_counter() - t0
    assert ble == ble_ref and zb == zb_ref and rf == b"".join(rf_ref)
    assert [rf[a:b] for a, b in zip(off[:-1], off[1:])] == rf_ref
    assert formats.ble_pcap_block(f[:0]) == b"" and formats.zigbee_pcap_block(f[f["proto"] == 3]) == b""
    assert t_vec < t_loop                                         # and it is the faster way (typically 20-50x)
