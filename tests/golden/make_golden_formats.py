#!/usr/bin/env python3
"""Regenerate tests/golden/formats_ref.json from the reference's own Python code.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden_formats.py

The vendored scapy fork (scapy-radio/scapy) is imported with the `six` shim of SURVEY 8c and asked
to BUILD the wire formats our snout_b200.formats module must reproduce:

  rftap        raw(RFtap(flags="dlt+qual", length32=4, dlt=195, qual=q) / Dot15d4FCS(psdu))
               for every PSDU of the reference's test/rftap.pcap and a few lqi values
  gnuradio     raw(GnuradioPacket(proto=2) / Dot15d4FCS(psdu)),  raw(GnuradioPacket(proto=3) / BTLE(...))
  wrpcap195    the bytes scapy's wrpcap() writes for [Dot15d4FCS(psdu), ...] with fixed time stamps
  fcs          Dot15d4FCS(psdu[:-2]) recomputed FCS for every PSDU (FCS-16 known answers)
  crc24        BTLE.compute_crc for the golden BLE frames
"""
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, f"{REF}/scapy-radio/scapy")

import six  # noqa: E402
import scapy  # noqa: E402
import scapy.modules  # noqa: E402

sys.modules["scapy.modules.six"] = six
sys.modules["scapy.modules.six.moves"] = six.moves
scapy.modules.six = six

from scapy.layers.bluetooth4LE import BTLE  # noqa: E402
from scapy.layers.dot15d4 import Dot15d4FCS  # noqa: E402
from scapy.layers.gnuradio import GnuradioPacket  # noqa: E402
from scapy.layers.rftap import RFtap  # noqa: E402
from scapy.packet import Raw  # noqa: E402
from scapy.utils import rdpcap, wrpcap  # noqa: E402

import numpy as np  # noqa: E402


def main():
    out = {"source": "vendored scapy-radio/scapy @ reference checkout; built by tests/golden/make_golden_formats.py"}
    pk = rdpcap(f"{REF}/scapy-radio/scapy/test/rftap.pcap")
    psdus = []
    rft = []
    for p in pk:
        raw = bytes(p)
        r = RFtap(raw)
        # (the fork's LE_XBitField reads length32 big-endian: 0x0400 -- a dissector quirk, the wire bytes are 04 00)
        assert r.dlt == 195 and r.magic == 0x52467461 and r.flags.dlt and r.flags.qual
        psdu = raw[16:]
        psdus.append(psdu)
        rft.append({"datagram_hex": raw.hex(), "qual": float(r.qual), "psdu_hex": psdu.hex()})
    # the fork's RFtap layer cannot BUILD (LE_XBitField.addfield raises), so for other LQI values the
    # datagram built by snout_b200.formats is pushed through the reference DISSECTOR and pinned
    # once it reads back dlt 195, qual = lqi / 255.0 (epy_block_0.py:21) and the PSDU as Dot15d4FCS
    from snout_b200 import _abi, formats
    built = []
    for lqi in (0, 8, 96, 200, 248, 255):
        psdu = psdus[lqi % len(psdus)]
        f = np.zeros(1, _abi.FRAME_DTYPE)[0]
        f["bytes"][: len(psdu)] = np.frombuffer(psdu, np.uint8)
        f["len"], f["lqi"], f["proto"] = len(psdu), lqi, 2
        dg = formats.rftap_datagram(f)
        r = RFtap(dg)
        q = struct.unpack("<f", struct.pack("<f", lqi / 255.0))[0]
        assert r.dlt == 195 and r.qual == q and isinstance(r.payload, Dot15d4FCS) and bytes(r.payload) == psdu
        built.append({"lqi": lqi, "psdu_hex": psdu.hex(), "datagram_hex": dg.hex(), "qual": q})
    out["rftap_pcap"] = rft
    out["rftap_built"] = built

    out["gnuradio_zigbee"] = [{"psdu_hex": p.hex(), "packet_hex": bytes(GnuradioPacket(proto=2) / Raw(p)).hex()} for p in psdus[:3]]
    g = np.load(f"{HERE}/btle_sample_iq_4msps.npz")
    ble = []
    for f in g["frames"]:
        b = bytes(f["bytes"][: int(f["len"])])
        body = struct.pack("<I", int(f["access_addr"])) + b
        pkt = GnuradioPacket(proto=3) / Raw(body)
        d = GnuradioPacket(bytes(pkt))
        assert isinstance(d.payload, BTLE) and d.payload.access_addr == 0x8E89BED6
        crc = BTLE.compute_crc(b[:-3])
        assert crc == b[-3:], (crc.hex(), b[-3:].hex())
        ble.append({"pdu_crc_hex": b.hex(), "access_addr": int(f["access_addr"]), "packet_hex": bytes(pkt).hex(),
                    "scapy_crc_hex": crc.hex()})
    out["gnuradio_ble"] = ble

    # wrpcap of Dot15d4FCS packets with fixed time stamps
    pkts = []
    for i, p in enumerate(psdus[:4]):
        d = Dot15d4FCS(p)
        d.time = 1000.0 + 0.25 * i
        pkts.append(d)
    tmp = "/tmp/_snout_b200_wrpcap195.pcap"
    wrpcap(tmp, pkts)
    out["wrpcap195"] = {"psdus_hex": [p.hex() for p in psdus[:4]], "times": [1000.0 + 0.25 * i for i in range(4)],
                        "file_hex": open(tmp, "rb").read().hex()}
    os.remove(tmp)

    fcs = []
    for p in psdus:
        d = Dot15d4FCS(p)
        want = d.compute_fcs(p[:-2])
        fcs.append({"frame_hex": p[:-2].hex(), "fcs_le_hex": bytes(want).hex()})
    out["fcs"] = fcs
    json.dump(out, open(f"{HERE}/formats_ref.json", "w"), indent=1)
    print("wrote formats_ref.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
