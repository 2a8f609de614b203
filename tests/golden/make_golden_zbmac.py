#!/usr/bin/env python3
"""Generates tests/golden/zbmac_ref.json by importing the REFERENCE dissector (the vendored scapy of
/root/reference/scapy-radio, conf.dot15d4_protocol = 'zigbee' as snout/cli.py:24 sets it) and reading off, for a set of
802.15.4 frames, the fields Snout consumes (snout/core/message.py:258-304, snout/util/zigbee.py:176-202): frame type,
sequence number, PAN ids, addresses, MAC command id, and whether the tree holds a ZLLScanResponse.  Run in the build
container only; the GPU box and the tests read the committed JSON.

Frames: the 55 frames of the reference's own test captures (tests/golden/zb_ref_frames.json) and frames BUILT with the
reference's scapy classes covering every addressing mode, PAN-id compression, security headers, MAC commands, beacons,
acks and inter-PAN ZLL scan requests / responses."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/scapy-radio/scapy")
import six  # noqa: E402
import scapy  # noqa: E402
import scapy.modules  # noqa: E402
sys.modules['scapy.modules.six'] = six
sys.modules['scapy.modules.six.moves'] = six.moves
scapy.modules.six = six
from scapy.config import conf  # noqa: E402
conf.dot15d4_protocol = 'zigbee'
from scapy.layers.dot15d4 import (Dot15d4, Dot15d4FCS, Dot15d4Data, Dot15d4Cmd, Dot15d4Beacon, Dot15d4Ack,  # noqa: E402
                                  Dot15d4AuxSecurityHeader)
from scapy.layers.zigbee import (ZigbeeNWKStub, ZigbeeAppDataPayloadStub, ZigbeeZLLCommissioningCluster,  # noqa: E402
                                 ZLLScanRequest, ZLLScanResponse)
from scapy.packet import Raw  # noqa: E402


def fields(raw: bytes) -> dict:
    p = Dot15d4FCS(raw)
    o = {"hex": raw.hex(), "frame_type": int(p.fcf_frametype), "seqnum": int(p.seqnum), "dest_mode": int(p.fcf_destaddrmode),
         "src_mode": int(p.fcf_srcaddrmode), "security": int(bool(p.fcf_security)), "ackreq": int(bool(p.fcf_ackreq)),
         "pending": int(bool(p.fcf_pending)), "panid_compress": int(bool(p.fcf_panidcompress))}
    for name in ("dest_panid", "dest_addr", "src_panid", "src_addr", "cmd_id"):
        v = None
        for layer in (Dot15d4Data, Dot15d4Cmd, Dot15d4Beacon):
            if p.haslayer(layer) and name in p[layer].fields:
                v = p[layer].fields[name]
        o[name] = None if v is None else int(v)
    o["zll_scan_response"] = int(p.haslayer(ZLLScanResponse))
    o["zll_command"] = int(p[ZigbeeZLLCommissioningCluster].command_identifier) if p.haslayer(ZigbeeZLLCommissioningCluster) else None
    o["interpan"] = int(p.haslayer(ZigbeeNWKStub))
    return o


def built(rng):
    out = []
    modes = [(2, 2), (2, 3), (3, 2), (3, 3), (2, 0), (0, 2), (0, 3), (3, 0)]
    for dm, sm in modes:
        for comp in (0, 1):
            for sec in (0, 1):
                kw = dict(fcf_frametype=1, fcf_destaddrmode=dm, fcf_srcaddrmode=sm, fcf_panidcompress=comp, fcf_security=sec,
                          fcf_ackreq=int(rng.integers(0, 2)), seqnum=int(rng.integers(0, 256)))
                d = dict(dest_panid=int(rng.integers(0, 65536)), dest_addr=int(rng.integers(0, 1 << (16 if dm == 2 else 63))) if dm else 0)
                if sm:
                    d["src_addr"] = int(rng.integers(0, 1 << (16 if sm == 2 else 63)))
                    if not comp:
                        d["src_panid"] = int(rng.integers(0, 65536))
                if sec:
                    d["aux_sec_header"] = Dot15d4AuxSecurityHeader(sec_sc_keyidmode=int(rng.integers(0, 4)), sec_sc_seclevel=5,
                                                                  sec_framecounter=int(rng.integers(0, 1 << 32)))
                pkt = Dot15d4FCS(**kw) / Dot15d4Data(**d) / Raw(bytes(rng.integers(0, 256, int(rng.integers(1, 20)), dtype=np.uint8) & 0xFC))
                out.append(bytes(pkt))
    for cmd in (1, 2, 3, 4, 6, 7, 8):
        pkt = Dot15d4FCS(fcf_frametype=3, fcf_destaddrmode=2, fcf_srcaddrmode=3, seqnum=cmd) / \
            Dot15d4Cmd(dest_panid=0xFFFF, dest_addr=0xFFFF, src_panid=0x1234, src_addr=0x0011223344556677, cmd_id=cmd)
        out.append(bytes(pkt))
    out.append(bytes(Dot15d4FCS(fcf_frametype=2, fcf_destaddrmode=0, seqnum=77) / Dot15d4Ack()))
    out.append(bytes(Dot15d4FCS(fcf_frametype=0, fcf_destaddrmode=0, fcf_srcaddrmode=2, seqnum=5) /
                     Dot15d4Beacon(src_panid=0xBEEF, src_addr=0x0001)))
    out.append(bytes(Dot15d4FCS(fcf_frametype=0, fcf_destaddrmode=0, fcf_srcaddrmode=3, seqnum=6) /
                     Dot15d4Beacon(src_panid=0xBEEF, src_addr=0x0102030405060708)))
    # inter-PAN ZLL commissioning: scan request (broadcast) and scan response (what zigbee.py:176-192 looks for)
    for cmd, body in ((0x00, ZLLScanRequest()), (0x01, ZLLScanResponse())):
        for man in (0, 1):
            for delivery in (0, 2, 3):
                pkt = Dot15d4FCS(fcf_frametype=1, fcf_destaddrmode=2, fcf_srcaddrmode=3, fcf_panidcompress=0, seqnum=int(rng.integers(0, 256))) / \
                    Dot15d4Data(dest_panid=0xFFFF, dest_addr=0xFFFF, src_panid=0x4242, src_addr=0x00178801020304AA) / \
                    ZigbeeNWKStub() / ZigbeeAppDataPayloadStub(delivery_mode=delivery, cluster=0x1000, profile=0xC05E) / \
                    ZigbeeZLLCommissioningCluster(manufacturer_specific=man, command_identifier=cmd, transaction_sequence=9) / body
                out.append(bytes(pkt))
    # inter-PAN but another profile / cluster
    pkt = Dot15d4FCS(fcf_frametype=1, fcf_destaddrmode=2, fcf_srcaddrmode=2, fcf_panidcompress=1, seqnum=3) / \
        Dot15d4Data(dest_panid=0x1111, dest_addr=0x2222, src_addr=0x3333) / ZigbeeNWKStub() / \
        ZigbeeAppDataPayloadStub(cluster=0x0006, profile=0x0104) / Raw(b"\x01\x02\x03")
    out.append(bytes(pkt))
    return out


def main():
    rng = np.random.default_rng(20260)
    g = json.load(open(os.path.join(HERE, "zb_ref_frames.json")))
    frames = [bytes.fromhex(h) for h in g["with_fcs"]]
    for h in g["without_fcs"]:
        b = bytes.fromhex(h)
        frames.append(bytes(Dot15d4FCS(b + b"\x00\x00"))[:len(b)] + Dot15d4FCS().compute_fcs(b))
    frames += built(rng)
    rows = [fields(f) for f in frames]
    json.dump({"source": "vendored scapy dissector (Dot15d4FCS ...), conf.dot15d4_protocol='zigbee'", "frames": rows},
              open(os.path.join(HERE, "zbmac_ref.json"), "w"), indent=0)
    print(len(rows), "frames ->", os.path.join(HERE, "zbmac_ref.json"),
          "| scan responses:", sum(r["zll_scan_response"] for r in rows), "| inter-PAN:", sum(r["interpan"] for r in rows))


if __name__ == "__main__":
    main()
