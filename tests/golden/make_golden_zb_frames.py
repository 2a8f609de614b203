#!/usr/bin/env python3
"""Generates tests/golden/zb_ref_frames.json from the 802.15.4 captures the reference ships for its own tests
(scapy-radio/scapy/test/pcaps/zigbee-transport-key-skke_1.pcap: DLT 195, FCS included;
 zigbee-join-authenticate.pcap: DLT 230, no FCS).  Run in the build container only."""
import json
import os
import struct

REF = "/root/reference/scapy-radio/scapy/test/pcaps"


def read_pcap(path):
    data = open(path, "rb").read()
    e = "<" if data[:4] == b"\xd4\xc3\xb2\xa1" else ">"
    linktype = struct.unpack_from(e + "I", data, 20)[0]
    off, out = 24, []
    while off + 16 <= len(data):
        _, _, caplen, _ = struct.unpack_from(e + "IIII", data, off)
        out.append(data[off + 16: off + 16 + caplen])
        off += 16 + caplen
    return linktype, out


def main():
    lt1, with_fcs = read_pcap(os.path.join(REF, "zigbee-transport-key-skke_1.pcap"))
    lt2, no_fcs = read_pcap(os.path.join(REF, "zigbee-join-authenticate.pcap"))
    assert lt1 == 195 and lt2 == 230
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump({"source": "scapy-radio/scapy/test/pcaps/zigbee-transport-key-skke_1.pcap (DLT 195), zigbee-join-authenticate.pcap (DLT 230)",
               "with_fcs": [f.hex() for f in with_fcs], "without_fcs": [f.hex() for f in no_fcs]},
              open(os.path.join(here, "zb_ref_frames.json"), "w"), indent=0)
    print(len(with_fcs), "frames with FCS,", len(no_fcs), "without")


if __name__ == "__main__":
    main()
