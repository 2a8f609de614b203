#!/usr/bin/env python3
"""Generates tests/golden/btle_connreq_ref.json: CONNECT_REQ payloads through the UNMODIFIED reference's own field
extraction (parse_adv_pdu_payload_byte, vendor/BTLE/host/btle-tools/src/btle_rx.c:1476-1557) and what receiver_controller
would start tracking (receiver_status, chm_is_full_map :2158-2163), via oracle/_ref/libbtle_ref.so (built by
`make -C oracle ref` from the sources under /root/reference).  Run in the build container only; the GPU box and the tests
read the committed JSON.  Every third payload carries the full channel map (the only kind btle_rx -o follows)."""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def main():
    lib = ctypes.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libbtle_ref.so"))
    fn = lib.btle_ref_parse_connect_req
    fn.restype = ctypes.c_int
    rng = np.random.default_rng(20261018)
    rows = []
    for i in range(24):
        p = bytearray(rng.integers(0, 256, 34, dtype=np.uint8).tobytes())
        if i % 3 == 0:
            p[28:33] = b"\xff\xff\xff\xff\x1f"
        out = (ctypes.c_uint32 * 12)()
        ia, aa, chm = (ctypes.c_uint8 * 6)(), (ctypes.c_uint8 * 6)(), (ctypes.c_uint8 * 5)()
        rc = fn(bytes(p), 34, out, ia, aa, chm)
        assert rc == 0
        rows.append(dict(payload=bytes(p).hex(), access_addr=out[0], crc_init=out[1], win_size=out[2], win_offset=out[3],
                         interval=out[4], latency=out[5], timeout=out[6], hop=out[7], sca=out[8], chm_full=out[9],
                         status_hop=out[10], status_interval=out[11], init_a_reversed=bytes(ia).hex(),
                         adv_a_reversed=bytes(aa).hex(), chm_reversed=bytes(chm).hex()))
    # wrong payload length: the reference prints an error and returns -1 (btle_rx.c:1477-1480)
    out = (ctypes.c_uint32 * 12)()
    bad = fn(bytes(33), 33, out, (ctypes.c_uint8 * 6)(), (ctypes.c_uint8 * 6)(), (ctypes.c_uint8 * 5)())
    json.dump(dict(source="parse_adv_pdu_payload_byte of the unmodified btle_rx.c (oracle/_ref), CONNECT_REQ", rows=rows,
                   rc_for_33_byte_payload=bad),
              open(os.path.join(HERE, "btle_connreq_ref.json"), "w"), indent=0)
    print(len(rows), "rows; rc for a 33-byte payload:", bad)


if __name__ == "__main__":
    main()
