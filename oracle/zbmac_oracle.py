"""TEST INFRASTRUCTURE ONLY -- Python restatement of the 802.15.4 MAC-header walk of SURVEY 8(f) N2.

Follows the reference's dissector field by field: scapy-radio/scapy/scapy/layers/dot15d4.py (Dot15d4 :87-100,
Dot15d4FCS :140-172, Dot15d4Data :216-231, Dot15d4Beacon :255-287, Dot15d4Cmd :294-331, util_srcpanid_present :359-364,
Dot15d4AuxSecurityHeader :184-209) and, for the inter-PAN path with conf.dot15d4_protocol = 'zigbee' (snout/cli.py:24),
scapy/layers/zigbee.py (ZigbeeNWKStub :717-731, ZigbeeAppDataPayloadStub :734-763, ZigbeeZLLCommissioningCluster :1027-1059).
Pinned against the imported reference dissector by tests/golden/make_golden_zbmac.py -> tests/golden/zbmac_ref.json."""
from __future__ import annotations

SECURITY, PENDING, ACKREQ, PANID_COMPRESS, DEST_PANID, DEST_ADDR, SRC_PANID, SRC_ADDR, INTERPAN, ZLL, ZLL_SCAN_RESPONSE, NO_ADDRESSING = (1 << i for i in range(12))
MALFORMED = 0x4000


def _le(b: bytes) -> int:
    return int.from_bytes(b, "little")


def parse(psdu: bytes) -> dict:
    """PSDU (MHR | payload | FCS) -> the fields of snrx_zbmac_t."""
    o = dict(dest_addr=0, src_addr=0, dest_panid=0, src_panid=0, fcf=0, present=0, seqnum=0, frame_type=0xFF, dest_mode=0,
             src_mode=0, cmd_id=0xFF, payload_off=0, zll_command=0xFF, cluster=0, profile=0)
    if len(psdu) < 5:
        o["present"] |= MALFORMED
        return o
    d = psdu[:-2]
    b0, b1 = d[0], d[1]
    o["fcf"] = b0 | (b1 << 8)
    o["frame_type"] = b0 & 7
    security, compress = (b0 >> 3) & 1, (b0 >> 6) & 1
    o["present"] |= (SECURITY if security else 0) | (PENDING if (b0 >> 4) & 1 else 0) | (ACKREQ if (b0 >> 5) & 1 else 0) | \
                    (PANID_COMPRESS if compress else 0)
    o["dest_mode"], o["src_mode"], o["seqnum"] = (b1 >> 2) & 3, (b1 >> 6) & 3, d[2]
    pos = 3
    alen = {2: 2, 3: 8}

    class Short(Exception):
        pass

    def take(k):
        nonlocal pos
        if pos + k > len(d):
            raise Short()
        v = d[pos:pos + k]
        pos += k
        return v

    ft = o["frame_type"]
    # dot15d4AddressField.getfield raises for a mode without a length (dot15d4.py:60-63): the layer is kept as raw bytes
    if (ft in (1, 3) and (o["dest_mode"] < 2 or o["src_mode"] == 1)) or (ft == 0 and o["src_mode"] < 2):
        o["present"] |= NO_ADDRESSING
        o["payload_off"] = pos
        return o
    try:
        if ft in (1, 3):
            o["dest_panid"] = _le(take(2)); o["present"] |= DEST_PANID
            if o["dest_mode"] in alen:
                o["dest_addr"] = _le(take(alen[o["dest_mode"]])); o["present"] |= DEST_ADDR
            if o["src_mode"] != 0 and not compress:
                o["src_panid"] = _le(take(2)); o["present"] |= SRC_PANID
            if o["src_mode"] in alen:
                o["src_addr"] = _le(take(alen[o["src_mode"]])); o["present"] |= SRC_ADDR
        elif ft == 0:
            o["src_panid"] = _le(take(2)); o["present"] |= SRC_PANID
            if o["src_mode"] in alen:
                o["src_addr"] = _le(take(alen[o["src_mode"]])); o["present"] |= SRC_ADDR
        else:
            o["payload_off"] = pos
            return o
        # the auxiliary security header is not skipped: the reference tests `fcf_security is True` on the int 1
        # (dot15d4.py:227-228, :265-266, :305-306), so its dissector never parses that header
        if ft == 3:
            o["cmd_id"] = take(1)[0]
        o["payload_off"] = pos
        if ft != 1:
            return o
        p = d[pos:]
        if len(p) < 1 or (p[0] & 3) != 3:
            return o
        o["present"] |= INTERPAN
        pos += 2
        aps = take(1)[0]
        if ((aps >> 2) & 3) == 3:
            pos += 2
        cp = take(4)
        o["cluster"], o["profile"] = _le(cp[:2]), _le(cp[2:])
        if (aps & 3) != 3 or o["profile"] != 0xC05E or o["cluster"] != 0x1000:
            return o
        zcl = take(1)[0]
        if (zcl >> 2) & 1:
            pos += 2
        o["zll_command"] = take(2)[1]
        o["present"] |= ZLL | (ZLL_SCAN_RESPONSE if o["zll_command"] == 1 else 0)
    except Short:
        o["present"] |= MALFORMED
    return o
