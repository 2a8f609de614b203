"""TEST INFRASTRUCTURE ONLY -- Python restatement of Snout's advertising parser for the analytics row.

Follows snout/core/protocols/btle/advertising.py statement by statement:
  AdvDataParser.get_ad_structure  :74-91     (len, data) fields, slices truncate silently
  BtlePDUPayload.parse_ad_structure :139-159  dispatch on the AD type, `if data:`
  parse_ad_type_0x01/0x06/0x11/0x16/0xff :161-232, word16be :58-60 (little endian despite its name)
  AppleTypeParser.get_type_data :93-110, parse_man_data_apple :224-290
The reference stores results in a dict (a repeated AD type overwrites the earlier one: last wins) and raises IndexError on
empty flag / short service / short manufacturer fields; on a dangling Apple TLV type byte it never returns.  Those inputs
are flagged MALFORMED here (and in csrc/ble_adv.cuh) instead.  Pinned against the imported reference on
tests/golden/adv_ref.json (tests/golden/make_golden_adv.py)."""
from __future__ import annotations

import numpy as np

FLAGS, UUID128, OOB, SERVICE_DATA, MANUFACTURER, UNKNOWN, MALFORMED, SENDER = (1 << i for i in range(8))
ADV_DATA_PDUS = (0, 2, 4, 6)          # AdvA + AD structures ("AdvA:.. Data:.." lines, message.py:226-233)


def parse_adv_data(adv: bytes) -> dict:
    """AD structures -> summary fields (the dict keys mirror include/snoutrx.h snrx_adv_t)."""
    o = dict(n_ad=0, ad_flags=0, present=0, company_id=0xFFFF, service_uuid=0xFFFF, unknown_type=0, apple_action=0xFF,
             oob_flags=0, apple_types=0)
    pos = 0
    while pos < len(adv):
        ad_len = adv[pos]
        data = adv[pos + 1: pos + ad_len + 1]
        pos += 1 + ad_len
        o["n_ad"] += 1
        if not data:
            continue
        t, v = data[0], data[1:]
        if t == 0x01:
            if v:
                o["ad_flags"] = v[0]; o["present"] |= FLAGS
            else:
                o["present"] |= MALFORMED
        elif t == 0x06:
            o["present"] |= UUID128
        elif t == 0x11:
            if v:
                o["oob_flags"] = v[0]; o["present"] |= OOB
            else:
                o["present"] |= MALFORMED
        elif t == 0x16:
            if len(v) >= 2:
                o["service_uuid"] = v[0] | (v[1] << 8); o["present"] |= SERVICE_DATA
            else:
                o["present"] |= MALFORMED
        elif t == 0xFF:
            if len(v) >= 2:
                o["company_id"] = v[0] | (v[1] << 8); o["present"] |= MANUFACTURER
                if o["company_id"] == 0x004C:
                    o["apple_types"], o["apple_action"] = 0, 0xFF
                    man, p = v[2:], 0
                    while p < len(man):
                        if p + 1 >= len(man):
                            o["present"] |= MALFORMED
                            break
                        at, al = man[p], man[p + 1]
                        ad = man[p + 2: p + al + 2]
                        p += 2 + al
                        if at < 32:
                            o["apple_types"] |= 1 << at
                        if at == 0x10:
                            if ad:
                                o["apple_action"] = ad[0] & 0x0F
                            else:
                                o["present"] |= MALFORMED
                        if at == 0x0C and len(ad) < 3:                # Handoff: apple_data[0], word16be(apple_data[1:3])
                            o["present"] |= MALFORMED
            else:
                o["present"] |= MALFORMED
        else:
            o["unknown_type"] = t; o["present"] |= UNKNOWN
    return o


def parse_record(pdu: bytes) -> dict:
    """header(2) | payload | crc(3) of one BLE record -> summary (all snrx_adv_t fields except `frame`)."""
    o = dict(adv_a=bytes(6), pdu_type=0xFF, tx_add=0, rx_add=0, adv_len=0, n_ad=0, ad_flags=0, present=0, company_id=0xFFFF,
             service_uuid=0xFFFF, unknown_type=0, apple_action=0xFF, oob_flags=0, apple_types=0)
    if len(pdu) < 5:
        o["present"] |= MALFORMED
        return o
    o["pdu_type"], o["tx_add"], o["rx_add"] = pdu[0] & 0x0F, (pdu[0] >> 6) & 1, (pdu[0] >> 7) & 1
    p = pdu[2: len(pdu) - 3]
    if o["pdu_type"] not in ADV_DATA_PDUS:
        if o["pdu_type"] == 5 and len(p) >= 12:
            o["adv_a"] = bytes(p[6:12]); o["present"] |= SENDER
        elif o["pdu_type"] in (1, 3) and len(p) >= 6:
            o["adv_a"] = bytes(p[:6]); o["present"] |= SENDER
        return o
    if len(p) < 6:
        o["present"] |= MALFORMED
        return o
    o["adv_a"] = bytes(p[:6])
    o["present"] |= SENDER
    o["adv_len"] = len(p) - 6
    d = parse_adv_data(bytes(p[6:]))
    d["present"] |= o["present"]
    o.update(d)
    return o


def summarize(frames: np.ndarray) -> list[dict]:
    out = []
    for f in frames:
        if f["proto"] != 3:
            out.append(parse_record(b""))
            out[-1]["present"] = 0
            continue
        out.append(parse_record(bytes(f["bytes"][: min(int(f["len"]), 47)])))
    return out


def devices(frames: np.ndarray) -> dict:
    """Sender table over `frames` in any order: {(adv_a bytes, tx_add): fields of snrx_device_t}."""
    tab: dict = {}
    for f, a in zip(frames, summarize(frames)):
        if not a["present"] & SENDER:
            continue
        k = (a["adv_a"], a["tx_add"])
        pos = (int(f["capture_id"]), int(f["sample_index"]))
        d = tab.setdefault(k, dict(packets=0, crc_ok=0, chan_mask=0, pdu_mask=0, present=0, ad_flags=0, apple_types=0,
                                   first=pos, last=pos, company=(None, 0xFFFF)))
        d["packets"] += 1
        d["crc_ok"] += int(f["crc_ok"])
        d["chan_mask"] |= 1 << int(f["channel"])
        d["pdu_mask"] |= 1 << a["pdu_type"]
        d["present"] |= a["present"]
        d["ad_flags"] |= a["ad_flags"]
        d["apple_types"] |= a["apple_types"]
        d["first"], d["last"] = min(d["first"], pos), max(d["last"], pos)
        if a["present"] & MANUFACTURER:
            cand = (pos, a["company_id"])
            if d["company"][0] is None or cand > d["company"]:
                d["company"] = cand
    return tab
