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
HINT_NEARBY_MASK, HINT_FITBIT = 0x0F, 0x10
FITBIT_UUID = "ba5689a6fabfa2bd01467d6e00fbabad"
MODEL_NONE, MODEL_FITBIT_CHARGE, MODEL_AIRPODS = 0, 1, 2
OS_NONE, OS_UNDECIDED, OS_IOS10, OS_IOS11, OS_IOS12, OS_WINDOWS10 = range(6)
VENDOR_NONE, VENDOR_COMPANY, VENDOR_FITBIT = 0, 1, 2
ADV_DATA_PDUS = (0, 2, 4, 6)          # AdvA + AD structures ("AdvA:.. Data:.." lines, message.py:226-233)


def parse_adv_data(adv: bytes) -> dict:
    """AD structures -> summary fields (the dict keys mirror include/snoutrx.h snrx_adv_t)."""
    o = dict(n_ad=0, ad_flags=0, present=0, company_id=0xFFFF, service_uuid=0xFFFF, unknown_type=0, apple_action=0xFF,
             oob_flags=0, apple_types=0, hints=0)
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
            o["hints"] &= ~HINT_FITBIT                                 # the dict entry is replaced: last one wins
            if FITBIT_UUID in bytes(v).hex():                          # device.py:187-190: substring of the hex text
                o["hints"] |= HINT_FITBIT
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
                    o["hints"] &= ~HINT_NEARBY_MASK
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
                                if not o["hints"] & HINT_NEARBY_MASK:  # Device.os stops at the first Nearby record (device.py:238-241)
                                    nd, hint = ad[1:], 1               # advertising.py:272-281
                                    if len(nd) == 1 and nd[0] == 0x00:
                                        hint = 2
                                    if len(nd) == 4:
                                        if nd[0] == 0x10:
                                            hint = 3
                                        if nd[0] in (0x18, 0x1C):
                                            hint = 4
                                    o["hints"] |= hint
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
             service_uuid=0xFFFF, unknown_type=0, apple_action=0xFF, oob_flags=0, apple_types=0, hints=0)
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
                                   first=pos, last=pos, company=(None, 0xFFFF), votes=[]))
        if f["crc_ok"]:                                               # only CRC0 lines become messages (message.py:225-226)
            d["votes"].append((pos, vendor_vote(a), model_vote(a), os_vote(a)))
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
    for d in tab.values():                                             # Device.vendor / .model / .os: the first packet that decides
        d["vendor"], d["model"], d["os"] = (VENDOR_NONE, 0xFFFF), MODEL_NONE, OS_NONE
        for _, v, m, o in sorted(d.pop("votes"), key=lambda t: t[0]):
            if v and d["vendor"][0] == VENDOR_NONE:
                d["vendor"] = v
            if m and d["model"] == MODEL_NONE:
                d["model"] = m
            if o and d["os"] == OS_NONE:
                d["os"] = o
    return tab


def vendor_vote(a: dict):
    """device.py:171-191 for one message: (kind, company id) or None."""
    if a["present"] & MALFORMED:
        return None                                                    # the reference parser raised: there is no message
    if a["present"] & MANUFACTURER:
        return (VENDOR_COMPANY, a["company_id"])
    if a["hints"] & HINT_FITBIT:
        return (VENDOR_FITBIT, 0xFFFF)
    return None


def model_vote(a: dict) -> int:
    """device.py:193-220 for one message."""
    if a["present"] & MALFORMED:
        return MODEL_NONE
    if a["hints"] & HINT_FITBIT:
        return MODEL_FITBIT_CHARGE
    if a["present"] & MANUFACTURER and a["company_id"] == 0x004C and a["apple_types"] & (1 << 7):
        return MODEL_AIRPODS
    return MODEL_NONE


def os_vote(a: dict) -> int:
    """device.py:222-246 for one message."""
    if a["present"] & MALFORMED or not a["present"] & MANUFACTURER:
        return OS_NONE
    if a["company_id"] == 0x004C and a["hints"] & HINT_NEARBY_MASK:
        return a["hints"] & HINT_NEARBY_MASK
    if a["company_id"] == 0x0006:
        return OS_WINDOWS10
    return OS_NONE


def fold_messages(advs: list[bytes]) -> dict:
    """Device.vendor / .model / .os of one sender whose messages carry these AdvData payloads, in order."""
    out = dict(vendor=(VENDOR_NONE, 0xFFFF), model=MODEL_NONE, os=OS_NONE)
    for adv in advs:
        a = parse_adv_data(adv)
        v, m, o = vendor_vote(a), model_vote(a), os_vote(a)
        if v and out["vendor"][0] == VENDOR_NONE:
            out["vendor"] = v
        if m and out["model"] == MODEL_NONE:
            out["model"] = m
        if o and out["os"] == OS_NONE:
            out["os"] = o
    return out
