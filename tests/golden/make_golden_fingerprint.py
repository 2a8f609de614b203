#!/usr/bin/env python3
"""Generates tests/golden/fingerprint_ref.json: Device.vendor / .model / .os of the UNMODIFIED reference
(/root/reference/snout/core/device.py:171-246) on seeded sequences of advertising payloads parsed by the unmodified
BtlePDUPayload (advertising.py) -- `appdirs` / `timeago` stubbed as in make_golden_adv.py.  Run in the build container only;
the GPU box and the tests read the committed JSON.

Each case is one sender: a list of AdvData payloads in the order they were received.  The payloads are built from the
structures the three properties look at (manufacturer data of Apple with Nearby / AirPods / other TLVs, of Microsoft and of
other companies, the 128-bit UUID list with and without the FitBit UUID, flags, service data) in random order and number."""
import importlib
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
for name in ("appdirs", "timeago"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["appdirs"].user_config_dir = lambda *a, **k: "/tmp"
sys.modules["appdirs"].user_data_dir = lambda *a, **k: "/tmp"
sys.path.insert(0, REF)
for pkg in ("snout", "snout.core"):
    m = types.ModuleType(pkg)
    m.__path__ = [os.path.join(REF, *pkg.split("."))]
    sys.modules[pkg] = m
protocols = importlib.import_module("snout.core.protocols")
adv = importlib.import_module("snout.core.protocols.btle.advertising")
device = importlib.import_module("snout.core.device")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import adv_oracle as ao                  # only to skip payloads it flags MALFORMED (the reference may hang on them)


def ad(t, data):
    return bytes([1 + len(data), t]) + bytes(data)


def apple(rng, kinds):
    tlv = b""
    for k in kinds:
        if k == "nearby10":
            tlv += bytes([0x10, 2, int(rng.integers(0, 256)), 0x00])
        elif k == "nearby11":
            tlv += bytes([0x10, 5, int(rng.integers(0, 256)), 0x10]) + bytes(rng.integers(0, 256, 3, dtype=np.uint8))
        elif k == "nearby12":
            tlv += bytes([0x10, 5, int(rng.integers(0, 256)), int(rng.choice([0x18, 0x1C]))]) + bytes(rng.integers(0, 256, 3, dtype=np.uint8))
        elif k == "nearby?":
            n = int(rng.choice([1, 3, 4, 6]))
            tlv += bytes([0x10, n]) + bytes(rng.integers(0x20, 256, n, dtype=np.uint8))
        elif k == "airpods":
            tlv += bytes([0x07, 4]) + bytes(rng.integers(0, 256, 4, dtype=np.uint8))
        elif k == "ibeacon":
            tlv += bytes([0x02, 3]) + bytes(rng.integers(0, 256, 3, dtype=np.uint8))
        elif k == "handoff":
            tlv += bytes([0x0C, 4]) + bytes(rng.integers(0, 256, 4, dtype=np.uint8))
    return ad(0xFF, b"\x4c\x00" + tlv)


FITBIT = bytes.fromhex("ba5689a6fabfa2bd01467d6e00fbabad")


def payload(rng):
    parts = []
    for _ in range(int(rng.integers(0, 4))):
        k = int(rng.integers(0, 10))
        if k == 0:
            parts.append(ad(0x01, [int(rng.integers(0, 32))]))
        elif k == 1:
            parts.append(ad(0x06, FITBIT))
        elif k == 2:
            parts.append(ad(0x06, bytes(rng.integers(0, 256, 16, dtype=np.uint8))))
        elif k == 3:
            parts.append(ad(0xFF, b"\x06\x00" + bytes(rng.integers(0, 256, int(rng.integers(0, 6)), dtype=np.uint8))))
        elif k == 4:
            parts.append(ad(0xFF, bytes(rng.integers(0, 256, 2, dtype=np.uint8)) + bytes(rng.integers(0, 256, 3, dtype=np.uint8))))
        elif k == 5:
            parts.append(ad(0x16, bytes(rng.integers(0, 256, 4, dtype=np.uint8))))
        else:
            kinds = list(rng.choice(["nearby10", "nearby11", "nearby12", "nearby?", "airpods", "ibeacon", "handoff"],
                                    size=int(rng.integers(0, 4))))
            parts.append(apple(rng, kinds))
    out = b""
    for q in parts:                                  # whole structures only: a truncated Apple TLV makes the reference's
        if len(out) + len(q) <= 31:                  # AppleTypeParser loop forever (advertising.py:104-110)
            out += q
    return out


class Msg:                                   # what Device reads of a message (device.py:181-246)
    def __init__(self, p):
        self.protocol = protocols.BTLE
        self.payload = p
        self.timestamp = 0.0


def main():
    rng = np.random.default_rng(20261019)
    cases = []
    for i in range(160):
        advs = [payload(rng) for _ in range(int(rng.integers(1, 6)))]
        msgs = []
        ok = True
        for a in advs:
            if ao.parse_adv_data(a)["present"] & ao.MALFORMED:        # inputs the reference raises on / never returns from
                ok = False
                break
            try:
                msgs.append(Msg(adv.BtlePDUPayload(a)))
            except Exception:
                ok = False                    # the generator only builds payloads the reference parser accepts
        if not ok:
            continue
        d = device.Device(protocols.BTLE, f"{i:012x}")
        d._messages_sent = msgs
        cases.append(dict(adv=[a.hex() for a in advs], vendor=d.vendor, model=d.model, os=d.os))
    # nibble-offset match of the hex substring test (device.py:189): the UUID shifted by four bits
    shifted = bytes.fromhex("0" + FITBIT.hex() + "0")
    d = device.Device(protocols.BTLE, "shifted")
    d._messages_sent = [Msg(adv.BtlePDUPayload(ad(0x06, shifted)))]
    cases.append(dict(adv=[ad(0x06, shifted).hex()], vendor=d.vendor, model=d.model, os=d.os))
    here = os.path.dirname(os.path.abspath(__file__))
    an = importlib.import_module("snout.core.protocols.btle.assigned_numbers")
    ids = set()
    for c in cases:
        for a in c["adv"]:
            r = ao.parse_adv_data(bytes.fromhex(a))
            if r["present"] & ao.MANUFACTURER:
                ids.add(r["company_id"])
    names = {str(i): an.company_ids.get(i, "??") for i in sorted(ids)}      # the names of the ids that occur (fixture data)
    json.dump(dict(company_names=names, source="Device.vendor/.model/.os of the unmodified snout/core/device.py on payloads parsed by the unmodified advertising.py",
                   cases=cases), open(os.path.join(here, "fingerprint_ref.json"), "w"), indent=0)
    from collections import Counter
    print(len(cases), Counter(c["model"] for c in cases), Counter(c["os"] for c in cases), Counter(c["vendor"] for c in cases).most_common(6))


if __name__ == "__main__":
    main()
