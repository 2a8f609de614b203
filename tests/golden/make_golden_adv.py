#!/usr/bin/env python3
"""Generates tests/golden/adv_ref.json by importing the REFERENCE advertising parser
(/root/reference/snout/core/protocols/btle/advertising.py, with `appdirs` / `timeago` stubbed: they are not installed
here and not used by the parser) and running it on a seeded set of AdvData payloads.  Run in the build container only;
the GPU box and the tests read the committed JSON.

Payloads are assembled from well-formed and truncated AD structures.  Inputs on which the reference raises IndexError
(empty flag fields ...) are recorded with "raises": true; inputs on which AppleTypeParser never returns (a dangling type
byte, advertising.py:104-110) are not generated."""
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

# the package __init__ files pull in the CLI and its dependencies: register empty packages that only carry the real
# __path__, so that the parser module and its `.assigned_numbers` import load unmodified from the reference tree
import importlib

for pkg in ("snout", "snout.core", "snout.core.protocols", "snout.core.protocols.btle"):
    m = types.ModuleType(pkg)
    m.__path__ = [os.path.join(REF, *pkg.split("."))]
    sys.modules[pkg] = m
sys.modules["snout.core.protocols"].BTLE = "btle"
adv = importlib.import_module("snout.core.protocols.btle.advertising")
an = importlib.import_module("snout.core.protocols.btle.assigned_numbers")


def jsonable(x):
    if isinstance(x, (bytes, bytearray)):
        return {"__bytes__": bytes(x).hex()}
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    return x


def structure(t, data):
    return bytes([1 + len(data), t]) + bytes(data)


def apple_tlvs(rng):
    out = b""
    for _ in range(int(rng.integers(1, 4))):
        t = int(rng.choice([0x02, 0x05, 0x07, 0x09, 0x0A, 0x0C, 0x0D, 0x0E, 0x0F, 0x10, 0x03, 0x12]))
        n = {0x0C: 3, 0x0D: 4, 0x0E: 8, 0x10: int(rng.choice([1, 2, 5]))}.get(t, int(rng.integers(0, 6)))
        out += bytes([t, n]) + bytes(rng.integers(0, 256, n, dtype=np.uint8))
    return out


def main():
    rng = np.random.default_rng(20261017)
    vectors = [bytes.fromhex("0201060aff4c001005011c569415"),          # the example of message.py:214
               b"", bytes([0]), bytes([0, 0, 0]), bytes.fromhex("02011a"), bytes.fromhex("0303aafe")]
    for _ in range(160):
        parts = []
        for _ in range(int(rng.integers(1, 5))):
            kind = int(rng.integers(0, 8))
            if kind == 0:
                parts.append(structure(0x01, [int(rng.integers(0, 32))]))
            elif kind == 1:
                parts.append(structure(0x06, rng.integers(0, 256, 16, dtype=np.uint8)))
            elif kind == 2:
                parts.append(structure(0x11, [int(rng.integers(0, 16))]))
            elif kind == 3:
                parts.append(structure(0x16, rng.integers(0, 256, int(rng.integers(2, 8)), dtype=np.uint8)))
            elif kind == 4:
                parts.append(structure(0xFF, bytes([0x4C, 0x00]) + apple_tlvs(rng)))
            elif kind == 5:
                cid = int(rng.choice([0x0006, 0x0075, 0x00E0, 0x1234]))
                parts.append(structure(0xFF, bytes([cid & 255, cid >> 8]) + bytes(rng.integers(0, 256, int(rng.integers(0, 8)), dtype=np.uint8))))
            elif kind == 6:
                parts.append(structure(int(rng.choice([0x02, 0x03, 0x08, 0x09, 0x0A, 0x19])), rng.integers(0, 256, int(rng.integers(0, 6)), dtype=np.uint8)))
            else:
                parts.append(bytes([0]))                              # zero-length structure
        payload = b"".join(parts)[:31]                                 # AdvData is at most 31 bytes: the last structure may be cut
        vectors.append(payload)
    out = []
    for v in vectors:
        # cut Apple payloads can leave a dangling TLV type byte: skip what would hang the reference
        try:
            import signal
            signal.signal(signal.SIGALRM, lambda *a: (_ for _ in ()).throw(TimeoutError()))
            signal.alarm(2)
            d = adv.BtlePDUPayload(v).dict()
            signal.alarm(0)
            out.append({"adv": v.hex(), "dict": jsonable(d)})
        except TimeoutError:
            out.append({"adv": v.hex(), "hangs": True})
        except (IndexError, KeyError):
            signal.alarm(0)
            out.append({"adv": v.hex(), "raises": True})
    names = {"flags": [adv.FLAG_LEL, adv.FLAG_LEG, adv.FLAG_BR, adv.FLAG_SLEBR, adv.FLAG_LEBRS], "oob_present": adv.FLAG_OOB,
             "keys": {"flags": adv.FLAGS, "oob": adv.SEC_MG_OOB_FLAGS, "service": adv.SERVICE_DATA, "manufacturer": adv.MANUFACTURER_SPECIFIC,
                      "unknown": adv.UNKNOWN, "company_id": adv.COMPANY_ID, "uuid128": an.ad_types[0x06]["name"]},
             "apple_types": {v: k for k, v in adv.APPLE_DATA_TYPES.items()}}
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump({"source": "snout/core/protocols/btle/advertising.py (imported unmodified)", "names": names, "vectors": out},
              open(os.path.join(here, "adv_ref.json"), "w"), indent=0)
    print(len(out), "vectors,", sum("raises" in o for o in out), "raise,", sum("hangs" in o for o in out), "hang")


if __name__ == "__main__":
    main()
