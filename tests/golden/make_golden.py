#!/usr/bin/env python3
"""Regenerate the committed golden fixtures from the reference itself.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Every fixture is an OUTPUT OF THE REFERENCE'S OWN CODE (compiled unmodified into oracle/_ref) or
one of the reference's own test vectors, so the tests that read them keep pinning the oracle and
the CUDA path on machines where /root/reference does not exist (the GPU box).

  btle_sample_iq_4msps.npz  vendor/BTLE/matlab/sample_iq_4msps.txt as int8 + the frames the
                            reference receiver finds in it (3 ADV_IND, CRC ok)
  btle_welcome.npz          usrp_replay_example/btle_ch37_iq_float32_welcom_msg.bin (x256 -> int8) + frames
  btle_synth_ref.npz        reference frames + receiver() stdout on seeded synthetic captures
  btle_boundary_ref.npz     reference frames for the golden capture delayed by 403..413 samples
                            (window-boundary duplicate rule)
  btle_mask_ref.npz         reference frames under partial access-address masks (-m)
  btle_tables_ref.npz       scramble_table[40][42], crc_table[256], crc_init_reorder(0x555555)
  zb_sink_ref.npz           CHIP_MAPPING[16] and, for seeded synthetic captures, the frames the
                            reference packet sink publishes when fed the oracle's soft chips
  kats.json                 CRC-24 / FCS-16 known answers from scapy's .uts files
  rftap.pcap                scapy-radio/scapy/test/rftap.pcap (RFtap wire format fixture)
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from snout_b200 import synth  # noqa: E402

REF = "/root/reference"


def frames_to_dict(fr):
    return {k: fr[k] for k in fr.dtype.names}


def main():
    oracle.build(ref=True)
    only = sys.argv[1] if len(sys.argv) > 1 else "all"      # all | ble | zb   (ble fixtures embed wall-clock text)
    if only in ("all", "ble"):
        make_ble()
    if only in ("all", "zb"):
        make_zb_and_kats()
    if only in ("all", "mask"):
        make_mask()
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f:32s} {os.path.getsize(os.path.join(HERE, f)):9d} bytes")


def make_mask():
    """btle_mask_ref.npz: reference frames with a partial access-address mask (-m, btle_rx.c:1395-1401, 2301)."""
    cap = synth.ble_capture(n=400_000, channel=37, seed=12, esn0_db=18, gap=(100, 1500))
    q = oracle.ble_quantize(cap.iq, 128.0)
    out = {"params": np.array([12, 18.0, 37, 400_000])}
    for mask in (0xFFFFFF00, 0x00FFFFFF, 0xFFFF0000, 0xFFFFFFFE):
        out[f"frames_{mask:08x}"] = oracle.ble_decode(q, 37, impl="reference", aa_mask=mask)
    np.savez_compressed(f"{HERE}/btle_mask_ref.npz", **out)


def make_ble():
    # ---- BLE golden capture
    txt = open(f"{REF}/vendor/BTLE/matlab/sample_iq_4msps.txt").read().replace(",", " ").split()
    g = np.array(txt, dtype=np.int64).astype(np.int8).reshape(-1, 2)
    fr = oracle.ble_decode(g, 37, impl="reference")
    assert len(fr) == 3 and fr["crc_ok"].all() and list(fr["sample_index"]) == [97892, 501906, 905891]
    np.savez_compressed(f"{HERE}/btle_sample_iq_4msps.npz", iq=g, frames=fr)

    w = np.fromfile(f"{REF}/vendor/BTLE/usrp_replay_example/btle_ch37_iq_float32_welcom_msg.bin", dtype=np.float32)
    wq = np.clip(np.rint(w * 256.0), -128, 127).astype(np.int8).reshape(-1, 2)
    frw = oracle.ble_decode(wq, 37, impl="reference")
    assert len(frw) == 1 and frw["crc_ok"][0] == 1
    np.savez_compressed(f"{HERE}/btle_welcome.npz", iq=wq, frames=frw, stdout=oracle.ble_reference_stdout(wq, 37))

    # ---- BLE seeded synthetic: reference frames + real receiver() text
    out = {}
    for seed, esn0, ch in ((1001, 30.0, 37), (1002, 24.0, 38), (1003, 20.0, 39), (1004, 26.0, 5)):
        cap = synth.ble_capture(n=600_000, channel=ch, seed=seed, esn0_db=esn0)
        q = oracle.ble_quantize(cap.iq, 128.0)
        f = oracle.ble_decode(q, ch, impl="reference")
        out[f"frames_{seed}"] = f
        out[f"stdout_{seed}"] = oracle.ble_reference_stdout(q, ch)
        out[f"params_{seed}"] = np.array([seed, esn0, ch, 600_000])
    np.savez_compressed(f"{HERE}/btle_synth_ref.npz", **out)

    # ---- window-boundary rule
    out = {}
    for d in range(403, 414):
        gd = np.concatenate([np.zeros((d, 2), np.int8), g[:200_000]])
        out[f"frames_{d}"] = oracle.ble_decode(gd, 37, impl="reference")
    np.savez_compressed(f"{HERE}/btle_boundary_ref.npz", **out)

    wt, ct, ci = oracle.ble_tables("reference")
    np.savez_compressed(f"{HERE}/btle_tables_ref.npz", scramble_table=wt, crc_table=ct, crc_init_internal=np.uint32(ci))



def make_zb_and_kats():
    # ---- Zigbee: reference sink on the oracle's soft chips
    out = {"chip_mapping": oracle.zb_chip_words("reference")}
    for seed, esn0 in ((2001, 30.0), (2002, 12.0), (2003, 9.0)):
        cap = synth.zigbee_capture(n=1_000_000, channel=11, seed=seed, esn0_db=esn0)
        z = oracle.zb_dc_remove(oracle.zb_quad_demod(cap.iq))
        _, chips, _ = oracle.zb_chain(z, 0, len(z), 0, len(z), want_chips=True)
        ref = oracle.zb_sink_reference(chips)
        out[f"end_chip_{seed}"] = np.array([e for e, _ in ref], dtype=np.int64)
        out[f"len_{seed}"] = np.array([len(b) for _, b in ref], dtype=np.int32)
        by = np.zeros((len(ref), 128), dtype=np.uint8)
        for i, (_, b) in enumerate(ref):
            by[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
        out[f"bytes_{seed}"] = by
        out[f"params_{seed}"] = np.array([seed, esn0, 11, 1_000_000])
    np.savez_compressed(f"{HERE}/zb_sink_ref.npz", **out)

    kats = {
        "ble_crc24": {"pdu_hex": "0006000000000000", "crc_tx_hex": "5a3960",
                      "source": "scapy-radio/scapy/test/bluetooth4LE.uts:12-17"},
        "fcs16": [],
        "source_fcs": "scapy-radio/scapy/test/dot15d4.uts:85-115",
    }
    import re
    uts = open(f"{REF}/scapy-radio/scapy/test/dot15d4.uts").read()
    for name, want in (("ieee802_firstfrag", 0xb539), ("ieee802_secfrag", 0xac66), ("ieee802_iphc", 0x16c1)):
        m = re.search(name + r' = b"([^"]+)"', uts)
        raw = eval('b"' + m.group(1) + '"')
        kats["fcs16"].append({"name": name, "frame_hex": raw.hex(), "fcs": want})
    json.dump(kats, open(f"{HERE}/kats.json", "w"), indent=1)
    shutil.copyfile(f"{REF}/scapy-radio/scapy/test/rftap.pcap", f"{HERE}/rftap.pcap")
    os.chmod(f"{HERE}/rftap.pcap", 0o644)


if __name__ == "__main__":
    main()
