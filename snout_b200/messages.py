"""Batch builders of Snout's per-packet message objects from GPU-decoded records.

`snout zigbee scan` turns every datagram of the receive flowgraph into scapy objects one at a time
(snout/util/zigbee.py:194-202 handle_packet: RFtap(pkt), wrpcap(pkt.payload); snout/core/message.py:258-304
ZigbeeMessage.fromraw: sender = src_addr, receiver = dest_addr, seq_number = seqnum, pan = {src_panid, dest_panid},
rftap = the RFtap fields, vuln = {'zll': haslayer(ZLLScanResponse)}; snout/util/zigbee.py:176-192 collects the ZLL scan
responses of a touchlink scan).  Here the MAC headers of a whole batch are read on the GPU (RxEngine.zb_mac_summary ->
csrc/zb_mac.cuh) and the same dictionaries are built from the summary rows without any per-packet dissection."""
from __future__ import annotations

import numpy as np

from . import _abi


def _addr(value: int, mode: int):
    """An address the way scapy's dot15d4AddressField holds it: an int (16 or 64 bit), None when absent."""
    return int(value) if mode in (2, 3) else None


def zigbee_messages(frames: np.ndarray, mac: np.ndarray, timestamp: float | None = None) -> list[dict]:
    """One dictionary per 802.15.4 record with the keyword arguments ZigbeeMessage.fromraw (message.py:286-303) passes to
    ZigbeeMessage(...): sender, receiver, seq_number, timestamp, pan, rftap, vuln -- plus channel, crc_ok and the record index.
    `frames`: records of a batch (RxEngine.poll), `mac`: RxEngine.zb_mac_summary() of the same batch."""
    out = []
    P = _abi
    for f, m in zip(frames, mac):
        if m["present"] & P.ZBMAC_NOT_ZIGBEE:
            continue
        pr = int(m["present"])
        out.append(dict(
            sender=_addr(m["src_addr"], int(m["src_mode"])) if pr & P.ZBMAC_SRC_ADDR else None,
            receiver=_addr(m["dest_addr"], int(m["dest_mode"])) if pr & P.ZBMAC_DEST_ADDR else None,
            seq_number=int(m["seqnum"]),
            timestamp=timestamp,
            pan=dict(src_panid=int(m["src_panid"]) if pr & P.ZBMAC_SRC_PANID else None,
                     dest_panid=int(m["dest_panid"]) if pr & P.ZBMAC_DEST_PANID else None),
            rftap=dict(dlt=195, qual=float(f["lqi"]) / 255.0),                       # epy_block_0.py:21, rftap_encap(2, 195)
            vuln=dict(zll=bool(pr & P.ZBMAC_ZLL_SCAN_RESPONSE)),
            channel=int(f["channel"]), crc_ok=bool(f["crc_ok"]), frame_type=int(m["frame_type"]), record=int(m["frame"]),
        ))
    return out


def zll_scan_responses(frames: np.ndarray, mac: np.ndarray) -> list[dict]:
    """What ZigbeeZLLScan.handle_packet keeps (zigbee.py:176-192): {'pkt': the 802.15.4 frame, 'qual', 'channel'} for every
    record whose dissection holds a ZLLScanResponse."""
    sel = (mac["present"] & _abi.ZBMAC_ZLL_SCAN_RESPONSE) != 0
    return [dict(pkt=bytes(f["bytes"][: int(f["len"])]), qual=float(f["lqi"]) / 255.0, channel=int(f["channel"]))
            for f in frames[mac["frame"][sel]]]


# ----------------------------------------------------------------------------------------------------- SURVEY 8(f) N1
# A few well-known Bluetooth SIG company identifiers; a Snout installation passes its own table
# (snout.core.protocols.btle.assigned_numbers.company_ids) as `company_names`.
COMPANY_NAMES = {0x0006: "Microsoft", 0x004C: "Apple, Inc.", 0x0075: "Samsung Electronics Co. Ltd.", 0x00E0: "Google", 0x0087: "Garmin International, Inc.",
                 0x0059: "Nordic Semiconductor ASA"}
_OS_TEXT = {_abi.OS_NONE: "-", _abi.OS_UNDECIDED: "-", _abi.OS_IOS10: "iOS 10", _abi.OS_IOS11: "iOS 11", _abi.OS_IOS12: "iOS 12",
            _abi.OS_WINDOWS10: "Windows 10 >= v10.0.10240.0"}
_MODEL_TEXT = {_abi.MODEL_NONE: "-", _abi.MODEL_FITBIT_CHARGE: "Charge / Charge HR", _abi.MODEL_AIRPODS: "AirPods"}


def device_fingerprints(devices: np.ndarray, company_names: dict | None = None) -> list[dict]:
    """Device.vendor / .model / .os (snout/core/device.py:171-246) of every row of the GPU sender table (RxEngine.devices()):
    the three answers are decided on the GPU by each sender's FIRST deciding packet (csrc/ble_adv.cuh); this only renders the
    text.  Unknown company ids read '??' as in the reference (advertising.py:225)."""
    names = COMPANY_NAMES if company_names is None else company_names
    out = []
    for d in devices:
        kind = int(d["vendor_kind"])
        vendor = "-" if kind == _abi.VENDOR_NONE else "FitBit" if kind == _abi.VENDOR_FITBIT else names.get(int(d["vendor_company"]), "??")
        out.append(dict(address=bytes(d["adv_a"])[::-1].hex(), tx_add=int(d["tx_add"]), vendor=vendor, model=_MODEL_TEXT[int(d["model"])],
                        os=_OS_TEXT[int(d["os"])]))
    return out
