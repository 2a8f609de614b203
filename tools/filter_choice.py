#!/usr/bin/env python3
"""Which prototype filter the headline runs on, by measurement: frame recall of the 40-channel BLE receiver
under near-far conditions (every burst's amplitude drawn from [-S, 0] dB, receiver noise 60 dB below full scale so that what
limits a weak frame is the leakage of its stronger neighbours through the channel filter) and throughput, for the 384-tap
(Kaiser beta 5, ~54 dB) and the 768-tap (beta 9, ~90 dB) prototypes.  Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snout_b200 import synth
from snout_b200.engine import RxEngine

out = {"workload": "ble_wb40, 0.2-s captures made on the GPU, Es/N0 60 dB at full amplitude, gap 200..9000 samples", "rows": []}
for spread in (0.0, 20.0, 40.0, 50.0, 60.0):
    x, truth = synth.wideband_capture_gpu(seconds=0.2, kind="ble", seed=8100, esn0_db=60.0, amp_db_spread=spread)
    sent = {(t.channel, bytes(t.data)) for t in truth if t.start + 2000 < len(x) // 24}
    row = {"amp_spread_db": spread, "frames_sent": len(sent)}
    for taps in (384, 768):
        with RxEngine("ble_wb40", max_samples=len(x), pfb_taps=taps, max_frames=1 << 17) as e:
            fr = e.run(x)
            for _ in range(3):
                e.run(x)
            ms = []
            for _ in range(8):
                e.process(x).poll(copy=False); ms.append(e.stats()["gpu_ms_frontend"])
        dec = {(int(f["channel"]), bytes(f["bytes"][: f["len"]])) for f in fr if f["crc_ok"]}
        row[f"recall_{taps}"] = round(len(sent & dec) / len(sent), 5)
        row[f"channelizer_ms_{taps}"] = round(float(np.median(ms)), 4)
    out["rows"].append(row)
print(json.dumps(out))
