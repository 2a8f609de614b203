#!/usr/bin/env python3
"""A/B aid: whole-batch wall time of one resident wideband capture with DEPTH batches in flight (a library built with
-DSNRX_LANES=DEPTH).  usage: SNRX_LIB=... python tools/ab_depth.py <workload> <seconds> <depth>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snout_b200 import synth
from snout_b200.engine import RxEngine
wl, sec, depth = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
kind = {"ble_wb40": "ble", "zb_wb16": "zigbee", "mixed_wb56": "mixed"}[wl]
rep = max(1, int(round(sec / 0.0983)))
x, _ = synth.wideband_capture_gpu(seconds=0.0983, kind=kind, seed=4000, device=0, repeat=rep)
eng = RxEngine(wl, max_samples=len(x), device=0, max_frames=1 << 19)
for _ in range(3):
    eng.process(x); eng.poll(copy=False)
n = 12
for _ in range(depth - 1): eng.process(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(n):
    eng.process(x); eng.poll(copy=False)
for _ in range(depth - 1): eng.poll(copy=False)
torch.cuda.synchronize()
b = (time.perf_counter() - t0) / n * 1e3
print(f"{wl} {sec:.2f}s depth {depth} lib={os.path.basename(os.environ.get('SNRX_LIB','default'))}: wall/step {b:.4f} ms = {len(x)/b/1e6:.1f} Gsamples/s")
