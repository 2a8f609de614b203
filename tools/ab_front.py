#!/usr/bin/env python3
"""A/B aid: front-end kernel time and whole-batch time of one resident wideband capture for any workload, one batch at a
time and two in flight.  usage: SNRX_LIB=... python tools/ab_front.py <ble_wb40|zb_wb16|mixed_wb56> [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from snout_b200 import synth
from snout_b200.engine import RxEngine
wl = sys.argv[1] if len(sys.argv) > 1 else "ble_wb40"
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 0.983
kind = {"ble_wb40": "ble", "zb_wb16": "zigbee", "mixed_wb56": "mixed"}[wl]
rep = max(1, int(round(sec / 0.0983)))
x, _ = synth.wideband_capture_gpu(seconds=0.0983, kind=kind, seed=4000, device=0, repeat=rep)
eng = RxEngine(wl, max_samples=len(x), device=0, max_frames=1 << 19)
def serial(n):
    fe, tot, nf = [], [], 0
    for i in range(n):
        eng.process(x); fr = eng.poll(copy=False); nf = len(fr)
        s = eng.stats(); fe.append(s["gpu_ms_frontend"]); tot.append(s["gpu_ms"])
    return np.median(fe[2:]), np.median(tot[2:]), nf
def piped(n):
    eng.process(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        eng.process(x); eng.poll(copy=False)
    eng.poll(copy=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
a = serial(8); b = piped(12)
print(f"{wl} {sec:.2f}s lib={os.path.basename(os.environ.get('SNRX_LIB','default'))}: serial frontend {a[0]:.4f} ms, batch {a[1]:.4f} ms, frames {a[2]} | two in flight wall/step {b:.4f} ms = {len(x)/b/1e6:.1f} Gsamples/s")
