#!/bin/bash
# A/B of the BLE channelizer launch shapes: SNRX_PFB_TILES = tiles per CTA of k_pfb_ble_run (0 = one-tile kernel only)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wb40" 2>&1 | tail -3
for t in ${AB_TILES:-0 2 4 8 16 0}; do
  SNRX_PFB_TILES=$t timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-c5 2>&1 | grep '^{' | tail -1 > gpurun_out/ab_pfb_$t.json
  python - <<P
import json
d=json.loads(open("gpurun_out/ab_pfb_$t.json").read())
print("tiles=$t value=%.0f ms/step=%.4f kernel_ms=%.4f frac=%.4f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"]))
P
done
