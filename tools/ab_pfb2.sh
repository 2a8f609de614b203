#!/bin/bash
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-c5 2>&1 | grep '^{' | tail -1 > gpurun_out/ab_$name.json
  python - <<P
import json
d=json.loads(open("gpurun_out/ab_$name.json").read())
print("$name value=%.0f ms/step=%.4f kernel_ms=%.4f frac=%.4f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"]))
P
}
run base SNRX_PFB_TILES=0
run t2_strided SNRX_PFB_TILES=2 SNRX_PFB_ORDER=1
run t4_strided SNRX_PFB_TILES=4 SNRX_PFB_ORDER=1
run t8_strided SNRX_PFB_TILES=8 SNRX_PFB_ORDER=1
run t54_strided SNRX_PFB_TILES=54 SNRX_PFB_ORDER=1
SNRX_PFB_TILES=8 SNRX_PFB_ORDER=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_pfb_ble_run" -s 6 -c 1 -o gpurun_out/prof_pfb_run -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/prof_pfb_run.log 2>&1
echo ncu rc=$?
