#!/bin/bash
# GPU box (1 GPU): one full ncu capture (with sources) of a zb_wb16 step's kernels -> gpurun_out/prof_zb_wb16.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_zb_rx|k_pfb_zb_warp|k_zb_iir_sum|k_zb_order" -s 10 -c 5 -o gpurun_out/prof_zb_wb16 -f \
    python bench.py --workload zb_wb16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_zb_wb16.log 2>&1
echo "full capture rc=$?"
