#!/bin/bash
# GPU box (1 GPU): ncu launch list of one bench run + one full capture (with sources) of the dominant kernels.
#   gpurun -- bash tools/gpu_profile.sh [workload]       -> gpurun_out/launches*.csv, gpurun_out/prof_*.ncu-rep
# Summaries for profiles/: python tools/ncu_summary.py gpurun_out/prof_pfb.ncu-rep > profiles/rNN_pfb_ncu_vM.json
W=${1:-ble_wb40}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
K="regex:k_pfb_ble"; [ "$W" != "ble_wb40" ] && K="regex:k_zb_rx|k_pfb_zb_warp|k_zb_iir_sum|k_pfb_ble"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 6 -c 3 -o gpurun_out/prof_$W -f \
    python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$W.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
