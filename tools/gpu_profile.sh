#!/bin/bash
# Runs on the GPU box: ncu launch list of one bench run + one full capture of the channelizer kernel.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pfb_ble -s 2 -c 1 -o gpurun_out/prof_pfb -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_pfb.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
