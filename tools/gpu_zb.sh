#!/bin/bash
# GPU box: full parity suite, then the Zigbee / mixed benches with launch lists (and the CTA-tile channelizer for A/B)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
for w in zb_wb16 mixed_wb56; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 | tee gpurun_out/bench_$w.json | cut -c1-330
done
SNRX_ZB_PFB=cta timeout 300 python bench.py --workload zb_wb16 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_zb_wb16_cta.json | cut -c1-330
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_zb_wb16.csv python bench.py --workload zb_wb16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zb_under_ncu.log 2>&1; echo rc=$?
