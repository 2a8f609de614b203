#!/bin/bash
# GPU box (1 GPU): full parity suite, smoke, then every bench workload (headline first) and the launch lists
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench.json | cut -c1-200
for w in zb_wb16 mixed_wb56 ble_nb zb_nb; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 6 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_$w.json | cut -c1-250
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?
