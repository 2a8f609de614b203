#!/bin/bash
# GPU box (1 GPU): parity suite, smoke, headline + zb_wb16 bench lines, launch list -- the short form of gpu_all.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench.json | cut -c1-200
timeout 300 python bench.py --workload zb_wb16 --steps 10 --warmup 3 --cpu-seconds 6 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_zb_wb16.json | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?
