#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, bench, launch list, one full ncu capture of the
# channelizer.  Output -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
if [ -n "$SECONDARY" ]; then
for w in zb_wb16 mixed_wb56; do
  echo "== bench $w"
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --cpu-seconds 6 2>&1 | tail -1 | tee gpurun_out/bench_$w.log
done
fi
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pfb_ble -s 2 -c 1 -o gpurun_out/prof_pfb -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_pfb.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
