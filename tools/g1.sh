mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for t in 10 40; do
for w in zb_wb16 mixed_wb56; do
  timeout 300 python bench.py --workload $w --tiles $t --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_${w}_t$t.json | cut -c1-330
done
done
timeout 300 python bench.py --workload zb_nb --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_zb_nb.json | cut -c1-330
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_zb_wb16.csv python bench.py --workload zb_wb16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?
