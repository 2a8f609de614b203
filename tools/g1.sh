mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "zb or mixed" 2>&1 | tail -5
for v in a b; do
  if [ $v = b ]; then export SNRX_LIB=$PWD/snout_b200/lib/libsnoutrx_b.so; fi
  for t in 10 100; do
    timeout 300 python bench.py --workload zb_wb16 --tiles $t --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_zb_wb16_${v}_t$t.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$v', $t, round(j['value']), j['ms_per_step'])"
  done
done
unset SNRX_LIB
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_zb_wb16_t100.csv python bench.py --workload zb_wb16 --tiles 100 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?
