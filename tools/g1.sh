mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed or zb_wb16_stage" 2>&1 | tail -3
for v in t1 np t2; do
  export SNRX_LIB=$PWD/snout_b200/lib/libsnoutrx_$v.so
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c5 2>&1 | grep '^{' | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$v', round(j['value']), j['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'], 'frac', round(j['roofline']['frac'],3))"
done
