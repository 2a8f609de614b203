#!/bin/bash
L=$PWD/snout_b200/lib
for v in _NB2 "" _NB2 ""; do
  SNRX_LIB=$L/libsnoutrx$v.so timeout 300 python bench.py --workload ble_nb --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/ab_ble_nb$v.json
  python - <<P
import json
d=json.load(open('gpurun_out/ab_ble_nb$v.json'))
print('ble_nb lib$v value',round(d['value']),'ms',round(d['ms_per_step'],3),'live',round(d['roofline']['kernel_ms'],3),'alone',round(d['roofline']['alone']['kernel_ms'],3),'frac alone',round(d['roofline']['alone']['frac'],3))
P
done
