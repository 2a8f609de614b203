#!/bin/bash
# GPU box: parity tests, then the bench for each channelizer tile width
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in 2 4 1; do
  echo "== SNRX_PFB_WARPS=$w"
  SNRX_PFB_WARPS=$w timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['value'])"
done
