mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exchange or zb_nb_parity" 2>&1 | tail -5
for w in ble_nb zb_nb; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 4 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_$w.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$w', round(j['value']), j['ms_per_step'], 'frac', round(j['roofline']['frac'],3), 'step_frac', round(j['roofline']['step_frac'],3), 'e2e', round(j['e2e']['value']), 'single', j['config']['single_capture'], 'cpu', j.get('cpu_baseline',{}).get('value'))"
done
timeout 400 python bench.py --taps 768 --steps 10 --warmup 3 --no-cpu-baseline --no-c5 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_ble_wb40_768.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('768 taps', round(j['value']), j['ms_per_step'], 'frac', round(j['roofline']['frac'],3))"
python - <<'PY'
import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import numpy as np, oracle
from gen_tables import PFB_DESIGNS, kaiser_lowpass
from snout_b200 import chanplan
oracle.build(native=True)
h = kaiser_lowpass(*PFB_DESIGNS["BLE_384"])
rng = np.random.default_rng(0)
x = (rng.standard_normal(4_800_000) + 1j*rng.standard_normal(4_800_000)).astype(np.complex64)
bins = [chanplan.ble_channel_bin(c) for c in range(40)]
for thr in (1, 4, 8, 16, os.cpu_count()):
    oracle.set_threads(thr)
    best = 1e9
    for rep in range(3):
        t0=time.perf_counter(); y = oracle.pfb(x, h, bins, fast=True, native=True); best=min(best,time.perf_counter()-t0)
    print('cpu channelizer threads',thr, f'{len(x)/best/1e6:.1f} Msamples/s')
PY
