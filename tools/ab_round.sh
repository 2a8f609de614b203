#!/bin/bash
# A/B: two vs three batches in flight (library built with -DSNRX_LANES=3) on the Zigbee / mixed workloads, then the parity soak
L=$PWD/snout_b200/lib
for w in "zb_wb16 4.9" "zb_wb16 9.83" "mixed_wb56 4.9"; do
  python tools/ab_depth.py $w 2 2>&1 | tail -1
  SNRX_LIB=$L/libsnoutrx_L3.so python tools/ab_depth.py $w 2 2>&1 | tail -1
  SNRX_LIB=$L/libsnoutrx_L3.so python tools/ab_depth.py $w 3 2>&1 | tail -1
done
FUZZ_BLE=40 FUZZ_ZB=60 timeout 600 python tools/fuzz_parity.py 2>&1 | tail -3 | tee gpurun_out/fuzz_parity.log
