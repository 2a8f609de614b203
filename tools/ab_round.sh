#!/bin/bash
# A/B against the previous commit's library (snout_b200/lib/libsnoutrx_PREV.so)
timeout 900 python -m pytest tests -m gpu -q -x -k "wb or wideband or mixed or pfb or chan" 2>&1 | tail -2
L=$PWD/snout_b200/lib
for v in _PREV "" _PREV ""; do
  echo "== ble_wb40 lib$v"; SNRX_LIB=$L/libsnoutrx$v.so python tools/ab_serial.py 2>&1 | tail -1 | cut -c1-150
done
for v in _PREV ""; do
  SNRX_LIB=$L/libsnoutrx$v.so python tools/ab_front.py zb_wb16 4.9 2>&1 | tail -1
done
