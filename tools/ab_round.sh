#!/bin/bash
# persistent multi-tile channelizer (k_pfb_ble_run, rewritten without divisions / ELECT loops): parity, then timing per tiles-per-CTA
SNRX_PFB_TILES=4 timeout 900 python -m pytest tests -m gpu -q -x -k "ble_wb or wideband or mixed or c5" 2>&1 | tail -2
for t in 0 2 4 8 16; do
  echo "== tiles $t"; SNRX_PFB_TILES=$t python tools/ab_serial.py 2>&1 | tail -1 | cut -c1-150
done
