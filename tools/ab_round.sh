#!/bin/bash
# A/B: 64-row z ring (4 CTAs of k_zb_rx per SM) against the 128-row ring (3 CTAs) with the same group-level pacing
timeout 900 python -m pytest tests -m gpu -q -x -k "nb or narrow or zigbee or zb or mixed or c5 or shard" 2>&1 | tail -3
L=$PWD/snout_b200/lib
for w in "zb_wb16 4.9" "zb_wb16 9.83" "mixed_wb56 4.9"; do
  python tools/ab_front.py $w 2>&1 | tail -1
  SNRX_LIB=$L/libsnoutrx_R128.so python tools/ab_front.py $w 2>&1 | tail -1
done
