#!/bin/bash
# A/B of k_zb_rx's chain order (SNRX_ZB_ORDER=0: natural order) on zb_wb16 / mixed_wb56
timeout 900 python -m pytest tests -m gpu -q -x -k "zb or zigbee or mixed or c5 or shard" 2>&1 | tail -4
for sec in 4.9 9.83; do
for o in 0 1; do
  echo "== zb_wb16 $sec s order=$o"; SNRX_ZB_ORDER=$o python tools/ab_front.py zb_wb16 $sec 2>&1 | tail -1
done
done
for o in 0 1; do
  echo "== mixed_wb56 4.9 s order=$o"; SNRX_ZB_ORDER=$o python tools/ab_front.py mixed_wb56 4.9 2>&1 | tail -1
done
