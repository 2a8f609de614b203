#!/bin/bash
# persistent multi-tile channelizer with staggered CTA starts (SNRX_PFB_STAGGER_NS per CTA slot of an SM)
for cfg in "0 0" "54 0" "54 400" "54 800" "27 400"; do
  set -- $cfg
  echo "== tiles $1 stagger $2 ns"; SNRX_PFB_TILES=$1 SNRX_PFB_STAGGER_NS=$2 python tools/ab_serial.py 2>&1 | tail -1 | cut -c1-150
done
