#!/bin/bash
# where the channelizer's time goes: builds of libsnoutrx.so with a phase removed (results are WRONG by construction), timed alone
for v in base NO_TAIL NO_STAGE; do
  echo "== $v"; SNRX_LIB=$PWD/snout_b200/lib/libsnoutrx_$v.so python tools/ab_serial.py 2>&1 | tail -1
done
