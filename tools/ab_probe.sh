#!/bin/bash
# where the channelizer's time goes: builds of libsnoutrx.so with a phase removed (results are WRONG by construction), timed alone
# build them with: nvcc ... -DSNRX_PROBE_NO_TAIL | -DSNRX_PROBE_NO_STAGE | -DSNRX_PROBE_SKIP_BACK -o snout_b200/lib/libsnoutrx_<v>.so snrx.cu
for v in base NO_TAIL NO_STAGE SKIP_BACK; do
  echo "== $v"; SNRX_LIB=$PWD/snout_b200/lib/libsnoutrx_$v.so python tools/ab_serial.py 2>&1 | tail -1
done
