#!/bin/bash
L=$PWD/snout_b200/lib
for v in _PREV "" _PREV "" _PREV ""; do
  SNRX_LIB=$L/libsnoutrx$v.so python tools/ab_front.py zb_wb16 4.9 2>&1 | tail -1
done
for v in _PREV ""; do
  SNRX_LIB=$L/libsnoutrx$v.so python tools/ab_front.py mixed_wb56 4.9 2>&1 | tail -1
done
