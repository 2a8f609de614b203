#!/bin/bash
# usage: tools/sass_hist.sh <mangled-function-name>   -- SASS opcode histogram of one kernel of libsnoutrx.so
cuobjdump -sass -fun "$1" snout_b200/lib/libsnoutrx.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn | awk '{n+=$1; print} END {print n, "TOTAL"}'
