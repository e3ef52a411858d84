#!/bin/bash
# What-if timing of the attention kernel with one resource's work removed (results are wrong on purpose).
#   tools/attn_whatif.sh "<flags>" "<flags>" ...
i=0
for flags in "$@"; do
  i=$((i+1))
  export AACLIP_LIB_PATH=/tmp/libaaclip_wi$i.so AACLIP_NVCC_EXTRA="$flags"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
done
