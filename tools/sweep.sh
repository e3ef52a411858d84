#!/bin/bash
# Kernel-variant sweep on the GPU box: each argument is a set of -D flags; builds to /tmp and times with $BENCH.
#   BENCH="python tools/attn_bench.py 32" tools/sweep.sh "-DAACLIP_ATTN_POLY_EVERY=0" "-DAACLIP_ATTN_POLY_EVERY=2"
i=0
for flags in "$@"; do
  i=$((i+1))
  AACLIP_LIB_PATH=/tmp/libaaclip_var$i.so AACLIP_NVCC_EXTRA="$flags" timeout 300 $BENCH 2>&1 | tail -${TAIL:-1}
done
