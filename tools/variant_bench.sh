#!/bin/bash
# Build-variant A/B on the GPU box, bench only (2 runs per flag set).  tools/variant_bench.sh "" "-DFOO=1"
i=0
for flags in "$@"; do
  i=$((i+1))
  export AACLIP_LIB_PATH=/tmp/libaaclip_vb$i.so AACLIP_NVCC_EXTRA="$flags"
  python -c "import sys; sys.path.insert(0, '.'); import aaclip_b200; from aaclip_b200 import _lib; _lib.build()" || continue
  for r in 1 2; do timeout -s KILL 200 python tools/bench_brief.py 2>&1 | tail -2; done
done
