#!/bin/bash
# Attention variant sweep on the GPU box: isolated timing + attention tests + in-step bench for each flag set.
#   tools/attn_sweep.sh "-DAACLIP_ATTN_QT=2" "-DAACLIP_ATTN_QT=1" ...
# A variant that hangs (protocol bug) is cut after 90 s and the rest of its steps are skipped.
i=0
for flags in "$@"; do
  i=$((i+1))
  export AACLIP_LIB_PATH=/tmp/libaaclip_attn$i.so AACLIP_NVCC_EXTRA="$flags"
  echo "=== [$flags]"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  if [ "${PIPESTATUS[0]}" != "0" ]; then echo "attn_bench failed or hung: skipping"; continue; fi
  timeout -s KILL 120 python -m pytest tests/test_kernels_gpu.py -q -x -k attention 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -12
  timeout -s KILL 200 python tools/bench_brief.py 2>&1 | tail -2
done
