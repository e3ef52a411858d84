#!/bin/bash
# Build-variant A/B on the GPU box: full GPU test suite (parity against the golden maps) + in-step bench per flag set.
#   tools/variant_check.sh "" "-DAACLIP_GELU_TANH_FIT=1"
i=0
for flags in "$@"; do
  i=$((i+1))
  export AACLIP_LIB_PATH=/tmp/libaaclip_vc$i.so AACLIP_NVCC_EXTRA="$flags"
  echo "=== [$flags]"
  python -c "import sys; sys.path.insert(0, '.'); import aaclip_b200; from aaclip_b200 import _lib; _lib.build()" || continue
  timeout -s KILL 400 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py tests/test_host_surface_gpu.py -q -x 2>&1 | grep -E "^E  .*assert|passed|failed|^FAILED" | head -6
  timeout -s KILL 120 python tools/map_error.py 2>&1 | tail -4
  for r in 1 2; do timeout -s KILL 200 python tools/bench_brief.py 2>&1 | tail -2; done
done
