#!/bin/bash
# round-2 GPU call 37: residual stream as (hi, lo) 16-bit planes: full GPU suite, map error, in-step A/B against the
# fp32-stream library of the previous commit
mkdir -p gpurun_out
{
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/map_error.py 2>&1 | tail -4
export AACLIP_NO_REBUILD=1
for r in 1 2 3; do
  for v in m0_prev m1_hilo; do
    export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
    echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
  done
done
} 2>&1 | tee gpurun_out/r02_run37_hilo_stream.txt
