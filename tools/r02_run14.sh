#!/bin/bash
# round-2 GPU call 14/15: clock64 timeline of one attention CTA: softmax warps and MMA issuer warps
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in "$@"; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"; timeout -s KILL 90 python tools/attn_timeline.py 2>&1 | tail -56
done
} 2>&1 | tee gpurun_out/r02_run15_timeline.txt
