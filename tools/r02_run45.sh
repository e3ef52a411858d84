#!/bin/bash
# round-2 GPU call 45: where Q tile 0 releases Q tile 1 (stagger of 1/6, 1/4, 1/2 of a tile period): both warps of a
# sub-partition have two MIO-wait phases half a period apart, so a half-period stagger makes them wait together
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in g2_half g1_quarter g0_sixth; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"; timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  timeout -s KILL 120 python -m pytest tests/test_kernels_gpu.py -q -x -k attention 2>&1 | tail -1
done
python tools/attn_power.py tools/variants/g2_half.so tools/variants/g1_quarter.so tools/variants/g0_sixth.so
python tools/attn_power.py tools/variants/g2_half.so tools/variants/g1_quarter.so tools/variants/g0_sixth.so
} 2>&1 | tee gpurun_out/r02_run45_stagger_point.txt
