#!/bin/bash
# round-2 GPU call 42: what-if (timing only, wrong results): the MLP hidden tensor aliased onto its first 9 472 / 4 864
# rows (78 / 40 MB: L2-resident), i.e. the upper bound of a fused c_fc -> c_proj schedule that keeps it in L2
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for r in 1 2 3; do
  for v in f0_base f1_alias9472 f2_alias4864; do
    export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
    echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
  done
done
} 2>&1 | tee gpurun_out/r02_run42_hid_alias_whatif.txt
