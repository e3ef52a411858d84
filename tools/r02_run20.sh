#!/bin/bash
# round-2 GPU call 20: what bounds the residual-epilogue GEMMs? what-if builds (timing only, wrong results) that drop the
# residual read / the 16-bit copy / the fp32 write, isolated and inside the step
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in r0_base r1_noresid r2_nox16 r3_noxstore r4_none r5_noprefetch; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 120 python tools/kernel_bench.py 2>&1 | grep -E "out  resid|proj resid|out  bias|adapter"
  timeout -s KILL 300 python tools/bench_brief.py --steps 10 --warmup 3 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run20_resid_whatif.txt
