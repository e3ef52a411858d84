#!/bin/bash
# round-2 GPU call 1: parity with PDL launches + next-tile residual prefetch, then A/B benches
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_run1_tests.txt
cat gpurun_out/r02_run1_tests.txt
{
echo "== default (PDL on, prefetch one tile ahead)"; python tools/bench_brief.py --steps 10 --warmup 3
echo "== AACLIP_PDL=0";  AACLIP_PDL=0 python tools/bench_brief.py --steps 10 --warmup 3
echo "== prefetch=1 (round-1 behaviour), PDL on"; AACLIP_LIB_PATH=$PWD/tools/variants/pf1.so python tools/bench_brief.py --steps 10 --warmup 3
echo "== default again"; python tools/bench_brief.py --steps 10 --warmup 3
} 2>&1 | tee gpurun_out/r02_run1_bench.txt
