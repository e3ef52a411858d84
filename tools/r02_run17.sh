#!/bin/bash
# round-2 GPU call 17: GEMM tail split A/B (tests + isolated GEMM shapes + in-step), attention LAZYMAX x POLY_EVERY combos
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in g0_nosplit g1_split; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -q -x 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -8
  timeout -s KILL 120 python tools/kernel_bench.py 2>&1 | tail -14
done
for v in a1_lazy_p3 a2_lazy_p2; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"; timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
done
echo "=== power-capped attention"
python tools/attn_power.py tools/variants/g1_split.so tools/variants/a1_lazy_p3.so tools/variants/a2_lazy_p2.so
echo "=== in-step"
for v in g0_nosplit g1_split g0_nosplit g1_split; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run17_gemm_tail_split.txt
