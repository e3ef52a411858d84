#!/bin/bash
# round-2 GPU call 32: L2 eviction hints on the residual-epilogue stores (fp32 stream evict_first, 16-bit operand copy
# evict_last) and on the attention output (evict_last): tests, in-step A/B, DRAM bytes without ncu's cache flush
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
AACLIP_LIB_PATH=$PWD/tools/variants/h1_hints.so timeout -s KILL 600 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py tests/test_text_gpu.py -x -q 2>&1 | tail -3
for r in 1 2 3; do
  for v in h0_nohints h1_hints; do
    export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
    echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
  done
done
CMD="python tools/profile_step.py 3 32 518"
for v in h0_nohints h1_hints; do
  AACLIP_LIB_PATH=$PWD/tools/variants/$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    --profile-from-start off --csv --log-file gpurun_out/traffic_nocc_$v.csv $CMD > gpurun_out/ncu_nocc_$v.log 2>&1
  tail -1 gpurun_out/ncu_nocc_$v.log
done
} 2>&1 | tee gpurun_out/r02_run32_l2_hints.txt
