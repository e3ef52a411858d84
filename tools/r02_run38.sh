#!/bin/bash
# round-2 GPU call 38: (hi, lo) residual stream with the mixed-precision FHADD / FHFMA instructions in the epilogue:
# kernel + forward tests, in-step A/B against the fp32-stream library and the first (hi, lo) cut
mkdir -p gpurun_out
{
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_text_gpu.py -q 2>&1 | tail -4
python tools/map_error.py 2>&1 | tail -4
export AACLIP_NO_REBUILD=1
for r in 1 2 3; do
  for v in m0_prev m1_hilo m2_hilo_mixed; do
    export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
    echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
  done
done
} 2>&1 | tee gpurun_out/r02_run38_hilo_mixed.txt
