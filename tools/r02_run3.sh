#!/bin/bash
# round-2 GPU call 3: pipelined residual epilogue (default) vs the previous form (rp0): parity, isolated kernels, step
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -x -q 2>&1 | tail -4
{
echo "== default (RESID_PIPE=1)"; python tools/kernel_bench.py 32 2>&1 | grep -E "resid|attention"
echo "== rp0"; AACLIP_LIB_PATH=$PWD/tools/variants/rp0.so python tools/kernel_bench.py 32 2>&1 | grep -E "resid|attention"
echo "== default"; python tools/bench_brief.py --steps 10 --warmup 3
echo "== rp0"; AACLIP_LIB_PATH=$PWD/tools/variants/rp0.so python tools/bench_brief.py --steps 10 --warmup 3
echo "== default"; python tools/bench_brief.py --steps 10 --warmup 3
} 2>&1 | tee gpurun_out/r02_run3_bench.txt
