#!/bin/bash
# round-2 GPU call 2: per-instruction stall profile of the residual-epilogue GEMM (out-proj K=1024, c_proj K=4096)
# gemm launches of one step: 0 patch | block i < 6: 1+5i qkv, out, fc, proj, adapter | block 6: 31 qkv, 32 out, 33 fc, 34 proj
mkdir -p gpurun_out
CMD="python tools/profile_step.py 3 32 518"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k gemm_pair_kernel -s 32 -c 3 -f \
    -o gpurun_out/prof_resid_r02a $CMD > gpurun_out/ncu_resid.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu_resid.log
ls -la gpurun_out/*.ncu-rep
