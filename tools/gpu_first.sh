#!/bin/bash
# first GPU round trip: each kernel family in its own process under a timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for fam in gemm ln maps attn; do
  timeout 300 python tools/gpu_kernel_check.py $fam > gpurun_out/check_$fam.log 2>&1
  echo "exit $fam: $?" >> gpurun_out/check_summary.txt
done
cat gpurun_out/check_summary.txt
tail -n 40 gpurun_out/check_gemm.log
tail -n 25 gpurun_out/check_attn.log
