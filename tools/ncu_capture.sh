#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one warm step, (2) --set full of the GEMM and attention kernels.
set -x
mkdir -p gpurun_out
CMD="python tools/profile_step.py 2 32 518"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 205 -c 205 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 108 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 24 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu_gemm.log gpurun_out/ncu_attn.log gpurun_out/ncu_launches.log
