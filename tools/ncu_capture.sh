#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one warm step, (2) --set full of the GEMM and attention kernels.
# Each ncu run is preceded by the same command without ncu (must exit 0). Usage: tools/ncu_capture.sh [tag]
TAG=${1:-cap}
mkdir -p gpurun_out
CMD="python tools/profile_step.py 3 32 518"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_pair -s 25 -c 5 -f \
    -o gpurun_out/prof_gemm_$TAG $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tcgen05 -s 5 -c 1 -f \
    -o gpurun_out/prof_attn_$TAG $CMD > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/plain.log gpurun_out/ncu_gemm.log gpurun_out/ncu_attn.log gpurun_out/ncu_launches.log
