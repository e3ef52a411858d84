#!/bin/bash
# ncu evidence of the round-2 kernels for profiles/ (each ncu run is preceded by the same command without ncu):
#  (1) per-launch time + DRAM bytes of one warm step, (2) --set full of block 6's four GEMMs, (3) --set full of one
#  attention launch, (4) the memory-bound kernels.  Usage: tools/r02_profile.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python tools/profile_step.py 3 32 518"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/traffic_$TAG.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k gemm_pair_kernel -s 31 -c 4 -f \
    -o gpurun_out/prof_gemm_$TAG $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tcgen05 -s 5 -c 1 -f \
    -o gpurun_out/prof_attn_$TAG $CMD > gpurun_out/ncu_attn.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off \
    -k regex:'adapter_blend|layernorm_kernel|l2norm_rows|similarity_kernel|mean_over_patches|patch_gather|anomaly_map' -c 12 -f \
    -o gpurun_out/prof_mem_$TAG $CMD > gpurun_out/ncu_mem.log 2>&1
tail -n 2 gpurun_out/plain.log gpurun_out/ncu_traffic.log gpurun_out/ncu_gemm.log gpurun_out/ncu_attn.log gpurun_out/ncu_mem.log
ls -la gpurun_out/*_$TAG*
