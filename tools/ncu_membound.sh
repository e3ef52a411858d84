#!/bin/bash
# ncu pass over the memory-bound kernels of one warm step (everything except the GEMM and attention kernels) plus the
# two pre-processing kernels: duration and DRAM bytes per launch -> achieved GB/s for DESIGN.md section 3.3.
TAG=${1:-cap}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
CMD="python tools/profile_step.py 3 32 518"
$CMD > gpurun_out/plain_mb.log 2>&1 && \
ncu --metrics $M --clock-control none --profile-from-start off --csv \
    -k regex:'layernorm|adapter_blend|l2norm|similarity|anomaly_map|mean_over|patch_gather|cls_row|image_score' \
    --log-file gpurun_out/membound_$TAG.csv $CMD > gpurun_out/ncu_mb.log 2>&1
PRE="python tools/preprocess_bench.py"
$PRE > gpurun_out/plain_pre.log 2>&1 && \
ncu --metrics $M --clock-control none --csv -k regex:resample -c 4 --log-file gpurun_out/membound_pre_$TAG.csv $PRE > gpurun_out/ncu_pre.log 2>&1
tail -n 2 gpurun_out/plain_mb.log gpurun_out/plain_pre.log
