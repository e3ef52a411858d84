#!/bin/bash
# round-2 GPU call 31: per-launch DRAM bytes of one warm step WITHOUT ncu's cache flush between launches
# (--cache-control none: inter-kernel L2 reuse is what the serpentine row order is about), serpentine off / on
mkdir -p gpurun_out
CMD="python tools/profile_step.py 3 32 518"
for s in 0 1; do
  AACLIP_SERPENTINE=$s ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    --profile-from-start off --csv --log-file gpurun_out/traffic_nocc_serp$s.csv $CMD > gpurun_out/ncu_nocc_serp$s.log 2>&1
  tail -1 gpurun_out/ncu_nocc_serp$s.log
done
