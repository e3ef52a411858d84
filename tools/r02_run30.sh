#!/bin/bash
# round-2 GPU call 30: serpentine row order (every GEMM / attention launch walks the rows opposite to its producer):
# parity tests, in-step A/B through AACLIP_SERPENTINE=0/1, per-launch DRAM bytes of one step with and without
mkdir -p gpurun_out
{
timeout -s KILL 600 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py tests/test_text_gpu.py -x -q 2>&1 | tail -3
for r in 1 2 3; do
  echo "--- AACLIP_SERPENTINE=0"; AACLIP_SERPENTINE=0 timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
  echo "--- AACLIP_SERPENTINE=1"; AACLIP_SERPENTINE=1 timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
CMD="python tools/profile_step.py 3 32 518"
for s in 0 1; do
  AACLIP_SERPENTINE=$s ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/traffic_serp$s.csv $CMD > gpurun_out/ncu_serp$s.log 2>&1
  tail -1 gpurun_out/ncu_serp$s.log
done
} 2>&1 | tee gpurun_out/r02_run30_serpentine.txt
