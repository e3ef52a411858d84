#!/bin/bash
# round-2 GPU call 33: images/s as a function of the batch per forward call (8 / 16 / 32 / 64): is a batch whose
# activations fit L2 better worth the shorter GEMM schedules?
mkdir -p gpurun_out
{
for b in 32 16 8 64 32 16; do
  echo "--- batch $b"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 --batch $b 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run33_batch_sweep.txt
