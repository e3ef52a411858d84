#!/bin/bash
# round-2 GPU call 5 (2 GPUs): bench.py at N=2 = the config-5 path (micro-batches, overlapped NCCL gather, dp parity)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_run5_bench_n2.json 2> gpurun_out/r02_run5_bench_n2.err
echo "rc=$?"; tail -c 2500 gpurun_out/r02_run5_bench_n2.json; tail -5 gpurun_out/r02_run5_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -c 600
