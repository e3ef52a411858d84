#!/bin/bash
# round-2 GPU call 43 (8 GPUs): BASELINE config 5 with the final library: global batch 1024 per step (128 per GPU in
# micro-batches of 32), gather overlapped, gathered rows of a foreign rank re-computed on rank 0
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02z_bench_n8_config5.json 2> gpurun_out/r02z_bench_n8.err
tail -c 600 gpurun_out/r02z_bench_n8_config5.json; tail -2 gpurun_out/r02z_bench_n8.err
