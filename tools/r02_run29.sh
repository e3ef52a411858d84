#!/bin/bash
# round-2 GPU call 29 (2 GPUs): config-5 path with the final library: 128 images per GPU per step in micro-batches of 32,
# gather overlapped, gathered rows re-computed on rank 0 (dp_parity); plus the two-device test that is skipped on one GPU
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02z_bench_n2_config5.json 2> gpurun_out/r02z_bench_n2.err
tail -c 1500 gpurun_out/r02z_bench_n2_config5.json; tail -3 gpurun_out/r02z_bench_n2.err
python -m pytest tests/test_forward_gpu.py -q -k "two" 2>&1 | tail -3
