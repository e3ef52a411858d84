#!/bin/bash
# round-2 GPU call 28: final evidence for the shipped library (persistent attention, natural item order)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; tail -c 300 gpurun_out/r02z_bench_n1.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02z_bench_n1_20steps.json 2>> gpurun_out/r02z_bench_n1.err
python tools/kernel_bench.py 2>&1 | tee gpurun_out/r02z_kernel_bench.txt | tail -3
bash tools/r02_profile.sh r02z
