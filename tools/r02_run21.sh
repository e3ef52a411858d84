#!/bin/bash
# round-2 GPU call 21: evidence for the final library of the round: GPU suite, default bench line, 20-step bench line,
# ncu launch list + traffic + --set full captures (tools/r02_profile.sh, tag r02b)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02b_gpu_tests.txt
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; tail -c 600 gpurun_out/r02b_bench_n1.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_n1_20steps.json 2>> gpurun_out/r02b_bench_n1.err
bash tools/r02_profile.sh r02b
