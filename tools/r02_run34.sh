#!/bin/bash
# round-2 GPU call 34: final state: full GPU suite, default bench line, 20-step bench line
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02z_gpu_tests.txt
python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; tail -c 300 gpurun_out/r02z_bench_n1.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02z_bench_n1_20steps.json 2>> gpurun_out/r02z_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
