#!/bin/bash
# round-2 GPU call 46: one more default bench line + 20-step line of the final library on another box (box-to-box spread)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02z_bench_n1_rerun.json 2> gpurun_out/r02z_bench_n1_rerun.err; tail -c 200 gpurun_out/r02z_bench_n1_rerun.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02z_bench_n1_20steps_rerun.json 2>> gpurun_out/r02z_bench_n1_rerun.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 400
