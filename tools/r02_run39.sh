#!/bin/bash
# round-2 GPU call 39: final evidence for the shipped library ((hi, lo) stream, serpentine order, persistent attention):
# full GPU suite, smoke, default + 20-step bench lines, isolated kernel timings, ncu captures (tag r02z)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02z_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; tail -c 300 gpurun_out/r02z_bench_n1.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02z_bench_n1_20steps.json 2>> gpurun_out/r02z_bench_n1.err
python tools/kernel_bench.py 2>&1 | tee gpurun_out/r02z_kernel_bench.txt | tail -3
bash tools/r02_profile.sh r02z
CMD="python tools/profile_step.py 3 32 518"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    --profile-from-start off --csv --log-file gpurun_out/traffic_nocc_r02z.csv $CMD > gpurun_out/ncu_nocc_r02z.log 2>&1
