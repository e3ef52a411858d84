#!/bin/bash
# round-2 GPU call 10: evaluation driver end to end on a synthetic on-disk dataset; half-precision issue-rate ubench
mkdir -p gpurun_out
python -m pytest tests/test_eval_script_gpu.py -x -q 2>&1 | tail -15
cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates3 pipe_rates3.cu && ./pipe_rates3 | tee ../../gpurun_out/r02_ubench_pipe_rates3.txt
