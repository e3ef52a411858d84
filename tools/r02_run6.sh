#!/bin/bash
# round-2 GPU call 6: query adapters inside aaclip_forward (f2) -- parity + cost of the IQM branch at B=32 @518
mkdir -p gpurun_out
python -m pytest tests/test_forward_gpu.py tests/test_host_surface_gpu.py tests/test_fp32_mode_gpu.py -x -q 2>&1 | tail -6
python tools/iqm_bench.py 32 768 2>&1 | tail -2 | tee gpurun_out/r02_run6_iqm.txt
python tools/iqm_bench.py 32 512 2>&1 | tail -1 | tee -a gpurun_out/r02_run6_iqm.txt
