#!/bin/bash
# round-2 GPU call 24: persistent attention with the O tile written by one TMA store per warp
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in q0_base p1_persist_stg p3_persist_tma; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  if [ "${PIPESTATUS[0]}" != "0" ]; then echo "attn_bench failed or hung: skipping"; continue; fi
  timeout -s KILL 90 python tools/attn_bench.py 8 257 2>&1 | tail -1
  timeout -s KILL 90 python tools/attn_bench.py 3 77 2>&1 | tail -1
  timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py tests/test_text_gpu.py -q -x -k "attention or text" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -8
done
echo "=== power-capped steady state"
python tools/attn_power.py tools/variants/q0_base.so tools/variants/p1_persist_stg.so tools/variants/p3_persist_tma.so
echo "=== in-step (3 rounds, alternating)"
for v in q0_base p3_persist_tma q0_base p3_persist_tma q0_base p3_persist_tma; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run24_attn_persistent_tma.txt
