#!/bin/bash
# round-2 GPU call 16: persistent attention with a dynamic schedule: correctness, isolated timing, timeline, in-step A/B
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in v0_base p1_persist p2_nodefer p3_noprobe; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  if [ "${PIPESTATUS[0]}" != "0" ]; then echo "attn_bench failed or hung: skipping"; continue; fi
  timeout -s KILL 90 python tools/attn_bench.py 8 257 2>&1 | tail -1
  timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py tests/test_text_gpu.py -q -x -k "attention or text" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -8
done
export AACLIP_LIB_PATH=$PWD/tools/variants/tl_persist.so
echo "=== timeline"; timeout -s KILL 90 python tools/attn_timeline.py 2>&1 | tail -72
echo "=== power-capped steady state"
python tools/attn_power.py tools/variants/v0_base.so tools/variants/p1_persist.so tools/variants/p2_nodefer.so
echo "=== in-step"
for v in v0_base p1_persist v0_base p1_persist; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run16_attn_persistent.txt
