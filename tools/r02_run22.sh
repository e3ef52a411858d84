#!/bin/bash
# round-2 GPU call 22: persistent attention, second cut (per-item re-stagger, decoded items in the ring, one issue site)
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in q0_base q1_persist q2_persist_nostagger q3_persist_nodefer; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  if [ "${PIPESTATUS[0]}" != "0" ]; then echo "attn_bench failed or hung: skipping"; continue; fi
  timeout -s KILL 90 python tools/attn_bench.py 8 257 2>&1 | tail -1
  timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py tests/test_text_gpu.py -q -x -k "attention or text" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -8
done
export AACLIP_LIB_PATH=$PWD/tools/variants/tl_pp.so
echo "=== timeline"; timeout -s KILL 90 python tools/attn_timeline_p.py 2>&1 | tail -70
echo "=== power-capped steady state"
python tools/attn_power.py tools/variants/q0_base.so tools/variants/q1_persist.so tools/variants/q3_persist_nodefer.so
echo "=== in-step"
for v in q0_base q1_persist q0_base q1_persist; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run22_attn_persistent2.txt
