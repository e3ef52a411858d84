#!/bin/bash
# round-2 GPU call 27: attention item order: natural (K/V of a head read once) vs one-tile items last (best tail balance)
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in s0_natural s1_singles_last; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py tests/test_text_gpu.py -q -x -k "attention or text" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -8
done
echo "=== power-capped steady state"
python tools/attn_power.py tools/variants/s0_natural.so tools/variants/s1_singles_last.so
echo "=== in-step (3 rounds, alternating)"
for v in s0_natural s1_singles_last s0_natural s1_singles_last s0_natural s1_singles_last; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run27_attn_item_order.txt
