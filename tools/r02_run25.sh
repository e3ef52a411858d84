#!/bin/bash
# round-2 GPU call 25: the persistent attention kernel as THE kernel: full GPU suite, map error, timeline, in-step bench
mkdir -p gpurun_out
{
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/map_error.py 2>&1 | tail -4
export AACLIP_NO_REBUILD=1
AACLIP_LIB_PATH=/root/repo/tools/variants/tl.so timeout -s KILL 90 python tools/attn_timeline.py 2>&1 | tail -70
for r in 1 2; do
  echo "--- new"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run25_persistent_default.txt
