#!/bin/bash
# round-2 GPU call 19: ln_post folded into the token-tap projections (seg / det / query): full GPU suite + map error + in-step bench
mkdir -p gpurun_out
{
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/map_error.py 2>&1 | tail -4
export AACLIP_NO_REBUILD=1
for r in 1 2; do
  echo "--- new"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run19_lnpost_fold.txt
