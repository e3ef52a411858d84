#!/bin/bash
# round-2 GPU call 8: final validation of the committed state: whole GPU suite (with prints), smoke, default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "err|passed|failed|saturated|IQM|Error" | tail -40 | tee gpurun_out/r02_run8_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_run8_smoke.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run8_bench.json 2> gpurun_out/r02_run8_bench.err
python tools/bench_brief.py --steps 20 --warmup 5 | tee gpurun_out/r02_run8_brief.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_run8_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches','clocks')}, d['e2e'], d['roofline'], d['config3_text_anchors'])
PY
