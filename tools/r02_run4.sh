#!/bin/bash
# round-2 GPU call 4: the whole GPU suite (new: IQM golden loop, autograd route, tokenizer-driven anchors, saturation
# counter) + the default bench line (config 2, with config 3 inside) + smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_run4_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_run4_smoke.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_run4_bench.json 2> gpurun_out/r02_run4_bench.err
tail -c 3000 gpurun_out/r02_run4_bench.json; tail -5 gpurun_out/r02_run4_bench.err
