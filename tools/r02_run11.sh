#!/bin/bash
# round-2 GPU call 11: CUDA-graph capture of the step: parity test + A/B bench (direct launches vs one graph per batch)
mkdir -p gpurun_out
python -m pytest tests/test_forward_gpu.py -x -q -k "graph or golden or streams" 2>&1 | tail -5
{
echo "== direct launches"; python tools/bench_brief.py --steps 20 --warmup 5
echo "== --graph";         python tools/bench_brief.py --steps 20 --warmup 5 --graph
echo "== direct launches"; python tools/bench_brief.py --steps 20 --warmup 5
echo "== --graph";         python tools/bench_brief.py --steps 20 --warmup 5 --graph
} 2>&1 | tee gpurun_out/r02_run11_graph_ab.txt
