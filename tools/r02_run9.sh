#!/bin/bash
# round-2 GPU call 9 (8 GPUs): bench.py at N=8 = BASELINE config 5 as written (global batch 1024), dp parity on hardware
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_run9_bench_n$N.json 2> gpurun_out/r02_run9_bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r02_run9_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_run9_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','dtype','gpu_launches','clocks','dp_parity')}, d['e2e'], d['config']['workload'])
PY
