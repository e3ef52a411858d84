#!/bin/bash
# round-2 GPU call 12: attention variants (trimmed last K/V tile, early mbarrier probes, S load ahead) A/B:
# isolated timing + error, attention tests, power-capped mJ/launch, then the in-step bench of base vs all.
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in v0_base v5_notrim_noall v2_trim v3_trim_probe v4_trim_ldahead v1_all; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  timeout -s KILL 90 python tools/attn_bench.py 8 257 2>&1 | tail -1
  timeout -s KILL 150 python -m pytest tests/test_kernels_gpu.py tests/test_text_gpu.py -q -x -k "attention or text" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -8
done
echo "=== power-capped steady state"
python tools/attn_power.py tools/variants/v0_base.so tools/variants/v2_trim.so tools/variants/v3_trim_probe.so tools/variants/v4_trim_ldahead.so tools/variants/v1_all.so
echo "=== in-step"
for v in v0_base v1_all v0_base v1_all; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "--- $v"; timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/r02_run12_attn_ab.txt
