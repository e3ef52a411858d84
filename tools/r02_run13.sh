#!/bin/bash
# round-2 GPU call 13: attention variants with ONE tile body for every width (blocks of 32 columns skipped at run time)
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in v0_base w5_none w6_probe_only w2_trim w3_trim_probe w4_trim_ldahead w1_all; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"
  timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
  timeout -s KILL 90 python tools/attn_bench.py 8 257 2>&1 | tail -1
  timeout -s KILL 150 python -m pytest tests/test_kernels_gpu.py tests/test_text_gpu.py -q -x -k "attention or text" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -8
done
echo "=== power-capped steady state"
python tools/attn_power.py tools/variants/v0_base.so tools/variants/w5_none.so tools/variants/w2_trim.so tools/variants/w1_all.so
} 2>&1 | tee gpurun_out/r02_run13_attn_ab.txt
