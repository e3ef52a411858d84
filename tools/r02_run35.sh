#!/bin/bash
# round-2 GPU call 35: knob sweep of the persistent attention kernel: isolated ms and power-capped mJ per launch
mkdir -p gpurun_out
export AACLIP_NO_REBUILD=1
{
for v in k0_base k1_poly3 k2_poly5 k3_poly8 k6_poly0 k4_kv4 k5_regs40; do
  export AACLIP_LIB_PATH=$PWD/tools/variants/$v.so
  echo "=== $v"; timeout -s KILL 90 python tools/attn_bench.py 32 2>&1 | tail -1
done
echo "=== power-capped steady state (two passes)"
python tools/attn_power.py tools/variants/k0_base.so tools/variants/k1_poly3.so tools/variants/k2_poly5.so tools/variants/k3_poly8.so tools/variants/k6_poly0.so tools/variants/k4_kv4.so tools/variants/k5_regs40.so
python tools/attn_power.py tools/variants/k0_base.so tools/variants/k1_poly3.so tools/variants/k2_poly5.so tools/variants/k3_poly8.so tools/variants/k6_poly0.so tools/variants/k4_kv4.so tools/variants/k5_regs40.so
} 2>&1 | tee gpurun_out/r02_run35_attn_knobs.txt
