#!/bin/bash
# round-2 GPU call 44: L2 evict_first hint on the TMA loads of dead A operands (hidden tensor in c_proj, attention output
# in out_proj): bit equality, in-step A/B through AACLIP_DEAD_OPERAND_HINT=0/1, DRAM bytes without cache flush
mkdir -p gpurun_out
{
for c in 0 1; do echo "hint $c: $(AACLIP_DEAD_OPERAND_HINT=$c python tools/maps_checksum.py 8 2>&1 | tail -1)"; done
for r in 1 2 3; do
  for c in 0 1; do
    echo "--- AACLIP_DEAD_OPERAND_HINT=$c"; AACLIP_DEAD_OPERAND_HINT=$c timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
  done
done
CMD="python tools/profile_step.py 3 32 518"
for c in 0 1; do
  AACLIP_DEAD_OPERAND_HINT=$c ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    --profile-from-start off --csv --log-file gpurun_out/traffic_nocc_dead$c.csv $CMD > gpurun_out/ncu_nocc_dead$c.log 2>&1
  tail -1 gpurun_out/ncu_nocc_dead$c.log
done
} 2>&1 | tee gpurun_out/r02_run44_dead_operand_hint.txt
