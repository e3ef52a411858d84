#!/bin/bash
# round-2 GPU call 41: MLP in row chunks (c_fc and c_proj of one chunk back to back: the chunk's hidden tensor stays in
# L2): bit-level equality of the outputs, in-step throughput for 1 / 4 / 5 / 6 / 8 chunks
mkdir -p gpurun_out
{
for c in 1 5; do echo "chunks $c: $(AACLIP_MLP_CHUNKS=$c python tools/maps_checksum.py 32 2>&1 | tail -1)"; done
for r in 1 2; do
  for c in 1 4 5 6 8; do
    echo "--- AACLIP_MLP_CHUNKS=$c"; AACLIP_MLP_CHUNKS=$c timeout -s KILL 300 python tools/bench_brief.py --steps 20 --warmup 5 2>&1 | tail -2
  done
done
} 2>&1 | tee gpurun_out/r02_run41_mlp_chunks.txt
