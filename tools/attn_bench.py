"""Attention kernel alone: timing (CUDA events) + max abs error against an fp32 torch reference on a few heads.
    python tools/attn_bench.py [B] [N]
AACLIP_LIB_PATH / AACLIP_NVCC_EXTRA select or build a kernel variant (see aa-clip-iqm_b200/_lib.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib, engine

_lib.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1370
torch.manual_seed(0)
qkv = (torch.randn(B, N, 3072, device="cuda") * 1.5).half()
out = engine.attention(qkv)
torch.cuda.synchronize()
q, k, v = qkv[:2].float().split(1024, dim=-1)
q = q.view(2, N, 16, 64).transpose(1, 2)
k = k.view(2, N, 16, 64).transpose(1, 2)
v = v.view(2, N, 16, 64).transpose(1, 2)
ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
err = (out[:2].float() - ref.transpose(1, 2).reshape(2, N, 1024)).abs().max().item()
for _ in range(3):
    engine.attention(qkv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    engine.attention(qkv)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"{os.environ.get('AACLIP_NVCC_EXTRA', ''):40s} attention B={B} N={N}: {ms:.4f} ms  "
      f"{4.0 * B * 16 * N * N * 64 / ms / 1e9:7.1f} TFLOP/s  max abs err {err:.2e}", flush=True)
