"""ncu driver: a few launches of one GEMM shape.   python tools/gemm_only.py [N] [K] [epi] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib, engine

_lib.build()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
epi = int(sys.argv[3]) if len(sys.argv) > 3 else _lib.EPI_BIAS_BF16
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32
M = 1370 * B
a = (torch.randn(M, K, device="cuda") * 0.5).half()
w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
bias = torch.randn(N, device="cuda") * 0.1
o16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
o32 = torch.zeros(M, N, device="cuda")
for _ in range(4):
    engine.gemm16(a, w, epi, bias=bias, out_f32=o32, out_16=o16 if epi < 3 else None)
torch.cuda.synchronize()
print("ok")
