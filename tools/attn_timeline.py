"""Developer tool: clock64() timeline of CTA 5 of the attention kernel (softmax warps 0 / 4 and the two MMA
issuers over the first 16 tiles the CTA processes; one item boundary is inside). Build with
    AACLIP_NVCC_EXTRA=-DAACLIP_ATTN_DBG_TIMELINE AACLIP_LIB_PATH=/tmp/tl.so python tools/attn_timeline.py """
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib, engine

_lib.build()
lib = _lib.load()
qkv = (torch.randn(32, 1370, 3072, device="cuda") * 1.5).half()
for _ in range(3):
    engine.attention(qkv)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 16 * 8))()
mb = (ctypes.c_longlong * (2 * 16 * 8))()
assert lib.aaclip_debug_attn_timeline(buf, mb) == 0
names = ["loop", "s_full", "ld+free", "max", "exp0", "o_full(+write-back)", "exp1", "arrive"]
t0 = min(buf[(t * 16 + 0) * 8 + 0] for t in range(2))
for t in range(2):
    print(f"Q tile {t}: (cycles since the first stamp; then deltas between consecutive stamps)")
    prev = None
    for j in range(16):
        row = [buf[(t * 16 + j) * 8 + k] for k in range(8)]
        d = [row[0] - (prev if prev is not None else row[0])] + [row[k] - row[k - 1] for k in range(1, 8)]
        prev = row[7]
        print(f"  g={j:2d} start {row[0] - t0:7d}  " + "  ".join(f"{n} {x:5d}" for n, x in zip(names, d)) +
              f"   total {row[7] - row[0]:5d}")
for t in range(2):
    print(f"MMA issuer of Q tile {t}: S(g): k_full wait, s_free wait, issue;  PV(g): v_full+p_full wait, issue")
    for j in range(16):
        row = [mb[(t * 16 + j) * 8 + k] for k in range(7)]
        arrive = buf[(t * 16 + j) * 8 + 7]
        print(f"  g={j:2d} S start {row[0] - t0:7d}  k_full {row[1] - row[0]:5d}  s_free {row[2] - row[1]:5d}  issue {row[3] - row[2]:5d}"
              f"   PV start {row[4] - t0:7d}  wait {row[5] - row[4]:5d}  issue {row[6] - row[5]:5d}"
              f"   [arrive@{arrive - t0:7d}  seen +{row[5] - arrive:5d}  issued +{row[6] - arrive:5d}]")
