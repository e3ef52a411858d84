"""ncu driver: a few launches of the attention kernel alone (B x 1370 x 16 heads)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
qkv = torch.randn(B, 1370, 3072, device="cuda").half()
for _ in range(4):
    out = engine.attention(qkv)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
