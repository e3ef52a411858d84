"""Times the device input transform (aaclip_preprocess_u8) alone: 32 decoded RGB images -> [32,3,518,518] fp32.
    python tools/preprocess_bench.py [H] [W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import preprocess

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = int(sys.argv[2]) if len(sys.argv) > 2 else H
B, S = 32, 518
src = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
pp = preprocess.ImagePreprocessor(S)
dst = torch.empty((B, 3, S, S), device="cuda")
for i in range(2):
    pp(src[i % 2], out=dst)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(10):
    pp(src[i % 2], out=dst)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{H}x{W} -> {S}: {ms:.3f} ms per {B} images, {B * (H * W * 3 + 3 * S * S * 4) / ms / 1e6:.0f} GB/s algorithmic")
