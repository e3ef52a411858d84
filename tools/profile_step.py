"""Minimal driver for ncu: N steps of the hot path (batch 32 @518 by default) with nothing else in the process.
The LAST step is bracketed by cudaProfilerStart/Stop: run ncu with --profile-from-start off to capture exactly one
warm step (weight packing / LayerNorm folding and the warm-up steps stay outside).
    python tools/profile_step.py [steps] [batch] [img_size]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import pipeline, weights

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 518
sd = weights.make_visual_state_dict(seed=0, img_size=S)
pred = pipeline.AnomalyPredictor(sd, weights.make_text_anchor(0), img_size=S, max_batch=B)
imgs = weights.make_images(0, B, S).cuda()
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    maps, scores = pred.predict_device(imgs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", pred.launches_per_batch, float(maps.sum()), float(scores.sum()))
