"""Prints a SHA-256 of the maps + scores of one seeded batch (B=8 @518): bit-level A/B of schedule-only changes
(AACLIP_MLP_CHUNKS, AACLIP_SERPENTINE, ...), which must not change a single bit.   python tools/maps_checksum.py [B]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import pipeline, weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = 518
pred = pipeline.AnomalyPredictor(weights.make_visual_state_dict(seed=0, img_size=S), weights.make_text_anchor(0), img_size=S,
                                 max_batch=B)
maps, scores = pred.predict_device(weights.make_images(0, B, S).cuda())
torch.cuda.synchronize()
h = hashlib.sha256(maps.cpu().numpy().tobytes() + scores.cpu().numpy().tobytes()).hexdigest()
print(f"maps+scores sha256 {h[:16]}  launches {pred.launches_per_batch}")
