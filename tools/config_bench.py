"""Timings of BASELINE.json configs 3 and 4 (the headline config 2 and config 5 are bench.py's).
  config 3: text-anchor build -- 27 classes x 16 prompts = 432 sentences [432,77] through the adapted text tower + the
            prompt-ensemble kernel (random-init weights, synthetic token ids with an end-of-text token per sentence).
  config 4: fp32-accuracy mode, ViT-L-14-336 @336, batch 128 (bench.py --dtype fp32).
    python tools/config_bench.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import engine, weights

n_cls, per = 27, 16
n = n_cls * per
g = torch.Generator().manual_seed(0)
tok = torch.zeros((n, 77), dtype=torch.int32)
for i in range(n):
    L = int(torch.randint(8, 20, (1,), generator=g))
    tok[i, 0] = 49406
    tok[i, 1:L] = torch.randint(300, 40000, (L - 1,), generator=g, dtype=torch.int32)
    tok[i, L] = 49407
tok = tok.cuda()
eng = engine.AaclipTextEngine("cuda", max_sentences=512)
eng.load_state_dict(weights.make_text_state_dict(seed=0))
sizes = [per // 2] * (2 * n_cls)
for _ in range(2):
    anchors = engine.text_anchors(eng.encode_text(tok), sizes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K):
    anchors = engine.text_anchors(eng.encode_text(tok), sizes)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"config": 3, "workload": f"{n} sentences x 77 tokens -> {n_cls} anchors [768,2]", "ms": ms,
                  "sentences_per_s": n / ms * 1e3, "tflops": n * 13.57e9 / (ms * 1e-3) / 1e12}), flush=True)
del eng
torch.cuda.empty_cache()
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--dtype", "fp32", "--img-size", "336", "--batch", "128",
                      "--steps", "2", "--warmup", "1", "--no-cpu-baseline"], capture_output=True, text=True)
if out.returncode == 0:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print(json.dumps({"config": 4, "workload": d["config"]["workload"], "images_per_s": d["value"], "ms_per_step": d["ms_per_step"],
                      "e2e_images_per_s": d["e2e"]["value"]}), flush=True)
else:
    print("config 4 failed:", out.stderr[-1500:])
