"""Times the reference-facing call of test_last.get_predictions -- model(image, text_embeddings=...) with the IQM branch
-- at batch 32 @518: the default route (query adapters inside aaclip_forward + factored cross-attention) against the
literal route (fp32 tokens exported, query adapters / memory projection / key-value projections in host PyTorch).
    python tools/iqm_bench.py [B] [hidden]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import weights
from aaclip_b200.model import AdaptedCLIP, create_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 768
S = 518
clip = create_model("ViT-L-14-336", S, pretrained=None, force_image_size=S, device="cuda")
model = AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=hidden, max_batch=B).to("cuda").eval()
model.load_state_dict(weights.make_visual_state_dict(seed=0, img_size=S), strict=False)
weights.fill_iqm_branch_(model, 0)
x = weights.make_images(0, B, S).cuda()
text = weights.make_text_anchor(0).cuda().unsqueeze(0).repeat(B, 1, 1)


def timeit(fn, n=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


with torch.no_grad():
    t_plain, _ = timeit(lambda: model(x))
    t_fast, o_fast = timeit(lambda: model(x, text_embeddings=text))
    model.iqm_factored = False
    dim = model._engine_query_dim
    model._engine_query_dim = lambda: 0          # literal route: tokens exported, everything of the branch in PyTorch
    model.refresh_engines()
    t_lit, o_lit = timeit(lambda: model(x, text_embeddings=text))
err = (o_fast[2].last_hidden_state - o_lit[2].last_hidden_state).abs().max().item()
print(f"B={B} @518 hidden={hidden}: tower only {t_plain:.1f} ms | with IQM branch: device query adapters + factored "
      f"attention {t_fast:.1f} ms (+{t_fast - t_plain:.1f}) | literal host branch {t_lit:.1f} ms (+{t_lit - t_plain:.1f}) | "
      f"last_hidden_state max abs diff between the two {err:.2e}")
