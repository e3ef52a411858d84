"""Experiment: the batch split into `n_sub` sub-batches that run on separate CUDA streams (independent images), so that
one sub-batch's kernels fill the idle tail waves of the other's. Prints images/s for n_sub = 1, 2, 4.
    python tools/two_stream.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib, engine as _engine, pipeline, weights

_lib.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = 518
sd = weights.make_visual_state_dict(seed=0, img_size=S)
pred = pipeline.AnomalyPredictor(sd, weights.make_text_anchor(0), img_size=S, max_batch=B)
imgs = [weights.make_images(0, B, S, first_index=7919 * j).cuda() for j in range(2)]
eng = pred.engine


def run(n_sub, streams, wss, i):
    x = imgs[i % 2]
    if n_sub == 1:
        return [pred.predict_device(x)]
    outs = []
    cur = torch.cuda.current_stream()
    ev0 = torch.cuda.Event()
    ev0.record(cur)
    step = B // n_sub
    for k in range(n_sub):
        with torch.cuda.stream(streams[k]):
            streams[k].wait_event(ev0)
            eng._ws = wss[k]
            outs.append(pred.predict_device(x[k * step:(k + 1) * step]))
            ev = torch.cuda.Event()
            ev.record(streams[k])
        cur.wait_event(ev)
    return outs


ref = run(1, None, None, 0)[0]
ref = (ref[0].clone(), ref[1].clone())
for n_sub in (1, 2, 4, 1, 2):
    streams = [torch.cuda.Stream() for _ in range(n_sub)]
    wss = [torch.empty(eng.workspace_bytes(B // n_sub), dtype=torch.uint8, device="cuda") for _ in range(n_sub)]
    if n_sub == 1:
        eng._ws = None
    for i in range(4):
        outs = run(n_sub, streams, wss, i)
    torch.cuda.synchronize()
    maps = torch.cat([o[0] for o in run(n_sub, streams, wss, 0)])
    torch.cuda.synchronize()
    same = torch.equal(maps, ref[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 12
    for i in range(K):
        run(n_sub, streams, wss, i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"n_sub={n_sub}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} images/s   maps bit-equal to the single-stream run: {same}", flush=True)
