"""Generates tests/golden/train_step_224.npz by executing the UNMODIFIED reference training step (TEST
INFRASTRUCTURE; build container only):   python oracle/gen_golden_train.py

train.train_image_adapter (train.py:115-236) is run for ONE batch on the reference AdaptedCLIP (seeded synthetic
weights, eval mode as in train.py:333, SGD with lr 0 so the weights stay put); the gradients it leaves on the adapter
and IQM parameters and the logged loss are stored (sub-sampled). The CPU test replays the same step through this
repository's mirror (autograd route + forward_utils losses) and compares.
"""
import logging
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle import ref_import  # noqa: E402

GRAD_KEYS = ["image_adapter.layer_adapters.0.fc.0.weight", "image_adapter.layer_adapters.5.fc.0.weight",
             "image_adapter.seg_proj.0.fc.weight", "image_adapter.seg_proj.3.fc.weight",
             "image_adapter.det_proj.fc.weight", "class_query_mlp.0.weight", "class_query_mlp.2.bias",
             "query_adapters.1.fc.weight", "iqm.encoder.layer.0.crossattention.attention.key.weight",
             "iqm.encoder.layer.1.output_query.dense.weight", "iqm.layernorm.weight"]


def main():
    ref_import.install()
    os.chdir(ref_import.REF_ROOT)
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    from model.adapter import AdaptedCLIP
    from model.clip import create_model
    import train as ref_train

    S, B, seed, hidden = 224, 2, 31, 768
    torch.set_num_threads(os.cpu_count())
    sd = weights.make_visual_state_dict(seed=seed, img_size=S, relu=False)
    clip = create_model("ViT-L-14-336", img_size=S, pretrained=None, force_image_size=S, device="cpu").eval()
    model = AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=hidden, iqm_num_layers=2, iqm_num_heads=8).eval()
    model.load_state_dict(sd, strict=False)
    weights.fill_iqm_branch_(model, seed)
    images = weights.make_images(seed, B, S)
    g = torch.Generator().manual_seed(seed)
    mask = (torch.rand(B, 1, S, S, generator=g) > 0.8).float()
    mask[0] = 0
    label = torch.tensor([0, 1])
    anchors = {"carpet": weights.make_text_anchor(seed), "grid": weights.make_text_anchor(seed + 1)}
    loader = [{"image": images, "mask": mask, "label": label, "class_name": ["carpet", "grid"]}]
    params = [p for n, p in model.named_parameters() if n.startswith(("image_adapter.", "iqm.", "class_query_mlp.",
                                                                     "query_adapters.", "iqm_layer_norm."))]
    opt = torch.optim.SGD(params, lr=0.0)
    sched = torch.optim.lr_scheduler.ConstantLR(opt, factor=1.0, total_iters=1)
    ref_train.tqdm = lambda x: x
    log = logging.getLogger("gen_golden_train")
    losses = []
    log.info = lambda msg, *a: losses.append(msg) if str(msg).startswith("loss:") else None
    torch.set_num_threads(os.cpu_count())
    with tempfile.TemporaryDirectory() as tmp:
        ref_train.train_image_adapter(model=model, text_embeddings=anchors, train_loader=loader, optimizer=opt,
                                      scheduler=sched, device="cpu", start_epoch=0, save_path=tmp, image_epoch=1,
                                      img_size=S, logger=log)
    loss = float(str(losses[-1]).split(":")[1])
    named = dict(model.named_parameters())
    fix = dict(img_size=np.int64(S), batch=np.int64(B), seed=np.int64(seed), hidden=np.int64(hidden),
               loss=np.float64(loss), mask=mask.numpy().astype(np.uint8), label=label.numpy())
    for k in GRAD_KEYS:
        gr = named[k].grad
        assert gr is not None, k
        fix["grad:" + k] = (gr[:16, :64] if gr.dim() == 2 else gr[:64]).numpy().copy()
        fix["gnorm:" + k] = np.float64(gr.double().norm())
    path = os.path.join(REPO, "tests", "golden", "train_step_224.npz")
    np.savez_compressed(path, **fix)
    print("loss", loss, {k: float(v) for k, v in fix.items() if k.startswith("gnorm:")}, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
