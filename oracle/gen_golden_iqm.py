"""Generates tests/golden/iqm_*.npz by executing the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only):

    python oracle/gen_golden_iqm.py

The reference AdaptedCLIP (model/adapter.py) is built at 224 px with the deterministic visual weights of
aa-clip-iqm_b200/weights.py and deterministic IQM-branch weights (weights.fill_iqm_branch_), put in eval mode, and the
body of the reference's evaluation loop is executed on a seeded batch: test_last.get_predictions (test_last.py:53-158)
through a two-item loader. Stored: the loop's outputs (fused 0.6 / 0.4 maps, the [B,2] image predictions of the fork),
the IQM branch's last_hidden_state, and the IQM branch applied to SEEDED tokens / CLS rows (a CPU-checkable
fixture for the host-side module alone, no tower involved).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle import ref_import  # noqa: E402

CASES = [("iqm_224_h768", 224, 2, 21, 768), ("iqm_224_h512", 224, 2, 22, 512)]


def main():
    ref_import.install()
    os.chdir(ref_import.REF_ROOT)
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    from model.adapter import AdaptedCLIP
    from model.clip import create_model
    import test_last as ref_tl

    torch.set_num_threads(os.cpu_count())
    for name, S, B, seed, hidden in CASES:
        sd = weights.make_visual_state_dict(seed=seed, img_size=S, relu=False)
        clip = create_model("ViT-L-14-336", img_size=S, pretrained=None, force_image_size=S, device="cpu").eval()
        model = AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=hidden, iqm_num_layers=2, iqm_num_heads=8).eval()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected
        n = weights.fill_iqm_branch_(model, seed)
        images = weights.make_images(seed, B, S)
        text = weights.make_text_anchor(seed)                           # [768,2] = text_embeddings[class_name]
        loader = [{"image": images, "mask": torch.zeros(B, 1, S, S), "label": torch.tensor([0, 1][:B]),
                   "file_name": [f"f{i}" for i in range(B)], "class_name": ["carpet"] * B}]
        ref_tl.tqdm = lambda x: x
        with torch.no_grad():
            torch.manual_seed(1234)                # h512: get_predictions draws a random proj_layer per batch
            masks, labels, preds, preds_image, files = ref_tl.get_predictions(
                model=model, class_text_embeddings=text, test_loader=loader, device="cpu", img_size=S, dataset="MVTec")
            seg, det, iqm = model(images, text_embeddings=text.unsqueeze(0).repeat(B, 1, 1))
            # the branch alone on seeded inputs (host module parity, CPU-checkable)
            g = torch.Generator().manual_seed(seed + 100)
            P = (S // 14) ** 2
            tok = [torch.randn(B, P, 1024, generator=g) for _ in range(4)]
            cls = torch.randn(B, 1024, generator=g)
            q = model.class_query_mlp(cls).unsqueeze(1).repeat(1, 2, 1) + model.pos_embedding[:, :2, :].repeat(B, 1, 1)
            vis = model.visual_feature_proj(torch.cat([model.query_adapters[i](t) for i, t in enumerate(tok)], dim=1))
            te = model.text_feature_proj(text.unsqueeze(0).repeat(B, 1, 1))
            alone = model.iqm(query_embeds=q, query_length=2, encoder_hidden_states=vis, text_encoder_hidden_states=te)
            alone = model.iqm_layer_norm(alone.last_hidden_state)
        fix = dict(img_size=np.int64(S), batch=np.int64(B), seed=np.int64(seed), hidden=np.int64(hidden),
                   n_iqm_params=np.int64(n), preds=preds, preds_image=preds_image,
                   iqm_last_hidden=iqm.last_hidden_state.numpy(), branch_alone=alone.numpy(),
                   det=det.numpy(), text=text.numpy(),
                   w_checksum=np.float64(sum(float(p.double().sum()) for p in model.iqm.parameters())))
        path = os.path.join(REPO, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **fix)
        print(name, "preds", preds.shape, float(preds.min()), float(preds.max()), "preds_image", preds_image.shape,
              "iqm", iqm.last_hidden_state.shape, os.path.getsize(path) / 1e6, "MB", flush=True)


if __name__ == "__main__":
    main()
