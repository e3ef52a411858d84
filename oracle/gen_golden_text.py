"""Generates tests/golden/text_*.npz by executing the UNMODIFIED reference text path (TEST INFRASTRUCTURE; build
container only):   python oracle/gen_golden_text.py      (needs /root/reference; ~1 min)

The reference's own tokenizer (model/tokenizer.py + its BPE vocabulary) turns the real prompt ensembles of
dataset/constants.py into token ids for a few MVTec / VisA classes; the reference AdaptedCLIP (random-init CLIP text
tower loaded with aa-clip-iqm_b200/weights.make_text_state_dict) encodes them with encode_text and builds the
[768,2] anchors with forward_utils.get_adapted_single_class_text_embedding. Stored: token ids, group sizes, the
reference sentence embeddings and anchors; plus the restatement-vs-reference differences.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import aaclip_oracle as orc  # noqa: E402
from oracle import ref_import  # noqa: E402

CLASSES = [("MVTec", "carpet"), ("MVTec", "pill"), ("VisA", "pcb1")]
SEED = 5


def main():
    ref_import.install()
    os.chdir(ref_import.REF_ROOT)
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    from model.adapter import AdaptedCLIP
    from model.clip import create_model
    from model.tokenizer import tokenize
    import forward_utils as ref_fu
    from dataset.constants import PROMPTS, REAL_NAMES

    torch.set_num_threads(os.cpu_count())
    sd = weights.make_text_state_dict(seed=SEED)
    clip = create_model("ViT-L-14-336", img_size=224, pretrained=None, force_image_size=224, device="cpu").eval()
    model = AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=512, iqm_num_layers=2, iqm_num_heads=8).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [k for k in missing if k.startswith(("clipmodel.transformer.", "clipmodel.token_embedding",
                                                    "clipmodel.positional_embedding", "clipmodel.ln_final",
                                                    "text_adapter."))], "synthetic state dict misses text tensors"
    tokens, sizes, embs, anchors = [], [], [], []
    with torch.no_grad():
        for ds, cls in CLASSES:
            real = REAL_NAMES[ds][cls]
            for states in (PROMPTS["prompt_normal"], PROMPTS["prompt_abnormal"]):
                sent = [t.format(s.format(real)) for s in states for t in PROMPTS["prompt_templates"]]
                tk = tokenize(sent)                      # int32 [n,77], the reference tokenizer
                tokens.append(tk)
                sizes.append(tk.shape[0])
                embs.append(model.encode_text(tk))       # the reference encoder
            anchors.append(ref_fu.get_adapted_single_class_text_embedding(model, ds, cls, "cpu"))
        tokens = torch.cat(tokens)
        embs = torch.cat(embs)
        anchors = torch.stack(anchors)
        o_emb = orc.encode_text(sd, tokens)
        o_anc, r = [], 0
        for c in range(len(CLASSES)):
            a, b = sizes[2 * c], sizes[2 * c + 1]
            o_anc.append(orc.text_anchor(o_emb[r:r + a], o_emb[r + a:r + a + b]))
            r += a + b
        o_anc = torch.stack(o_anc)
    d_emb = (o_emb - embs).abs().max().item()
    d_anc = (o_anc - anchors).abs().max().item()
    path = os.path.join(REPO, "tests", "golden", "text_3classes.npz")
    np.savez_compressed(path, tokens=tokens.numpy().astype(np.int32), sizes=np.array(sizes, dtype=np.int64),
                        emb=embs.numpy(), anchors=anchors.numpy(), seed=np.int64(SEED),
                        classes=np.array([f"{d}/{c}" for d, c in CLASSES]),
                        w_checksum=np.float64(sum(float(v.double().sum()) for v in sd.values())),
                        oracle_vs_ref_emb=np.float64(d_emb), oracle_vs_ref_anchor=np.float64(d_anc))
    print(f"wrote {path}: {tokens.shape[0]} sentences, emb |max| {embs.abs().max():.3f}, "
          f"oracle_vs_ref_emb={d_emb:.2e} oracle_vs_ref_anchor={d_anc:.2e}; first ids {tokens[0, :10].tolist()}")


if __name__ == "__main__":
    main()
