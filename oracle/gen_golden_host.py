"""Generates the host-surface fixtures by executing the UNMODIFIED reference (TEST INFRASTRUCTURE; build container
only):   python oracle/gen_golden_host.py
  tests/golden/ref_state_dict_keys.json : key -> shape of the reference AdaptedCLIP.state_dict() (the contract our
                                          host mirror must reproduce so that reference checkpoints load unchanged)
  tests/golden/pos_embed_resize.npz     : reference resize_pos_embed (model/model.py:396-427) of a seeded 24x24 grid to
                                          37x37 (518 px) and 16x16 (224 px), column-subsampled
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle import ref_import  # noqa: E402


def main():
    ref_import.install()
    os.chdir(ref_import.REF_ROOT)
    from model.adapter import AdaptedCLIP
    from model.clip import create_model
    from model.model import resize_pos_embed

    clip = create_model("ViT-L-14-336", img_size=518, pretrained=None, force_image_size=518, device="cpu").eval()
    out = {}
    for relu in (False, True):
        model = AdaptedCLIP(clip_model=clip, relu=relu, iqm_hidden_size=512, iqm_num_layers=2, iqm_num_heads=8)
        keys = {k: list(v.shape) for k, v in model.state_dict().items()}    # every key, the IQM branch included
        out["relu" if relu else "norelu"] = keys
    with open(os.path.join(REPO, "tests", "golden", "ref_state_dict_keys.json"), "w") as f:
        json.dump(out, f)
    print("keys:", len(out["norelu"]), len(out["relu"]))

    g = torch.Generator().manual_seed(77)
    old = torch.randn(577, 1024, generator=g) * 0.03
    fix = {"seed": np.int64(77)}
    for S in (518, 224):
        gs = S // 14
        sd = {"visual.positional_embedding": old.clone()}
        stub = types.SimpleNamespace(visual=types.SimpleNamespace(grid_size=(gs, gs)))
        resize_pos_embed(sd, stub)
        new = sd["visual.positional_embedding"]
        assert new.shape == (gs * gs + 1, 1024)
        fix[f"pos_{S}"] = new[:, ::16].numpy().copy()
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "pos_embed_resize.npz"), **fix)
    print("pos embed fixtures written", {k: v.shape for k, v in fix.items() if k != "seed"})


if __name__ == "__main__":
    main()
