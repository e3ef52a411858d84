"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only).

    python oracle/gen_golden.py            # needs /root/reference; ~1 min on 8 cores

For each case the reference AdaptedCLIP (model/adapter.py) is built at the case's image size, loaded with the
deterministic synthetic weights of aa-clip-iqm_b200/weights.py (same seed => same tensors on the GPU box), run on the
deterministic synthetic images, and its outputs are stored: patch features (row subsample), det token, CLS / tapped
tokens (subsample), the reference's own calculate_similarity_map outputs (test mode per domain, summed over levels as
test_last.py:149 does; train mode for level 0) and image scores. The same run checks oracle/aaclip_oracle.py against
the reference and records the max abs differences in the fixture (`oracle_vs_ref_*`).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import aaclip_oracle as orc  # noqa: E402
from oracle import ref_import  # noqa: E402

CASES = [
    # name, img_size, batch, seed, relu
    ("vitl_224_b2", 224, 2, 1, False),
    ("vitl_518_b1", 518, 1, 2, False),
    ("vitl_224_b1_relu", 224, 1, 3, True),
    ("vitl_336_b2", 336, 2, 4, False),      # BASELINE config 4 (fp32-accuracy mode @336)
    ("vitl_518_b2", 518, 2, 5, False),      # a second and third reference-pinned image at the headline size
]


def main(only=None):
    ref_import.install()
    os.chdir(ref_import.REF_ROOT)  # the reference uses ./dataset/... relative paths at import time
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    from model.adapter import AdaptedCLIP
    from model.clip import create_model
    import forward_utils as ref_fu

    torch.set_num_threads(os.cpu_count())
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, S, B, seed, relu in CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        sd = weights.make_visual_state_dict(seed=seed, img_size=S, relu=relu)
        clip = create_model("ViT-L-14-336", img_size=S, pretrained=None, force_image_size=S, device="cpu").eval()
        model = AdaptedCLIP(clip_model=clip, relu=relu, iqm_hidden_size=512, iqm_num_layers=2, iqm_num_heads=8).eval()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert not [k for k in missing if k.startswith(("image_encoder.", "image_adapter."))
                    and k != "image_encoder.proj"], "synthetic state dict does not cover the visual path"
        images = weights.make_images(seed, B, S)
        text = weights.make_text_anchor(seed)
        G = S // 14
        P = G * G
        with torch.no_grad():
            seg, det, iqm = model(images)  # text_embeddings=None -> no IQM branch (model/adapter.py:186-188)
            assert iqm is None
            maps_ind = torch.cat([ref_fu.calculate_similarity_map(f, text, S, test=True, domain="Industrial")
                                  for f in seg], dim=1).sum(1)
            maps_med = torch.cat([ref_fu.calculate_similarity_map(f, text, S, test=True, domain="Medical")
                                  for f in seg], dim=1).sum(1)
            train0 = ref_fu.calculate_similarity_map(seg[0], text, S, test=False)
            score = ((det @ text)[:, 1] + 1) / 2
            # the restatement, same inputs
            o = orc.predict(sd, images, text, S, domain="Industrial")
            o_med = orc.anomaly_maps(o["seg"], text, S, "Medical")
            o_train0 = orc.calculate_similarity_map(o["seg"][0], text, S, test=False)
        diffs = dict(
            oracle_vs_ref_seg=max((a - b).abs().max().item() for a, b in zip(o["seg"], seg)),
            oracle_vs_ref_det=(o["det"] - det).abs().max().item(),
            oracle_vs_ref_maps=(o["maps"] - maps_ind).abs().max().item(),
            oracle_vs_ref_maps_med=(o_med - maps_med).abs().max().item(),
            oracle_vs_ref_train0=(o_train0 - train0).abs().max().item(),
            oracle_vs_ref_score=(o["score"] - score).abs().max().item(),
        )
        rows = np.arange(0, P, max(1, P // 32))[:40]
        fix = dict(
            img_size=np.int64(S), batch=np.int64(B), seed=np.int64(seed), relu=np.int64(relu),
            rows=rows.astype(np.int64),
            seg_rows=np.stack([s[:, rows, :].numpy() for s in seg]),          # [4,B,len(rows),768]
            det=det.numpy(), score=score.numpy(),
            maps_industrial=maps_ind.numpy(), maps_medical=maps_med.numpy()[:, ::2, ::2].copy(),
            train0=train0.numpy()[:, :, ::4, ::4].copy(),
            cls=o["cls"].numpy(),  # reference exposes x[:,0] only inside the IQM branch; oracle value, see diffs
            w_checksum=np.float64(sum(float(v.double().sum()) for v in sd.values())),
            img_checksum=np.float64(float(images.double().sum())),
            text=text.numpy(),
            **{k: np.float64(v) for k, v in diffs.items()},
        )
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **fix)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB) in {time.time() - t0:.1f}s; "
              f"maps range [{maps_ind.min():.3f},{maps_ind.max():.3f}]; " +
              " ".join(f"{k}={v:.2e}" for k, v in diffs.items()), flush=True)
        del model, clip, sd


if __name__ == "__main__":
    main(sys.argv[1:] or None)
