#!/usr/bin/env python
"""Single-image anomaly map (the reference's README.md:48-55 names a ``single_image_test.py`` that is not in its tree;
this one is written against the same call surface -- create_model / AdaptedCLIP / get_adapted_*_text_embedding /
calculate_similarity_map -- on the sm_100a engine).

    python single_image_test.py --image path/to.png --class_name carpet --dataset MVTec --save_path ckpt/baseline
    python single_image_test.py --synthetic            # no files needed: seeded weights, image and anchors

Writes <out>.npy (the [S,S] fp32 anomaly map, summed over the four levels) and prints the image score.
"""
import argparse
import glob
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import aaclip_b200  # noqa: E402,F401
from aaclip_b200 import forward_utils as fu  # noqa: E402
from aaclip_b200 import weights  # noqa: E402
from aaclip_b200.model import AdaptedCLIP, create_model  # noqa: E402

from aaclip_b200 import preprocess  # noqa: E402


def load_image(path: str, size: int, device: str) -> torch.Tensor:
    """dataset/__init__.py:86,150-159: decode to RGB on the host; bicubic resize to size x size, ToTensor and the CLIP
    mean/std normalisation run on the device (aaclip_preprocess_u8, bit-exact with PIL + torchvision)."""
    from PIL import Image
    rgb = np.asarray(Image.open(path).convert("RGB"))
    return preprocess.ImagePreprocessor(size, device=device)(rgb)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--image")
    ap.add_argument("--class_name", default="object")
    ap.add_argument("--dataset", default="MVTec")
    ap.add_argument("--img_size", type=int, default=518)
    ap.add_argument("--relu", action="store_true")
    ap.add_argument("--save_path", default="ckpt/baseline", help="directory holding text_adapter.pth / image_adapter_*.pth")
    ap.add_argument("--domain", default="Industrial", choices=["Industrial", "Medical"])
    ap.add_argument("--out", default="anomaly_map")
    ap.add_argument("--synthetic", action="store_true", help="seeded random weights / image / anchors (no files)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("single_image_test.py needs a CUDA sm_100a device (there is no CPU path)")
    S, dev = args.img_size, "cuda"
    if args.synthetic:
        clip = create_model("ViT-L-14-336", S, device=dev, pretrained=None, force_image_size=S)
        model = AdaptedCLIP(clip_model=clip, relu=args.relu).to(dev).eval()
        model.load_state_dict(weights.make_visual_state_dict(seed=0, img_size=S, relu=args.relu), strict=False)
        image = weights.make_images(0, 1, S)
        text = weights.make_text_anchor(0).to(dev)
    else:
        if not args.image:
            raise SystemExit("--image is required unless --synthetic")
        clip = create_model("ViT-L-14-336", S, device=dev, pretrained="openai", require_pretrained=True)
        model = AdaptedCLIP(clip_model=clip, relu=args.relu).to(dev).eval()
        tfile = glob.glob(os.path.join(args.save_path, "text_adapter.pth"))
        if tfile:
            model.text_adapter.load_state_dict(torch.load(tfile[0])["text_adapter"])
        files = sorted(glob.glob(os.path.join(args.save_path, "image_adapter_*.pth")),
                       key=lambda p: int(p.split("_")[-1].split(".")[0]))
        if not files:
            raise SystemExit(f"no image_adapter_*.pth under {args.save_path}")
        model.image_adapter.load_state_dict(torch.load(files[-1])["image_adapter"])
        image = load_image(args.image, S, dev)
        with torch.no_grad():
            text = fu.get_adapted_single_class_text_embedding(model, args.dataset, args.class_name, dev)
    with torch.no_grad():
        seg, det, _ = model(image.to(dev))
        amap, score = fu.anomaly_maps(seg, text, det, S, args.domain)
    np.save(args.out + ".npy", amap[0].cpu().numpy())
    print(f"anomaly map {tuple(amap.shape[1:])} -> {args.out}.npy  range [{float(amap.min()):.3f}, {float(amap.max()):.3f}]  "
          f"image score {float(score[0]):.4f}")


if __name__ == "__main__":
    main()
