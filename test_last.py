#!/usr/bin/env python
"""Evaluation of the last adapter checkpoint on a dataset -- the role of the reference's ``test_last.py`` (same command
line, same per-class table of pixel / image AUROC and AP, same ``get_predictions`` return values), written against
this repository's engine:

    python test_last.py --dataset MVTec --save_path ckpt/baseline [--relu] [--img_size 518] [--image_batch_size 32]

What is different from the reference's script (test_last.py:53-330), none of it in the numbers it prints:
  * the per-image transform runs on the GPU (``--device_transform``, default): the loader workers only decode, the
    bicubic resize + ToTensor + Normalize is aaclip_preprocess_u8 (bit-exact with PIL + torchvision) when all images
    of a batch have one size; otherwise the loader falls back to the reference's CPU transform for that dataset;
  * the four text maps of a batch come from one fused call (level sum on the device) and the IQM branch's four
    projections from the tower's own GEMMs;
  * ``--device_metrics`` evaluates AUROC / AP on the GPU (metrics_eval_device) instead of scikit-learn on the host;
  * ``--well_posed_image_score`` replaces the fork's image-score broadcast (test_last.py:90-91 yields a [B,2] array that
    carries image #1's feature for every image) by ((det . text)[:, 1] + 1) / 2 per image. The default
    reproduces the reference's arithmetic, quirk included, so that its tables are comparable;
  * ``--synthetic_weights`` runs without any checkpoint file (seeded random weights; for smoke tests).
"""
import argparse
import logging
import os
import sys
from glob import glob

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from dataset import DOMAINS, get_dataset  # noqa: E402
from forward_utils import anomaly_maps, get_adapted_text_embedding, metrics_eval  # noqa: E402
from model.adapter import AdaptedCLIP  # noqa: E402
from model.clip import create_model  # noqa: E402
from utils import setup_seed  # noqa: E402


def get_predictions(model: nn.Module, class_text_embeddings: torch.Tensor, test_loader, device, img_size: int,
                    dataset: str = "MVTec", well_posed_image_score: bool = False, preprocessor=None):
    """-> (masks [n,1,S,S], labels [n], preds [n,1,S,S], preds_image, file_names): test_last.py:53-158.
    preds = 0.6 * sum_l text_map_l + 0.4 * sum_l iqm_map_l; preds_image is the fork's [n,2] array unless
    `well_posed_image_score` ([n])."""
    iqm_weight, text_weight = 0.4, 0.6
    masks, labels, preds, preds_image, file_names = [], [], [], [], []
    for batch in test_loader:
        assert len(set(batch["class_name"])) == 1, "mixed class not supported"
        if "image_u8" in batch:
            image = preprocessor(batch["image_u8"].to(device, non_blocking=True))
        else:
            image = batch["image"].to(device)
        masks.append(batch["mask"].cpu().numpy())
        labels.append(np.asarray(batch["label"]))
        file_names.extend(batch["file_name"])
        epoch_text_feature = class_text_embeddings.unsqueeze(0).repeat(image.size(0), 1, 1)
        patch_features, det_feature, iqm_outputs = model(image, text_embeddings=epoch_text_feature)
        text_map, score = anomaly_maps(patch_features, epoch_text_feature, det_feature, img_size, DOMAINS[dataset])
        if well_posed_image_score:
            preds_image.append(score.cpu().numpy())
        else:
            pred = det_feature @ epoch_text_feature              # [B,B,2]: the fork's broadcast (test_last.py:90)
            preds_image.append(((pred[:, 1] + 1) / 2).cpu().numpy())
        q = iqm_outputs.last_hidden_state
        norm_query, abnorm_query = q[:, 0, :], q[:, 1, :]
        iqm_map = 0
        for f in patch_features:
            if norm_query.shape[-1] != f.shape[-1]:              # test_last.py:111-118: a fresh random projection
                proj_layer = nn.Linear(norm_query.shape[-1], f.shape[-1]).to(f.device)
                norm_query, abnorm_query = proj_layer(norm_query), proj_layer(abnorm_query)
            diff = (F.cosine_similarity(f, abnorm_query.unsqueeze(1), dim=-1)
                    - F.cosine_similarity(f, norm_query.unsqueeze(1), dim=-1))
            B, L = diff.shape
            H = int(L ** 0.5)
            assert H * H == L, f"L={L} is not a perfect square"
            iqm_map = iqm_map + F.interpolate(torch.sigmoid(diff).view(B, 1, H, H), size=(img_size, img_size),
                                              mode="bilinear", align_corners=False)
        preds.append((text_map.unsqueeze(1) * text_weight + iqm_map * iqm_weight).cpu().numpy())
    return (np.concatenate(masks, axis=0), np.concatenate(labels, axis=0), np.concatenate(preds, axis=0),
            np.concatenate(preds_image, axis=0), file_names)


def build_parser():
    p = argparse.ArgumentParser(description="Evaluation of the last adapter checkpoint")
    p.add_argument("--model_name", type=str, default="ViT-L-14-336")
    p.add_argument("--img_size", type=int, default=518)
    p.add_argument("--relu", action="store_true")
    p.add_argument("--dataset", type=str, default="MVTec")
    p.add_argument("--shot", type=int, default=4)
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--image_batch_size", type=int, default=32)
    p.add_argument("--seed", type=int, default=111)
    p.add_argument("--save_path", type=str, default="ckpt/baseline")
    p.add_argument("--visualize", action="store_true")
    p.add_argument("--text_norm_weight", type=float, default=0.1)
    p.add_argument("--text_adapt_weight", type=float, default=0.1)
    p.add_argument("--image_adapt_weight", type=float, default=0.1)
    p.add_argument("--text_adapt_until", type=int, default=3)
    p.add_argument("--image_adapt_until", type=int, default=6)
    p.add_argument("--iqm_hidden_size", type=int, default=512)
    p.add_argument("--iqm_num_layers", type=int, default=2)
    p.add_argument("--iqm_num_heads", type=int, default=8)
    p.add_argument("--iqm_weight", type=float, default=0.7)
    # additions of this repository
    p.add_argument("--data_path", type=str, default=None, help="dataset root (default: dataset.constants.DATA_PATH)")
    p.add_argument("--metadata_root", type=str, default=None, help="directory holding <dataset>/full-shot.jsonl")
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--no_device_transform", dest="device_transform", action="store_false")
    p.add_argument("--device_metrics", action="store_true")
    p.add_argument("--well_posed_image_score", action="store_true")
    p.add_argument("--synthetic_weights", action="store_true")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.visualize:
        raise SystemExit("--visualize is not provided by this build")
    setup_seed(args.seed)
    os.makedirs(args.save_path, exist_ok=True)
    logger = logging.getLogger("test_last")
    logging.basicConfig(filename=os.path.join(args.save_path, "test.log"), encoding="utf-8", level=logging.INFO, force=True)
    logger.info("args: %s", vars(args))
    if not torch.cuda.is_available():
        raise SystemExit("test_last.py needs a CUDA sm_100a device (there is no CPU path)")
    device = torch.device("cuda:0")
    # ---- model (test_last.py:208-247)
    if args.synthetic_weights:
        from aaclip_b200 import weights
        clip_model = create_model(args.model_name, args.img_size, device=device, pretrained=None,
                                  force_image_size=args.img_size)
    else:
        clip_model = create_model(model_name=args.model_name, img_size=args.img_size, device=device, pretrained="openai",
                                  require_pretrained=True)
    clip_model.eval()
    model = AdaptedCLIP(clip_model=clip_model, text_adapt_weight=args.text_adapt_weight,
                        image_adapt_weight=args.image_adapt_weight, text_adapt_until=args.text_adapt_until,
                        image_adapt_until=args.image_adapt_until, relu=args.relu, iqm_hidden_size=args.iqm_hidden_size,
                        iqm_num_layers=args.iqm_num_layers, iqm_num_heads=args.iqm_num_heads,
                        max_batch=args.image_batch_size).to(device)
    model.eval()
    adapt_text = True
    if args.synthetic_weights:
        sd = dict(weights.make_visual_state_dict(seed=args.seed, img_size=args.img_size, relu=args.relu))
        sd.update(weights.make_text_state_dict(seed=args.seed))
        model.load_state_dict(sd, strict=False)
    else:
        text_file = glob(args.save_path + "/text_adapter.pth")
        if len(text_file) > 0:
            model.text_adapter.load_state_dict(torch.load(text_file[0])["text_adapter"])
        else:
            adapt_text = False
        files = glob(args.save_path + "/image_adapter_*.pth")
        assert len(files) > 0, "image adapter checkpoint not found"
        file = sorted(files, key=lambda x: int(x.split("_")[-1].split(".")[0]))[-1]      # the last epoch, numerically
        checkpoint = torch.load(file)
        model.image_adapter.load_state_dict(checkpoint["image_adapter"])
        logger.info("load model from epoch %d", checkpoint["epoch"])
    # ---- data (test_last.py:252-262)
    image_datasets = get_dataset(args.dataset, args.img_size, None, args.shot, "test", logger=logger,
                                 data_path=args.data_path, metadata_root=args.metadata_root,
                                 raw_uint8=args.device_transform)
    preprocessor = None
    if args.device_transform:
        from aaclip_b200 import preprocess
        preprocessor = preprocess.ImagePreprocessor(args.img_size, device=device)
    with torch.no_grad():
        text_embeddings = get_adapted_text_embedding(model if adapt_text else clip_model, args.dataset, device)
    # ---- per-class evaluation (test_last.py:273-321)
    from pandas import DataFrame, Series
    cols = ["class name", "pixel AUC", "pixel AP", "image AUC", "image AP"]
    df = DataFrame(columns=cols)
    for class_name, image_dataset in image_datasets.items():
        loader = torch.utils.data.DataLoader(image_dataset, batch_size=args.image_batch_size, shuffle=True,
                                             num_workers=args.num_workers, pin_memory=True)
        with torch.no_grad():
            masks, labels, preds, preds_image, _ = get_predictions(
                model=model, class_text_embeddings=text_embeddings[class_name], test_loader=loader, device=device,
                img_size=args.img_size, dataset=args.dataset, well_posed_image_score=args.well_posed_image_score,
                preprocessor=preprocessor)
        if args.device_metrics:
            from aaclip_b200.metrics import metrics_eval_device
            result = metrics_eval_device([torch.from_numpy(masks).to(device)], torch.from_numpy(labels).to(device),
                                         [torch.from_numpy(preds).to(device)], torch.from_numpy(preds_image).to(device),
                                         class_name, domain=DOMAINS[args.dataset])
        else:
            result = metrics_eval(masks, labels, preds, preds_image, class_name, domain=DOMAINS[args.dataset])
        df.loc[len(df)] = Series(result)
    average_row = df[cols[1:]].mean()
    average_row["class name"] = "Average"
    df.loc[len(df)] = average_row
    logger.info("final results:\n%s", df.to_string(index=False, justify="center"))
    print(df.to_string(index=False, justify="center"))
    return df


if __name__ == "__main__":
    main()
