"""Mirror of the reference's scoring / evaluation glue (forward_utils.py) for the hot path: same function names,
arguments and return shapes; the device work goes through the C ABI (aaclip_score_maps, aaclip_encode_text,
aaclip_text_anchors). ``metrics_eval`` keeps the reference's host semantics (rows a11: numpy + scikit-learn)."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from . import engine as _engine

# dataset/constants.py:137-147 -- the prompt ensemble that DEFINES the text anchors (data, not code)
PROMPTS = {
    "prompt_normal": ["{}", "a {}", "the {}"],
    "prompt_abnormal": ["a damaged {}", "a broken {}", "a {} with flaw", "a {} with defect", "a {} with damage"],
    "prompt_templates": ["{}.", "a photo of {}."],
}
DOMAIN_BLUR = {"Industrial": (7, 1.0), "Medical": (9, 1.5)}       # forward_utils.py:205-206


def calculate_similarity_map(patch_features: torch.Tensor, epoch_text_feature: torch.Tensor, img_size: int,
                             test: bool = False, domain: str = "Medical") -> torch.Tensor:
    """forward_utils.py:196-216. patch_features [B,P,768] unit-norm; epoch_text_feature [768,2] or [B,768,2].
    test=True -> [B,1,S,S] = upsample(blur((s1+1-s0)/2)); test=False -> [B,2,S,S] = softmax(upsample(s))."""
    ks, sg = DOMAIN_BLUR["Industrial" if domain == "Industrial" else "Medical"]
    if epoch_text_feature.shape[-1] != 2:
        raise AssertionError("calculate_similarity_map expects two anchors (normal, abnormal)")   # :204 `assert C == 2`
    maps, _ = _engine.score_maps([patch_features], epoch_text_feature, img_size, test=test, ksize=ks, sigma=sg)
    return maps.unsqueeze(1) if test else maps


def anomaly_maps(patch_features: Sequence[torch.Tensor], epoch_text_feature: torch.Tensor, det_feature: torch.Tensor,
                 img_size: int, domain: str = "Industrial"):
    """One fused call for what test_last.get_predictions does per batch (test_last.py:90-100,149): the level-summed
    test-mode map [B,S,S] and the image score ((det . text[:,1]) + 1) / 2 [B]."""
    ks, sg = DOMAIN_BLUR["Industrial" if domain == "Industrial" else "Medical"]
    return _engine.score_maps(list(patch_features), epoch_text_feature, img_size, test=True, ksize=ks, sigma=sg,
                              det=det_feature)


def _reference_tokenizer() -> Callable:
    try:
        from model.tokenizer import tokenize   # the reference's BPE tokenizer, if its tree is importable
        return tokenize
    except Exception as e:   # pragma: no cover
        raise RuntimeError("no tokenizer: pass tokenize=<callable(list[str]) -> int tensor [n,77]> (the reference's "
                           "model.tokenizer.tokenize); BPE tokenisation is outside this build") from e


def get_adapted_single_class_text_embedding(model, dataset_name: str, class_name: str, device,
                                            tokenize: Optional[Callable] = None,
                                            real_name: Optional[str] = None) -> torch.Tensor:
    """forward_utils.py:138-162 -> [768,2] unit-norm columns (normal, abnormal). `real_name` overrides the lookup in
    the reference's dataset.constants.REAL_NAMES (used when that module is not importable)."""
    if real_name is None:
        if class_name == "object":
            real_name = class_name
        else:
            from dataset.constants import CLASS_NAMES, REAL_NAMES
            assert class_name in CLASS_NAMES[dataset_name], (
                f"class_name {class_name} not found; available class_names: {CLASS_NAMES[dataset_name]}")
            real_name = REAL_NAMES[dataset_name][class_name]
    tokenize = tokenize or _reference_tokenizer()
    sentences, sizes = [], []
    for states in (PROMPTS["prompt_normal"], PROMPTS["prompt_abnormal"]):
        sent = [t.format(s.format(real_name)) for s in states for t in PROMPTS["prompt_templates"]]
        sentences += sent
        sizes.append(len(sent))
    emb = model.encode_text(tokenize(sentences).to(device))       # both states in one engine call
    return _engine.text_anchors(emb, sizes)[0]


def get_adapted_text_embedding(model, dataset_name: str, device, tokenize: Optional[Callable] = None,
                               real_names: Optional[Dict[str, str]] = None) -> Dict[str, torch.Tensor]:
    """forward_utils.py:185-193 -> {class_name: [768,2]}. All classes are encoded in ONE batched engine call
    (the reference loops class by class, two encoder calls each)."""
    if real_names is None:
        from dataset.constants import CLASS_NAMES, REAL_NAMES
        real_names = {c: REAL_NAMES[dataset_name][c] for c in CLASS_NAMES[dataset_name]}
    tokenize = tokenize or _reference_tokenizer()
    sentences, sizes = [], []
    for real in real_names.values():
        for states in (PROMPTS["prompt_normal"], PROMPTS["prompt_abnormal"]):
            sent = [t.format(s.format(real)) for s in states for t in PROMPTS["prompt_templates"]]
            sentences += sent
            sizes.append(len(sent))
    emb = model.encode_text(tokenize(sentences).to(device))
    anchors = _engine.text_anchors(emb, sizes)
    return {c: anchors[i] for i, c in enumerate(real_names)}


def metrics_eval(pixel_label: np.ndarray, image_label: np.ndarray, pixel_preds: np.ndarray, image_preds: np.ndarray,
                 class_names: str, domain: str) -> dict:
    """forward_utils.py:233-308, same semantics: global min-max of maps and scores over the class, image score =
    0.5*max-pixel + 0.5*det score (Industrial) or max-pixel (Medical), pixel / image AUROC and AP (scikit-learn),
    rounded to 4 decimals and scaled by 100."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    pixel_preds = np.asarray(pixel_preds, dtype=np.float32)
    image_preds = np.asarray(image_preds, dtype=np.float32)
    if pixel_preds.max() != 1:
        pixel_preds = (pixel_preds - pixel_preds.min()) / (pixel_preds.max() - pixel_preds.min())
    if image_preds.max() != 1:
        image_preds = (image_preds - image_preds.min()) / (image_preds.max() - image_preds.min())
    if pixel_preds.ndim == 4 and pixel_preds.shape[1] == 1:
        pixel_preds = pixel_preds.squeeze(1)
    elif pixel_preds.ndim == 2:
        side = int(pixel_preds.shape[1] ** 0.5)
        if side * side == pixel_preds.shape[1]:
            pixel_preds = pixel_preds.reshape(pixel_preds.shape[0], side, side)
    if image_preds.ndim == 2 and image_preds.shape[1] == 2:
        image_preds = image_preds[:, 0]            # the fork's [B,2] score quirk, kept as is
    elif image_preds.ndim > 1:
        image_preds = image_preds.flatten()
    pmax = pixel_preds.max(axis=(1, 2))
    image_preds = pmax * 0.5 + image_preds * 0.5 if domain != "Medical" else pmax
    pl, pp = np.asarray(pixel_label).flatten(), pixel_preds.flatten()
    pixel_auc, pixel_ap = roc_auc_score(pl, pp), average_precision_score(pl, pp)
    il = np.asarray(image_label)
    if il.max() != il.min():
        image_auc = roc_auc_score(il.flatten(), image_preds.flatten())
        image_ap = average_precision_score(il.flatten(), image_preds.flatten())
    else:
        image_auc = image_ap = 0
    return {"class name": class_names, "pixel AUC": round(pixel_auc, 4) * 100, "pixel AP": round(pixel_ap, 4) * 100,
            "image AUC": round(image_auc, 4) * 100, "image AP": round(image_ap, 4) * 100}
