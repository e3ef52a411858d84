"""Mirror of the reference's scoring / evaluation glue (forward_utils.py) for the hot path: same function names,
arguments and return shapes; the device work goes through the C ABI (aaclip_score_maps, aaclip_encode_text,
aaclip_text_anchors). ``metrics_eval`` keeps the reference's host semantics (rows a11: numpy + scikit-learn).

Two routes, as in model/adapter.py: under ``torch.no_grad()`` (evaluation) everything below runs in the sm_100a engine
and raises without it; when autograd needs a gradient through a call (train.py:76-93,150-166) the same arithmetic is
expressed in differentiable PyTorch ops. ``calculate_seg_loss`` is training-only host code."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import engine as _engine
from .constants import CLASS_NAMES, DATA_PATH, DOMAINS, PROMPTS, REAL_NAMES, class_prompts  # noqa: F401
from .tokenizer import tokenize as _tokenize

DOMAIN_BLUR = {"Industrial": (7, 1.0), "Medical": (9, 1.5)}       # forward_utils.py:205-206


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _gaussian_blur(x: torch.Tensor, ksize: int, sigma: float) -> torch.Tensor:
    """kornia.filters.gaussian_blur2d(x, (k,k), (s,s)) as published: normalised exp(-t^2 / 2 s^2) taps on
    t = -(k-1)/2 .. (k-1)/2, reflect border, separable depthwise convolution."""
    t = torch.arange(ksize, dtype=x.dtype, device=x.device) - (ksize - 1) / 2
    w = torch.exp(-0.5 * (t / sigma) ** 2)
    w = w / w.sum()
    C = x.shape[1]
    x = F.pad(x, (ksize // 2,) * 4, mode="reflect")
    x = F.conv2d(x, w.view(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)
    return F.conv2d(x, w.view(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)


def _similarity_map_autograd(patch_features, text, img_size, test, ks, sg):
    s = 100.0 * torch.matmul(patch_features, text)                      # [B,P,2]
    B, P, C = s.shape
    H = int(np.sqrt(P))
    s = s.permute(0, 2, 1).reshape(B, C, H, H)
    if test:
        m = ((s[:, 1] + 1 - s[:, 0]) / 2).unsqueeze(1)
        m = _gaussian_blur(m, ks, sg)
        return F.interpolate(m, size=img_size, mode="bilinear", align_corners=True)
    return torch.softmax(F.interpolate(s, size=img_size, mode="bilinear", align_corners=True), dim=1)


def calculate_similarity_map(patch_features: torch.Tensor, epoch_text_feature: torch.Tensor, img_size: int,
                             test: bool = False, domain: str = "Medical") -> torch.Tensor:
    """forward_utils.py:196-216. patch_features [B,P,768] unit-norm; epoch_text_feature [768,2] or [B,768,2].
    test=True -> [B,1,S,S] = upsample(blur((s1+1-s0)/2)); test=False -> [B,2,S,S] = softmax(upsample(s))."""
    ks, sg = DOMAIN_BLUR["Industrial" if domain == "Industrial" else "Medical"]
    if epoch_text_feature.shape[-1] != 2:
        raise AssertionError("calculate_similarity_map expects two anchors (normal, abnormal)")   # :204 `assert C == 2`
    if _wants_grad(patch_features, epoch_text_feature):
        return _similarity_map_autograd(patch_features, epoch_text_feature, img_size, test, ks, sg)
    maps, _ = _engine.score_maps([patch_features], epoch_text_feature, img_size, test=test, ksize=ks, sigma=sg)
    return maps.unsqueeze(1) if test else maps


def anomaly_maps(patch_features: Sequence[torch.Tensor], epoch_text_feature: torch.Tensor, det_feature: torch.Tensor,
                 img_size: int, domain: str = "Industrial"):
    """One fused call for what test_last.get_predictions does per batch (test_last.py:90-100,149): the level-summed
    test-mode map [B,S,S] and the image score ((det . text[:,1]) + 1) / 2 [B]."""
    ks, sg = DOMAIN_BLUR["Industrial" if domain == "Industrial" else "Medical"]
    return _engine.score_maps(list(patch_features), epoch_text_feature, img_size, test=True, ksize=ks, sigma=sg,
                              det=det_feature)


# ---------------------------------------------------------------------------------------------- training losses
def _focal_loss(prob: torch.Tensor, target: torch.Tensor, gamma: float = 2.0, smooth: float = 1e-5) -> torch.Tensor:
    """Focal loss on class probabilities [N,C,H,W] against integer masks [N,1,H,W] with label smoothing `smooth`,
    alpha = 1, mean over pixels (forward_utils.py:21-107 at its defaults)."""
    C = prob.shape[1]
    p = prob.permute(0, 2, 3, 1).reshape(-1, C)
    idx = target.reshape(-1, 1).long()
    onehot = torch.zeros_like(p).scatter_(1, idx, 1.0).clamp(smooth / (C - 1), 1.0 - smooth)
    pt = (onehot * p).sum(1) + smooth
    return (-((1 - pt) ** gamma) * pt.log()).mean()


def _dice_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """1 - mean_n (2 <p,t> + 1) / (sum p + sum t + 1)   (forward_utils.py:110-125)."""
    N = target.shape[0]
    p, t = pred.reshape(N, -1), target.reshape(N, -1)
    return 1 - ((2 * (p * t).sum(1) + 1) / (p.sum(1) + t.sum(1) + 1)).sum() / N


def calculate_seg_loss(patch_preds: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """focal + dice(normal channel, 1 - mask) + dice(abnormal channel, mask)   (forward_utils.py:222-226)."""
    return (_focal_loss(patch_preds, mask) + _dice_loss(patch_preds[:, 0], 1 - mask)
            + _dice_loss(patch_preds[:, 1], mask))


# ---------------------------------------------------------------------------------------------- text anchors
def _real_name(dataset_name: str, class_name: str) -> str:
    if class_name == "object":
        return class_name
    assert class_name in CLASS_NAMES[dataset_name], (
        f"class_name {class_name} not found; available class_names: {CLASS_NAMES[dataset_name]}")
    return REAL_NAMES[dataset_name][class_name]


def _encode_groups(model, groups, device, tokenize):
    """groups: list of sentence lists (normal, abnormal, normal, abnormal, ...) -> anchors [n_classes, E, 2]."""
    sentences = [s for g in groups for s in g]
    sizes = [len(g) for g in groups]
    emb = model.encode_text(tokenize(sentences).to(device))          # every sentence in one encoder call
    if _wants_grad(emb):
        # training route (train.py:66-70 differentiates through the anchors): forward_utils.py:155-161 in torch ops
        cols, o = [], 0
        for n in sizes:
            e = emb[o:o + n]
            e = e / e.norm(dim=-1, keepdim=True)
            m = e.mean(dim=0)
            cols.append(m / m.norm())
            o += n
        return torch.stack([torch.stack(cols[i:i + 2], dim=1) for i in range(0, len(cols), 2)])
    return _engine.text_anchors(emb, sizes)


def get_adapted_single_class_text_embedding(model, dataset_name: str, class_name: str, device,
                                            tokenize: Optional[Callable] = None,
                                            real_name: Optional[str] = None) -> torch.Tensor:
    """forward_utils.py:138-162 -> [768,2] unit-norm columns (normal, abnormal). `tokenize` / `real_name` override the
    built-in BPE tokenizer and the REAL_NAMES lookup."""
    real = real_name if real_name is not None else _real_name(dataset_name, class_name)
    return _encode_groups(model, class_prompts(real), device, tokenize or _tokenize)[0]


def get_adapted_single_sentence_text_embedding(model, dataset_name: str, class_name: str, device) -> torch.Tensor:
    """forward_utils.py:165-182 -> the 16 unit-norm sentence embeddings of a class [16,768] (no ensembling)."""
    sentences = [s for g in class_prompts(_real_name(dataset_name, class_name)) for s in g]
    return F.normalize(model.encode_text(_tokenize(sentences).to(device)), dim=-1)


def get_adapted_text_embedding(model, dataset_name: str, device, tokenize: Optional[Callable] = None,
                               real_names: Optional[Dict[str, str]] = None) -> Dict[str, torch.Tensor]:
    """forward_utils.py:185-193 -> {class_name: [768,2]}. All classes are encoded in ONE batched engine call
    (the reference loops class by class, two encoder calls each)."""
    if real_names is None:
        real_names = {c: REAL_NAMES[dataset_name][c] for c in CLASS_NAMES[dataset_name]}
    groups = [g for real in real_names.values() for g in class_prompts(real)]
    anchors = _encode_groups(model, groups, device, tokenize or _tokenize)
    return {c: anchors[i] for i, c in enumerate(real_names)}


def visualize(*args, **kwargs):
    raise NotImplementedError("visualize (forward_utils.py:318-354: OpenCV heat-map overlays written to disk) is "
                              "outside the hot path and is not provided")


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
