"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU fp32 restatement (plain torch tensor ops, no nn.Module from the reference) of the reference algorithm on the
AA-CLIP inference hot path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this; the product path (aa-clip-iqm_b200/) never does.

Pinning: oracle/gen_golden.py runs the UNMODIFIED reference (/root/reference, imported through oracle/ref_import.py)
and this restatement on the same weights and inputs in the build container and commits the reference's outputs under
tests/golden/; tests/test_oracle_golden.py checks this file against those vectors (fp32, <= 2e-5 abs on maps).
The reference ships no tests or golden vectors of its own. kornia's gaussian_blur2d is a
third-party dependency that is not installable here: its published algorithm is restated in `gaussian_blur2d` below
and that one function is "parity unpinned" against real kornia 0.6.9 (cross-checked against torchvision instead).

All file:line citations are relative to the reference root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

LEVELS = (6, 12, 18, 24)


# ------------------------------------------------------------------------------------------------ blocks
def layer_norm(x, w, b):
    """model/transformer.py LayerNorm (eps 1e-5; model/model.py:96 selects the plain, non-fp32-cast variant)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def multi_head_attention(x, in_w, in_b, out_w, out_b, heads: int, attn_mask: Optional[torch.Tensor] = None):
    """nn.MultiheadAttention forward as called at model/transformer.py:200,226-237 (batch_first=False, slow path):
    packed in-projection, q scaled by head_dim**-0.5 BEFORE q@k^T, additive float mask, softmax, out-projection.
    x: [B, N, D] here (the reference holds it as [N, B, D]; per-image math is identical)."""
    B, N, D = x.shape
    hd = D // heads
    qkv = x @ in_w.t() + in_b
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, N, heads, hd).transpose(1, 2) * (hd ** -0.5)
    k = k.view(B, N, heads, hd).transpose(1, 2)
    v = v.view(B, N, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if attn_mask is not None:
        s = s + attn_mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, D)
    return o @ out_w.t() + out_b


def residual_attention_block(x, sd: Dict[str, torch.Tensor], prefix: str, heads: int, quick_gelu: bool = False,
                             attn_mask: Optional[torch.Tensor] = None):
    """ResidualAttentionBlock.forward, model/transformer.py:239-258 (ls_1/ls_2 are Identity: ls_init_value=None)."""
    h = layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"])
    x = x + multi_head_attention(h, sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"],
                                 sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"], heads,
                                 attn_mask)
    h = layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    h = h @ sd[prefix + "mlp.c_fc.weight"].t() + sd[prefix + "mlp.c_fc.bias"]
    # nn.GELU (erf) is what create_model builds (model/model.py:84, config has no quick_gelu key);
    # QuickGELU = x*sigmoid(1.702x) only with force_quick_gelu (model/transformer.py:46-49, model/clip.py:151-153)
    h = h * torch.sigmoid(1.702 * h) if quick_gelu else F.gelu(h)
    h = h @ sd[prefix + "mlp.c_proj.weight"].t() + sd[prefix + "mlp.c_proj.bias"]
    return x + h


def adapter_blend(x, w_adapter, weight: float):
    """model/adapter.py:163-170: SimpleAdapter = Linear(no bias)+LeakyReLU(0.01) (model/adapter_modules.py:6-13);
    rescale to the token norm (no eps), then i_w*adapt + (1-i_w)*x."""
    a = F.leaky_relu(x @ w_adapter.t(), 0.01)
    a = a * x.norm(dim=-1, keepdim=True) / a.norm(dim=-1, keepdim=True)
    return weight * a + (1 - weight) * x


def simple_proj(t, sd, prefix):
    """SimpleProj, model/adapter_modules.py:16-26: Linear(no bias) [+ LeakyReLU when built with relu=True]."""
    if prefix + "fc.0.weight" in sd:
        return F.leaky_relu(t @ sd[prefix + "fc.0.weight"].t(), 0.01)
    return t @ sd[prefix + "fc.weight"].t()


# ------------------------------------------------------------------------------------------------ visual tower
def visual_forward(sd: Dict[str, torch.Tensor], images: torch.Tensor, patch: int = 14, heads: int = 16,
                   layers: int = 24, adapt_until: int = 6, levels: Sequence[int] = LEVELS,
                   adapt_weight: float = 0.1, quick_gelu: bool = False, return_stages: bool = False):
    """AdaptedCLIP.forward without the IQM branch, model/adapter.py:137-184.

    Returns dict(seg=[4 x [B,P,E]], det=[B,E], tokens=[4 x [B,P,W]] (post ln_post), cls=[B,W] (final x[:,0])).
    """
    p = "image_encoder."
    images = images.float()
    # patch embedding conv, model/adapter.py:139-141 (VisionTransformer.conv1, model/transformer.py:359-365)
    x = F.conv2d(images, sd[p + "conv1.weight"], stride=patch)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    # CLS + positional embedding, model/adapter.py:143-153 ; patch_dropout is identity in eval (:155)
    cls = sd[p + "class_embedding"] + torch.zeros(x.shape[0], 1, x.shape[-1])
    x = torch.cat([cls, x], dim=1) + sd[p + "positional_embedding"]
    x = layer_norm(x, sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"])  # :156
    stages = {"ln_pre": x} if return_stages else None
    taps: List[torch.Tensor] = []
    for i in range(layers):  # :161 (hard-coded 24 in the reference)
        x = residual_attention_block(x, sd, f"{p}transformer.resblocks.{i}.", heads, quick_gelu)
        if i < adapt_until:  # :163
            x = adapter_blend(x, sd[f"image_adapter.layer_adapters.{i}.fc.0.weight"], adapt_weight)
        if i + 1 in levels:  # :171-172, post-blend, CLS dropped
            taps.append(x[:, 1:, :])
        if return_stages:
            stages[f"block{i}"] = x
    tokens = [layer_norm(t, sd[p + "ln_post.weight"], sd[p + "ln_post.bias"]) for t in taps]  # :176
    seg = [F.normalize(simple_proj(t, sd, f"image_adapter.seg_proj.{l}."), dim=-1) for l, t in enumerate(tokens)]  # :179-182
    det = F.normalize(simple_proj(tokens[-1], sd, "image_adapter.det_proj."), dim=-1).mean(1)  # :183-184
    out = dict(seg=seg, det=det, tokens=tokens, cls=x[:, 0, :])
    if return_stages:
        out["stages"] = stages
    return out


# ------------------------------------------------------------------------------------------------ scoring
def gaussian_taps(ksize: int, sigma: float) -> torch.Tensor:
    x = torch.arange(ksize, dtype=torch.float32) - ksize // 2
    if ksize % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * float(sigma) ** 2))
    return g / g.sum()


def gaussian_blur2d(x: torch.Tensor, ksize: int, sigma: float) -> torch.Tensor:
    """kornia==0.6.9 filters.gaussian_blur2d as called at forward_utils.py:208-210 (border_type='reflect'):
    2-D kernel = outer product of normalised 1-D Gaussian taps, reflect padding ksize//2, depthwise correlation.
    PARITY UNPINNED against kornia itself (not installed / not vendored)."""
    w = gaussian_taps(ksize, sigma)
    k2 = (w[:, None] @ w[None, :])[None, None]
    c = x.shape[1]
    xp = F.pad(x, (ksize // 2,) * 4, mode="reflect")
    return F.conv2d(xp, k2.expand(c, 1, ksize, ksize), groups=c)


def calculate_similarity_map(patch_features, text, img_size: int, test: bool = False, domain: str = "Medical"):
    """forward_utils.calculate_similarity_map, forward_utils.py:196-216. text: [E,2] or [B,E,2]."""
    s = 100.0 * torch.matmul(patch_features, text)
    B, L, C = s.shape
    H = int(math.isqrt(L))
    pred = s.permute(0, 2, 1).reshape(B, C, H, H)
    if test:
        assert C == 2
        sigma = 1 if domain == "Industrial" else 1.5
        ksize = 7 if domain == "Industrial" else 9
        pred = (pred[:, 1] + 1 - pred[:, 0]) / 2
        pred = gaussian_blur2d(pred.unsqueeze(1), ksize, sigma)
    pred = F.interpolate(pred, size=img_size, mode="bilinear", align_corners=True)
    if not test and C > 1:
        pred = torch.softmax(pred, dim=1)
    return pred


def anomaly_maps(seg: Sequence[torch.Tensor], text, img_size: int, domain: str = "Industrial") -> torch.Tensor:
    """Sum of the per-level test-mode maps, test_last.py:96-100,149 (the no-IQM branch) -> [B,S,S]."""
    maps = [calculate_similarity_map(f, text, img_size, test=True, domain=domain) for f in seg]
    return torch.cat(maps, dim=1).sum(1)


def image_score(det: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
    """Image-level score at the well-posed level: upstream AA-CLIP semantics of
    test_last.py:90-92 with a 2-D anchor: ((det @ text)[:,1] + 1) / 2. With per-image anchors [B,E,2] the same
    formula is applied image by image (the fork's [B,B,2] broadcast is a host-side bug, not reproduced here)."""
    if text.dim() == 2:
        return ((det @ text)[:, 1] + 1) / 2
    return (torch.einsum("be,be->b", det, text[:, :, 1]) + 1) / 2


def predict(sd, images, text, img_size: int, domain: str = "Industrial", **kw):
    """Whole hot path for a batch: tower -> summed anomaly maps [B,S,S] + image scores [B]."""
    out = visual_forward(sd, images, **kw)
    out["maps"] = anomaly_maps(out["seg"], text, img_size, domain)
    out["score"] = image_score(out["det"], text)
    return out


# ------------------------------------------------------------------------------------------------ text anchors
def encode_text(sd: Dict[str, torch.Tensor], tokens: torch.Tensor, heads: int = 12, layers: int = 12,
                adapt_until: int = 3, adapt_weight: float = 0.1, quick_gelu: bool = False) -> torch.Tensor:
    """AdaptedCLIP.encode_text(text, adapt_text=True), model/adapter.py:273-304. tokens: int [n, 77] -> [n, 768].
    Causal additive mask as built by TextTransformer.build_attention_mask (model/transformer.py:629-635)."""
    p = "clipmodel."
    x = sd[p + "token_embedding.weight"][tokens.long()] + sd[p + "positional_embedding"]      # :277-281
    T = x.shape[1]
    mask = torch.full((T, T), float("-inf")).triu_(1)
    for i in range(layers):                                                                      # :284 (hard-coded 12)
        x = residual_attention_block(x, sd, f"{p}transformer.resblocks.{i}.", heads, quick_gelu, attn_mask=mask)
        if i < adapt_until:                                                                      # :288-295
            x = adapter_blend(x, sd[f"text_adapter.{i}.fc.0.weight"], adapt_weight)
    x = layer_norm(x, sd[p + "ln_final.weight"], sd[p + "ln_final.bias"])                        # :297
    eot = x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)]                                     # :298 (EOT = highest id)
    return F.leaky_relu(eot @ sd[f"text_adapter.{adapt_until}.fc.0.weight"].t(), 0.01)           # SimpleProj(relu=True)


def text_anchor(emb_normal: torch.Tensor, emb_abnormal: torch.Tensor) -> torch.Tensor:
    """forward_utils.get_adapted_single_class_text_embedding, forward_utils.py:152-161: per state, L2-normalise each
    sentence embedding, average, re-normalise; stack the two states -> [E, 2]."""
    feats = []
    for e in (emb_normal, emb_abnormal):
        e = e / e.norm(dim=-1, keepdim=True)
        m = e.mean(dim=0)
        feats.append(m / m.norm())
    return torch.stack(feats, dim=1)
