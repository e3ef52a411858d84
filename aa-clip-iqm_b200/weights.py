"""Deterministic synthetic weights for the adapted ViT-L-14-336 visual tower (and text tower).

There are no pretrained checkpoints or network access in this environment, so the bench, the smoke
test and the parity tests all use random weights of the reference architecture, generated from a seed with a CPU
torch.Generator so that the build container (which runs the reference to make golden vectors) and the GPU box see
bit-identical tensors. Key names are the reference's ``AdaptedCLIP.state_dict()`` keys ([probed]):
``image_encoder.*`` for the CLIP visual tower, ``image_adapter.*`` for adapters and projections.

Scales follow the reference's effective initialisation (PyTorch defaults for Linear/MultiheadAttention,
xavier_uniform for adapters -- model/adapter.py:107-123) but, unlike it, biases and LayerNorm affine parameters are
non-trivial so that every term of the path is exercised by the parity tests.
"""
from __future__ import annotations

from typing import Dict

import torch

VIT_L_14 = dict(patch=14, width=1024, heads=16, layers=24, mlp_dim=4096, embed_dim=768, adapt_until=6,
                levels=(6, 12, 18, 24))
TEXT_B = dict(context=77, vocab=49408, width=768, heads=12, layers=12, mlp_dim=3072, adapt_until=3)


def make_visual_state_dict(seed: int = 0, img_size: int = 518, relu: bool = False, cfg: dict = VIT_L_14
                           ) -> Dict[str, torch.Tensor]:
    """fp32 CPU tensors, keys relative to AdaptedCLIP (image_encoder.* / image_adapter.*)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    W, F, E, ps = cfg["width"], cfg["mlp_dim"], cfg["embed_dim"], cfg["patch"]
    n_tok = (img_size // ps) ** 2 + 1

    def rn(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    p = "image_encoder."
    sd[p + "conv1.weight"] = rn(W, 3, ps, ps, std=(3 * ps * ps) ** -0.5)
    sd[p + "class_embedding"] = rn(W, std=W ** -0.5)
    sd[p + "positional_embedding"] = rn(n_tok, W, std=W ** -0.5)
    for ln in ("ln_pre", "ln_post"):
        sd[p + ln + ".weight"] = rn(W, std=0.1, mean=1.0)
        sd[p + ln + ".bias"] = rn(W, std=0.05)
    for i in range(cfg["layers"]):
        b = f"{p}transformer.resblocks.{i}."
        sd[b + "ln_1.weight"] = rn(W, std=0.1, mean=1.0)
        sd[b + "ln_1.bias"] = rn(W, std=0.05)
        sd[b + "attn.in_proj_weight"] = rn(3 * W, W, std=0.022)
        sd[b + "attn.in_proj_bias"] = rn(3 * W, std=0.02)
        sd[b + "attn.out_proj.weight"] = rn(W, W, std=0.018)
        sd[b + "attn.out_proj.bias"] = rn(W, std=0.02)
        sd[b + "ln_2.weight"] = rn(W, std=0.1, mean=1.0)
        sd[b + "ln_2.bias"] = rn(W, std=0.05)
        sd[b + "mlp.c_fc.weight"] = rn(F, W, std=0.018)
        sd[b + "mlp.c_fc.bias"] = rn(F, std=0.02)
        sd[b + "mlp.c_proj.weight"] = rn(W, F, std=0.009)
        sd[b + "mlp.c_proj.bias"] = rn(W, std=0.02)
    a = "image_adapter."
    for i in range(cfg["adapt_until"]):
        sd[f"{a}layer_adapters.{i}.fc.0.weight"] = rn(W, W, std=0.031)
    fc = "fc.0.weight" if relu else "fc.weight"
    for l in range(len(cfg["levels"])):
        sd[f"{a}seg_proj.{l}.{fc}"] = rn(E, W, std=0.033)
    sd[f"{a}det_proj.{fc}"] = rn(E, W, std=0.033)
    return sd


def make_text_state_dict(seed: int = 0, cfg: dict = TEXT_B) -> Dict[str, torch.Tensor]:
    """fp32 CPU tensors of the adapted text tower, keys relative to AdaptedCLIP (clipmodel.* / text_adapter.*).
    Scales follow the reference's init (model/transformer.py TextTransformer.init_parameters: token embedding
    std 0.02, positional 0.01, attn std W^-0.5, proj std (2W*layers)^-0.5... approximated) with non-trivial
    biases / LayerNorm affines so every term is exercised."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 29)
    W, F, L = cfg["width"], cfg["mlp_dim"], cfg["layers"]

    def rn(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    p = "clipmodel."
    sd[p + "token_embedding.weight"] = rn(cfg["vocab"], W, std=0.02)
    sd[p + "positional_embedding"] = rn(cfg["context"], W, std=0.01)
    for i in range(L):
        b = f"{p}transformer.resblocks.{i}."
        sd[b + "ln_1.weight"] = rn(W, std=0.1, mean=1.0)
        sd[b + "ln_1.bias"] = rn(W, std=0.05)
        sd[b + "attn.in_proj_weight"] = rn(3 * W, W, std=W ** -0.5)
        sd[b + "attn.in_proj_bias"] = rn(3 * W, std=0.02)
        sd[b + "attn.out_proj.weight"] = rn(W, W, std=(2 * W * L) ** -0.5 * 4)
        sd[b + "attn.out_proj.bias"] = rn(W, std=0.02)
        sd[b + "ln_2.weight"] = rn(W, std=0.1, mean=1.0)
        sd[b + "ln_2.bias"] = rn(W, std=0.05)
        sd[b + "mlp.c_fc.weight"] = rn(F, W, std=(2 * W) ** -0.5)
        sd[b + "mlp.c_fc.bias"] = rn(F, std=0.02)
        sd[b + "mlp.c_proj.weight"] = rn(W, F, std=(2 * W * L) ** -0.5 * 2)
        sd[b + "mlp.c_proj.bias"] = rn(W, std=0.02)
    sd[p + "ln_final.weight"] = rn(W, std=0.1, mean=1.0)
    sd[p + "ln_final.bias"] = rn(W, std=0.05)
    for i in range(cfg["adapt_until"] + 1):
        sd[f"text_adapter.{i}.fc.0.weight"] = rn(W, W, std=0.036)
    return sd


def make_text_anchor(seed: int = 0, embed_dim: int = 768, batch: int | None = None) -> torch.Tensor:
    """Unit-norm synthetic text anchors [E,2] (column 0 normal, 1 abnormal) or [B,E,2]."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 3)
    shape = (embed_dim, 2) if batch is None else (batch, embed_dim, 2)
    t = torch.randn(*shape, generator=g, dtype=torch.float32)
    return t / t.norm(dim=-2, keepdim=True)


def make_images(seed: int, batch: int, img_size: int, first_index: int = 0) -> torch.Tensor:
    """Synthetic CLIP-normalised-looking images; image i depends only on (seed, first_index + i), so any
    data-parallel sharding reproduces the same per-image inputs."""
    out = torch.empty(batch, 3, img_size, img_size, dtype=torch.float32)
    g = torch.Generator(device="cpu")
    for i in range(batch):
        g.manual_seed(1_000_000 * seed + 1000 + first_index + i)
        out[i] = torch.randn(3, img_size, img_size, generator=g, dtype=torch.float32)
    return out


IQM_BRANCH_MODULES = ("iqm", "class_query_mlp", "query_adapters", "iqm_layer_norm", "visual_feature_proj",
                      "text_feature_proj")


def fill_iqm_branch_(model, seed: int = 0, n_levels: int = 4) -> int:
    """Deterministic weights for the IQM branch of an ``AdaptedCLIP`` -- the reference's class or this repository's
    mirror: both expose the same sub-modules under the same parameter names (model/adapter.py:56-96). The two layers
    the reference creates lazily on its first forward (``visual_feature_proj`` Linear(hidden, hidden),
    ``text_feature_proj`` Linear(2, 768), model/adapter.py:215-218,242-243) are created here, so that both sides run
    the same tensors. Parameters are filled in sorted-name order from one CPU generator: the same seed gives the same
    tensors in the build container (golden generation with the reference) and on the GPU box. Returns the count."""
    import torch.nn as nn
    hidden = model.iqm_hidden_size
    dev = next(model.iqm.parameters()).device
    if model.visual_feature_proj is None:
        model.visual_feature_proj = nn.Linear(hidden, hidden).to(dev)
    if model.text_feature_proj is None:
        model.text_feature_proj = nn.Linear(2, 768).to(dev)
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 41)
    n = 0
    with torch.no_grad():
        for mod_name in IQM_BRANCH_MODULES:
            mod = getattr(model, mod_name)
            for name, p in sorted(mod.named_parameters(), key=lambda kv: kv[0]):
                if p.dim() >= 2:
                    t = torch.randn(p.shape, generator=g) * (p.shape[-1] ** -0.5)
                elif name.endswith("LayerNorm.weight") or name in ("weight",) or name.endswith("layernorm.weight"):
                    t = torch.randn(p.shape, generator=g) * 0.1 + 1.0
                else:
                    t = torch.randn(p.shape, generator=g) * 0.05
                p.copy_(t.to(p.device))
                n += 1
    return n
