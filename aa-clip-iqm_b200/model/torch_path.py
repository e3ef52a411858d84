"""Autograd route of the mirror: the towers as plain PyTorch ops on the SAME ``nn.Parameter``s the engine packs.

This is the TRAINING path, not a fallback for inference:
``AdaptedCLIP.forward`` / ``encode_text`` come here only when autograd is enabled and a parameter requires grad
(train.py:150,219 run the model that way to train the adapters); under ``torch.no_grad()`` the sm_100a engine runs and
raises when it cannot. It also backs the two un-adapted CLIP helpers of the reference surface that are not on the hot
path -- ``ClipWeights.encode_image`` (forward_original, train.py stage 1) and ``ClipWeights.encode_text``
(test_last.py:264-271 when no text-adapter checkpoint exists).

Parameters are addressed by their state_dict names (``transformer.resblocks.3.attn.in_proj_weight`` ...), so the same
code serves ``ClipWeights`` and a reference ``CLIP`` instance. Behaviour follows model/adapter.py:137-184,273-304 and
model/transformer.py:239-258,488-537; the reference's quirks on this path are kept (adapter norm-rescale without an
epsilon; erf-GELU unless the model was built with quick_gelu).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


def named(module: torch.nn.Module) -> Dict[str, torch.Tensor]:
    out = dict(module.named_parameters())
    out.update(dict(module.named_buffers()))
    return out


def _act(x: torch.Tensor, quick: bool) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x) if quick else F.gelu(x)


def residual_block(x: torch.Tensor, P: Dict[str, torch.Tensor], pre: str, causal: bool, quick_gelu: bool) -> torch.Tensor:
    """One pre-LN residual attention block on x [B, N, W] (model/transformer.py:239-258; head_dim 64)."""
    B, N, W = x.shape
    heads = W // 64
    h = F.layer_norm(x, (W,), P[pre + "ln_1.weight"], P[pre + "ln_1.bias"], 1e-5)
    qkv = F.linear(h, P[pre + "attn.in_proj_weight"], P[pre + "attn.in_proj_bias"])
    q, k, v = (t.view(B, N, heads, 64).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    a = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
    a = a.transpose(1, 2).reshape(B, N, W)
    x = x + F.linear(a, P[pre + "attn.out_proj.weight"], P[pre + "attn.out_proj.bias"])
    h = F.layer_norm(x, (W,), P[pre + "ln_2.weight"], P[pre + "ln_2.bias"], 1e-5)
    h = _act(F.linear(h, P[pre + "mlp.c_fc.weight"], P[pre + "mlp.c_fc.bias"]), quick_gelu)
    return x + F.linear(h, P[pre + "mlp.c_proj.weight"], P[pre + "mlp.c_proj.bias"])


def blend(x: torch.Tensor, adapter: torch.nn.Module, w: float) -> torch.Tensor:
    """x <- w * a * ||x|| / ||a|| + (1 - w) * x with a = adapter(x) (model/adapter.py:163-170)."""
    a = adapter(x)
    a = a * x.norm(dim=-1, keepdim=True) / a.norm(dim=-1, keepdim=True)
    return w * a + (1 - w) * x


def n_blocks(P: Dict[str, torch.Tensor], prefix: str = "transformer.resblocks.") -> int:
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in P if k.startswith(prefix))


def visual_embed(images: torch.Tensor, P: Dict[str, torch.Tensor]) -> torch.Tensor:
    """patch embedding + CLS + positional embedding + ln_pre -> [B, N, W] (model/adapter.py:139-156)."""
    w = P["conv1.weight"]
    x = F.conv2d(images, w, stride=w.shape[-1])
    x = x.flatten(2).transpose(1, 2)
    cls = P["class_embedding"].to(x.dtype).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + P["positional_embedding"].to(x.dtype)
    return F.layer_norm(x, (x.shape[-1],), P["ln_pre.weight"], P["ln_pre.bias"], 1e-5)


def visual_tower(images: torch.Tensor, P: Dict[str, torch.Tensor], layer_adapters: Sequence[torch.nn.Module],
                 adapt_weight: float, levels: Sequence[int], quick_gelu: bool
                 ) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """-> (ln_post'ed patch tokens [B,P,W] per level, final CLS row [B,W] WITHOUT ln_post) -- model/adapter.py:137-176
    and the input of the IQM branch (model/adapter.py:190)."""
    x = visual_embed(images, P)
    taps = []
    for i in range(n_blocks(P)):
        x = residual_block(x, P, f"transformer.resblocks.{i}.", False, quick_gelu)
        if i < len(layer_adapters):
            x = blend(x, layer_adapters[i], adapt_weight)
        if i + 1 in levels:
            taps.append(x[:, 1:, :])
    W = x.shape[-1]
    tokens = [F.layer_norm(t, (W,), P["ln_post.weight"], P["ln_post.bias"], 1e-5) for t in taps]
    return tokens, x[:, 0, :]


def clip_encode_image(images: torch.Tensor, P: Dict[str, torch.Tensor], out_layers: Sequence[int], quick_gelu: bool):
    """Un-adapted CLIP.encode_image(image, out_layers) (model/model.py:185-188, model/transformer.py:488-537):
    (ln_post(CLS) @ proj, [all tokens [B,N,W] after block l for l in out_layers])."""
    x = visual_embed(images, P)
    outs = []
    for i in range(n_blocks(P)):
        x = residual_block(x, P, f"transformer.resblocks.{i}.", False, quick_gelu)
        if i + 1 in out_layers:
            outs.append(x)
    pooled = F.layer_norm(x[:, 0], (x.shape[-1],), P["ln_post.weight"], P["ln_post.bias"], 1e-5)
    if "proj" in P:
        pooled = pooled @ P["proj"]
    return pooled, outs


def text_tower(tokens: torch.Tensor, P: Dict[str, torch.Tensor], adapters: Sequence[torch.nn.Module],
               adapt_weight: float, quick_gelu: bool) -> torch.Tensor:
    """token ids [n,77] -> ln_final'ed EOT rows [n,W] (model/adapter.py:273-298 up to the final projection; with no
    adapters this is CLIP.encode_text up to the text projection, model/model.py:190-201)."""
    ids = tokens.long()
    x = F.embedding(ids, P["token_embedding.weight"]) + P["positional_embedding"]
    for i in range(n_blocks(P)):
        x = residual_block(x, P, f"transformer.resblocks.{i}.", True, quick_gelu)
        if i < len(adapters):
            x = blend(x, adapters[i], adapt_weight)
    x = F.layer_norm(x, (x.shape[-1],), P["ln_final.weight"], P["ln_final.bias"], 1e-5)
    return x[torch.arange(x.shape[0], device=x.device), ids.argmax(dim=-1)]
