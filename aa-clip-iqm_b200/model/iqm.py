"""The fork's IQM query module as host-side PyTorch: a two-query Q-Former-style encoder that
reads the visual tower's multi-level tokens and the text anchors.

Restates the behaviour of the reference's ``model/iqm.py:23-673`` as ``AdaptedCLIP.forward`` actually drives it
(model/adapter.py:246-258): ``IQM(query_embeds, query_length, encoder_hidden_states, text_encoder_hidden_states)`` with
all masks at their all-ones defaults (they add exactly zero to every score), absolute position type, no cache, no head
pruning. Parameter names and shapes -- hence ``state_dict`` keys -- are the reference's:

    iqm.layernorm
    iqm.encoder.layer.<i>.attention|crossattention|text_crossattention . attention.{query,key,value} / output.{dense,LayerNorm}
    iqm.encoder.layer.<i>.intermediate.dense, output.{dense,LayerNorm}, intermediate_query.dense, output_query.{dense,LayerNorm}

so a checkpoint of the reference module loads with ``strict=True`` (pinned by tests/golden/ref_state_dict_keys.json),
and the outputs match the reference module on the same weights (tests/golden/iqm_*.npz, <= 1e-5).
This module has no dependency on ``transformers`` (the reference builds on HuggingFace 4.x classes that 5.x removed).

Per layer (model/iqm.py:233-330), with q = the two query rows:
    h  = LN(q + W_o SelfAttn(q))
    hc = LN(h + W_o' CrossAttn(h -> visual tokens))
    ht = LN(hc + W_o'' CrossAttn(hc -> text tokens))
    m  = 0.4 h + 0.3 hc + 0.3 ht
    q' = LN(m + W_2 gelu(W_1 m))            (the *_query feed-forward; `intermediate` / `output` serve rows beyond
                                             query_length, which AdaptedCLIP never passes -- kept for the key set)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn


@dataclass
class IQMConfig:
    hidden_size: int = 768
    num_hidden_layers: int = 2
    num_attention_heads: int = 8
    intermediate_size: int = 2048
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    layer_norm_eps: float = 1e-12
    cross_attention_frequency: int = 1
    encoder_hidden_size: int = 768
    text_encoder_hidden_size: int = 768


class _Heads(nn.Module):
    """query / key / value projections + scaled dot-product attention over `heads` heads (model/iqm.py:23-137)."""

    def __init__(self, cfg: IQMConfig, kv_width: int):
        super().__init__()
        if cfg.hidden_size % cfg.num_attention_heads:
            raise ValueError(f"The hidden size ({cfg.hidden_size}) is not a multiple of the number of attention heads "
                             f"({cfg.num_attention_heads})")
        self.heads = cfg.num_attention_heads
        self.head_dim = cfg.hidden_size // cfg.num_attention_heads
        self.query = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.key = nn.Linear(kv_width, cfg.hidden_size)
        self.value = nn.Linear(kv_width, cfg.hidden_size)
        self.dropout = nn.Dropout(cfg.attention_probs_dropout_prob)

    def _split(self, t: torch.Tensor) -> torch.Tensor:          # [B, n, H] -> [B, heads, n, head_dim]
        B, n, _ = t.shape
        return t.view(B, n, self.heads, self.head_dim).transpose(1, 2)

    def forward(self, x: torch.Tensor, memory=None) -> torch.Tensor:
        if isinstance(memory, FactoredMemory):
            return self.forward_factored(x, memory)
        src = x if memory is None else memory
        q, k, v = self._split(self.query(x)), self._split(self.key(src)), self._split(self.value(src))
        p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(self.head_dim), dim=-1)
        ctx = self.dropout(p) @ v
        B, _, n, _ = ctx.shape
        return ctx.transpose(1, 2).reshape(B, n, self.heads * self.head_dim)

    def forward_factored(self, x: torch.Tensor, mem: FactoredMemory) -> torch.Tensor:
        """Cross-attention of a FEW queries over a LONG memory u_i = Wp a_i + bp without projecting the memory.

        The reference computes k_i = Wk u_i + bk and v_i = Wv u_i + bv for all 5 476 memory rows of every image
        (3 x 5476 x H x H MACs per layer) to serve 2 queries. Per head h, with q the projected query:
            q . k_i   = (Wp^T Wk_h^T q) . a_i + (Wk_h^T q) . bp + q . bk_h        -> project the QUERY back instead
            sum_i p_i v_i = Wv_h (Wp (sum_i p_i a_i) + bp) + bv_h                  (sum_i p_i = 1)
        so the memory is touched by two skinny products per layer ([n_q * heads, Ha] x [Ha, P] and its transpose
        partner) and the concatenation over levels is a concatenation of the small score rows. Same result as
        `forward` up to fp32 re-association (tests: <= 1e-5 on the reference fixture)."""
        B, n, _ = x.shape
        h, d = self.heads, self.head_dim
        Wp, bp = mem.proj.weight, mem.proj.bias                     # [H, Ha], [H]
        q = self._split(self.query(x))                              # [B, h, n, d]
        Wk = self.key.weight.view(h, d, -1)                         # [h, d, H]
        g = torch.einsum("bhnd,hdk->bhnk", q, Wk)                   # Wk_h^T q          [B, h, n, H]
        const = g @ bp + torch.einsum("bhnd,hd->bhn", q, self.key.bias.view(h, d))
        gp = (g @ Wp).reshape(B, h * n, -1)                         # Wp^T Wk_h^T q     [B, h*n, Ha]
        scores = torch.cat([gp @ a.transpose(1, 2) for a in mem.levels], dim=-1)           # [B, h*n, sum P_l]
        scores = (scores + const.reshape(B, h * n, 1)) / math.sqrt(d)
        p = self.dropout(torch.softmax(scores, dim=-1))
        s, o = None, 0
        for a in mem.levels:                                        # sum_i p_i a_i     [B, h*n, Ha]
            part = p[:, :, o:o + a.shape[1]] @ a
            s = part if s is None else s + part
            o += a.shape[1]
        t = (s @ Wp.t() + bp).view(B, h, n, -1)                     # Wp s + bp         [B, h, n, H]
        ctx = torch.einsum("bhnk,hdk->bhnd", t, self.value.weight.view(h, d, -1)) + self.value.bias.view(1, h, 1, d)
        return ctx.transpose(1, 2).reshape(B, n, h * d)


class FactoredMemory:
    """The visual memory of the cross-attention in factored form: ``levels`` = the query-adapter outputs a_l
    [B, P_l, Ha] (one tensor per tapped level, NOT concatenated) and ``proj`` = the Linear(Ha, H) the reference applies
    to their concatenation (AdaptedCLIP.visual_feature_proj). The memory the reference materialises is
    u = proj(cat(levels, dim=1)) [B, sum P_l, H]; `_Heads.forward_factored` attends over it without building it."""

    def __init__(self, levels, proj: nn.Linear):
        self.levels, self.proj = list(levels), proj

    def materialise(self) -> torch.Tensor:
        return self.proj(torch.cat(self.levels, dim=1))


class _ResidualNorm(nn.Module):
    """dense -> dropout -> LayerNorm(out + residual)   (IQM_SelfOutput / IQM_Output, model/iqm.py:140-151,215-226)."""

    def __init__(self, cfg: IQMConfig, in_width: int):
        super().__init__()
        self.dense = nn.Linear(in_width, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)

    def forward(self, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        return self.LayerNorm(self.dropout(self.dense(x)) + residual)


class _AttentionBlock(nn.Module):
    def __init__(self, cfg: IQMConfig, kv_width: int):
        super().__init__()
        self.attention = _Heads(cfg, kv_width)
        self.output = _ResidualNorm(cfg, cfg.hidden_size)

    def forward(self, x: torch.Tensor, memory=None) -> torch.Tensor:
        return self.output(self.attention(x, memory), x)


class _Expand(nn.Module):
    def __init__(self, cfg: IQMConfig):
        super().__init__()
        self.dense = nn.Linear(cfg.hidden_size, cfg.intermediate_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.gelu(self.dense(x))               # hidden_act = "gelu" (exact erf form)


class IQMLayer(nn.Module):
    def __init__(self, cfg: IQMConfig, layer_idx: int):
        super().__init__()
        self.attention = _AttentionBlock(cfg, cfg.hidden_size)
        self.has_cross_attention = layer_idx % cfg.cross_attention_frequency == 0
        if self.has_cross_attention:
            self.crossattention = _AttentionBlock(cfg, cfg.encoder_hidden_size)
            self.text_crossattention = _AttentionBlock(cfg, cfg.text_encoder_hidden_size)
        self.intermediate = _Expand(cfg)
        self.output = _ResidualNorm(cfg, cfg.intermediate_size)
        self.intermediate_query = _Expand(cfg)
        self.output_query = _ResidualNorm(cfg, cfg.intermediate_size)

    def forward(self, x: torch.Tensor, visual: Optional[torch.Tensor], text: Optional[torch.Tensor],
                query_length: int) -> torch.Tensor:
        h = self.attention(x)
        if query_length <= 0:
            return self.output(self.intermediate(h), h)
        q = h[:, :query_length]
        if self.has_cross_attention:
            if visual is None:
                raise ValueError("encoder_hidden_states must be given for cross-attention layers")
            hc = self.crossattention(q, visual)
            ht = self.text_crossattention(hc, text)
            q = 0.4 * q + 0.3 * hc + 0.3 * ht
        out = self.output_query(self.intermediate_query(q), q)
        if h.shape[1] > query_length:
            rest = h[:, query_length:]
            out = torch.cat([out, self.output(self.intermediate(rest), rest)], dim=1)
        return out


class IQMEncoder(nn.Module):
    def __init__(self, cfg: IQMConfig):
        super().__init__()
        self.layer = nn.ModuleList([IQMLayer(cfg, i) for i in range(cfg.num_hidden_layers)])


class IQM(nn.Module):
    def __init__(self, config: IQMConfig):
        super().__init__()
        self.config = config
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.encoder = IQMEncoder(config)

    def forward(self, query_embeds: torch.Tensor, query_length: Optional[int] = None,
                encoder_hidden_states: Optional[torch.Tensor] = None,
                text_encoder_hidden_states: Optional[torch.Tensor] = None):
        """-> namespace with ``last_hidden_state`` [B, n_query, hidden] and ``pooler_output`` (row 0), the two fields of
        the reference's BaseModelOutputWithPoolingAndCrossAttentions that callers read (test_last.py:104-106)."""
        n_q = query_embeds.shape[1] if query_length is None else query_length
        x = self.dropout(self.layernorm(query_embeds))
        for layer in self.encoder.layer:
            x = layer(x, encoder_hidden_states, text_encoder_hidden_states, n_q)
        return SimpleNamespace(last_hidden_state=x, pooler_output=x[:, 0, :])
