"""Mirror of the reference's ``AdaptedCLIP`` (model/adapter.py:10-304): same constructor arguments, sub-module names
and state_dict keys (``image_adapter.{layer_adapters,seg_proj,det_proj}``, ``text_adapter``, ``iqm``,
``class_query_mlp``, ``query_adapters``, ``clipmodel``, ``image_encoder`` ...), same ``forward`` / ``encode_text`` /
``forward_original`` signatures and return values -- so that the reference's evaluation loop
(test_last.get_predictions, test_last.py:53-158), its checkpoint loading (test_last.py:231-247) and its training steps
(train.py:150,219) run against this class.

Where the work happens:
  * inference (``torch.no_grad()``, or nothing requires grad): the adapted visual tower, the adapted text tower and
    every fused op run in libaaclip_b200.so (sm_100a) through the C ABI. There is NO PyTorch / CPU fallback on this
    route: without a CUDA sm_100a device or the built library it raises.
  * the fork's IQM branch (model/adapter.py:186-269) is host-side PyTorch on the tensors the
    engine exports (ln_post'ed tokens of the four levels, the final CLS row). Its weights are random at test time in
    the reference too (they are never checkpointed).
  * training (autograd enabled and some parameter requires grad): plain PyTorch ops on the same parameters
    (model/torch_path.py) -- the route train.py needs; section 8f rank 4 lists native training kernels as future work.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib
from .. import engine as _engine
from . import torch_path as _tp
from .adapter_modules import SimpleAdapter, SimpleProj
from .iqm import IQM, FactoredMemory, IQMConfig


class AdaptedCLIP(nn.Module):
    def __init__(self, clip_model, text_adapt_weight: float = 0.1, image_adapt_weight: float = 0.1,
                 text_adapt_until: int = 3, image_adapt_until: int = 6, levels: list = [6, 12, 18, 24],
                 relu: bool = True, iqm_hidden_size: int = 768, iqm_num_layers: int = 2, iqm_num_heads: int = 8,
                 operand_dtype: int = _lib.DT_F16, max_batch: int = 64, **kwargs):
        super().__init__()
        self.clipmodel = clip_model
        self.image_encoder = clip_model.visual
        self.text_adapt_until = text_adapt_until
        self.image_adapt_until = image_adapt_until
        self.t_w = text_adapt_weight
        self.i_w = image_adapt_weight
        self.levels = list(levels)
        self.relu = bool(relu)
        W = clip_model.visual.conv1.weight.shape[0]
        Wt = clip_model.token_embedding.weight.shape[1]
        E = 768
        self.image_adapter = nn.ModuleDict({
            "layer_adapters": nn.ModuleList([SimpleAdapter(W, W) for _ in range(image_adapt_until)]),
            "seg_proj": nn.ModuleList([SimpleProj(W, E, relu) for _ in range(len(levels))]),
            "det_proj": SimpleProj(W, E, relu),
        })
        self.text_adapter = nn.ModuleList([SimpleAdapter(Wt, Wt) for _ in range(text_adapt_until)]
                                          + [SimpleProj(Wt, Wt, relu=True)])
        # ---- IQM branch (model/adapter.py:56-96)
        self.iqm_config = IQMConfig(hidden_size=iqm_hidden_size, num_hidden_layers=iqm_num_layers,
                                    num_attention_heads=iqm_num_heads, encoder_hidden_size=iqm_hidden_size,
                                    text_encoder_hidden_size=768)
        self.iqm = IQM(self.iqm_config)
        self.class_query_mlp = nn.Sequential(nn.Linear(W, iqm_hidden_size), nn.ReLU(),
                                             nn.Linear(iqm_hidden_size, iqm_hidden_size))
        self.query_adapters = nn.ModuleList([SimpleProj(W, iqm_hidden_size, relu) for _ in range(len(levels))])
        self.iqm_hidden_size = iqm_hidden_size
        self.visual_feature_proj = None       # created on first use, like the reference (model/adapter.py:215-218)
        self.text_feature_proj = None         # (model/adapter.py:242-243)
        self.pos_embedding = nn.Parameter(_sinusoid(512, iqm_hidden_size))
        self.visual_weight = nn.Parameter(torch.tensor(0.6))    # held for the key set; the reference computes the
        self.text_weight = nn.Parameter(torch.tensor(0.4))      # normalised pair and never uses it (adapter.py:236-244)
        self.iqm_dropout = nn.Dropout(0.1)
        self.iqm_layer_norm = nn.LayerNorm(iqm_hidden_size)
        self.iqm_factored = True              # cross-attention over the visual memory in factored form (model/iqm.py)
        for mod in (self.image_adapter, self.text_adapter, self.iqm, self.class_query_mlp, self.query_adapters):
            for p in mod.parameters():                                # model/adapter.py:107-123
                if p.dim() > 1:
                    nn.init.xavier_uniform_(p)
        self._operand_dtype, self._max_batch = operand_dtype, max_batch
        self._vis: Optional[_engine.AaclipEngine] = None
        self._txt: Optional[_engine.AaclipTextEngine] = None
        self._vis_stamp = self._txt_stamp = None
        self._vis_list = self._txt_list = None      # cached (key, tensor) lists the stamps are taken over

    # ------------------------------------------------------------------------------------------ engine plumbing
    def _device(self) -> torch.device:
        return self.image_encoder.conv1.weight.device

    def _training_route(self) -> bool:
        """True when the call must be differentiable: autograd on and at least one parameter wants a gradient."""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _need_cuda(self):
        if self._device().type != "cuda":
            raise RuntimeError("AdaptedCLIP (aaclip_b200) inference needs its parameters on a CUDA sm_100a device "
                               "(model.to('cuda')); there is no CPU path")

    @staticmethod
    def _stamp(pairs):
        return tuple((t.data_ptr(), t._version) for _, t in pairs)

    def _visual_pairs(self):
        owner = (id(self.image_encoder), id(self.image_adapter), id(self.query_adapters))
        if self._vis_list is None or self._vis_list[0] != owner:
            sd = {"image_encoder." + k: v for k, v in _tp.named(self.image_encoder).items()
                  if k != "proj" and "attn_pool" not in k}
            sd.update({"image_adapter." + k: v for k, v in _tp.named(self.image_adapter).items()})
            if self._engine_query_dim():
                sd.update({"query_adapters." + k: v for k, v in _tp.named(self.query_adapters).items()})
            self._vis_list = (owner, list(sd.items()))
        return self._vis_list[1]

    def _text_pairs(self):
        owner = (id(self.clipmodel), id(self.text_adapter))
        if self._txt_list is None or self._txt_list[0] != owner:
            keep = ("token_embedding.", "positional_embedding", "transformer.", "ln_final.")
            sd = {"clipmodel." + k: v for k, v in _tp.named(self.clipmodel).items() if k.startswith(keep)}
            sd.update({"text_adapter." + k: v for k, v in _tp.named(self.text_adapter).items()})
            self._txt_list = (owner, list(sd.items()))
        return self._txt_list[1]

    def _engine_query_dim(self) -> int:
        """Width of the query adapters when the engine can run them (a multiple of 256: 512 and 768 both are), else 0
        (then the IQM branch projects the exported fp32 tokens on the host)."""
        return self.iqm_hidden_size if self.iqm_hidden_size % 256 == 0 else 0

    def refresh_engines(self):
        """Forget the packed copies (call after replacing a sub-module object; in-place updates -- load_state_dict,
        .to(), optimizer steps -- are noticed automatically through the tensors' version counters)."""
        self._vis_list = self._txt_list = self._vis_stamp = self._txt_stamp = None

    def _visual_engine(self) -> _engine.AaclipEngine:
        pairs = self._visual_pairs()
        stamp = self._stamp(pairs)
        if self._vis is None or stamp != self._vis_stamp:     # first call, or weights were (re)loaded / moved
            sd = dict(pairs)
            W = self.image_encoder.conv1.weight.shape[0]
            ps = self.image_encoder.conv1.weight.shape[-1]
            n_tok = self.image_encoder.positional_embedding.shape[0]
            grid = int(round((n_tok - 1) ** 0.5))
            layers = _tp.n_blocks(sd, "image_encoder.transformer.resblocks.")
            if self._vis is not None:
                self._vis.close()
            self._vis = _engine.AaclipEngine(
                self._device(), img_size=grid * ps, patch=ps, width=W, heads=W // 64, layers=layers,
                mlp_dim=sd["image_encoder.transformer.resblocks.0.mlp.c_fc.weight"].shape[0], embed_dim=768,
                adapt_until=self.image_adapt_until, levels=tuple(self.levels),
                act=_lib.ACT_QUICK_GELU if getattr(self.clipmodel, "quick_gelu", False) else _lib.ACT_GELU_ERF,
                proj_relu=int(self.relu), adapt_weight=self.i_w, max_batch=self._max_batch,
                operand_dtype=self._operand_dtype, query_dim=self._engine_query_dim())
            self._vis.load_state_dict(sd)
            if self._vis.missing:
                raise RuntimeError(f"{self._vis.missing} visual-path tensors missing from the wrapped CLIP model")
            self._vis_stamp = stamp
        return self._vis

    def _text_engine(self) -> _engine.AaclipTextEngine:
        pairs = self._text_pairs()
        stamp = self._stamp(pairs)
        if self._txt is None or stamp != self._txt_stamp:
            sd = dict(pairs)
            vocab, Wt = sd["clipmodel.token_embedding.weight"].shape
            layers = _tp.n_blocks(sd, "clipmodel.transformer.resblocks.")
            if self._txt is not None:
                self._txt.close()
            self._txt = _engine.AaclipTextEngine(
                self._device(), context=sd["clipmodel.positional_embedding"].shape[0], vocab=vocab, width=Wt,
                heads=Wt // 64, layers=layers, mlp_dim=sd["clipmodel.transformer.resblocks.0.mlp.c_fc.weight"].shape[0],
                embed_dim=Wt, adapt_until=self.text_adapt_until, adapt_weight=self.t_w,
                act=_lib.ACT_QUICK_GELU if getattr(self.clipmodel, "quick_gelu", False) else _lib.ACT_GELU_ERF,
                operand_dtype=self._operand_dtype)
            self._txt.load_state_dict(sd)
            if self._txt.missing:
                raise RuntimeError(f"{self._txt.missing} text-path tensors missing from the wrapped CLIP model")
            self._txt_stamp = stamp
        return self._txt

    # ------------------------------------------------------------------------------------------ the two routes
    def _tower_engine(self, x: torch.Tensor, want_iqm: bool):
        """-> (seg, det, tokens | None, cls | None, query_tokens | None). With the IQM branch wanted, the engine also
        runs the four query-adapter GEMMs and the fp32 tokens are not exported at all."""
        self._need_cuda()
        eng = self._visual_engine()
        on_device = want_iqm and self._engine_query_dim() > 0
        chunks = [eng.forward(x[i:i + self._max_batch], want_tokens=want_iqm and not on_device, want_cls=want_iqm,
                              want_query=True) if on_device else
                  eng.forward(x[i:i + self._max_batch], want_tokens=want_iqm, want_cls=want_iqm) + (None,)
                  for i in range(0, x.shape[0], self._max_batch)]      # larger batches run in max_batch chunks
        if len(chunks) == 1:
            return chunks[0]
        L = len(self.levels)
        seg = [torch.cat([c[0][l] for c in chunks]) for l in range(L)]
        det = torch.cat([c[1] for c in chunks])
        tokens = [torch.cat([c[2][l] for c in chunks]) for l in range(L)] if chunks[0][2] is not None else None
        cls = torch.cat([c[3] for c in chunks]) if want_iqm else None
        query = [torch.cat([c[4][l] for c in chunks]) for l in range(L)] if on_device else None
        return seg, det, tokens, cls, query

    def _tower_autograd(self, x: torch.Tensor):
        P = _tp.named(self.image_encoder)
        tokens, cls = _tp.visual_tower(x, P, self.image_adapter["layer_adapters"], self.i_w, self.levels,
                                       bool(getattr(self.clipmodel, "quick_gelu", False)))
        seg = [F.normalize(self.image_adapter["seg_proj"][i](t), dim=-1) for i, t in enumerate(tokens)]
        det = F.normalize(self.image_adapter["det_proj"](tokens[-1]), dim=-1).mean(1)
        return seg, det, tokens, cls, None

    # ------------------------------------------------------------------------------------------ IQM branch
    def _iqm_branch(self, tokens: Optional[List[torch.Tensor]], cls_feature: torch.Tensor, text_embeddings: torch.Tensor,
                    query_tokens: Optional[List[torch.Tensor]] = None):
        """model/adapter.py:186-269 on the tower's exports. tokens: ln_post'ed [B,P,W] per level; cls_feature: the
        final block's CLS row [B,W] (no ln_post); query_tokens: ``query_adapters[l](tokens[l])`` [B,P,hidden] when the
        engine has already computed them (then `tokens` is not needed).

        ``self.iqm_factored`` (default True): the projected visual memory proj(cat(levels)) [B, 4P, hidden] is never
        materialised and never pushed through the IQM's key / value projections -- the cross-attention runs in its
        factored form (model/iqm.py: FactoredMemory); False reproduces the reference's operation order literally."""
        B = cls_feature.shape[0]
        class_query = self.class_query_mlp(cls_feature).unsqueeze(1).repeat(1, 2, 1)      # normal / abnormal query
        query_embeds = class_query + self.pos_embedding[:, :2, :].repeat(B, 1, 1)
        if query_tokens is None:
            query_tokens = [self.query_adapters[i](t) for i, t in enumerate(tokens)]       # 4 x [B, P, hidden]
        width = int(query_tokens[0].shape[-1])
        if self.visual_feature_proj is None or self.visual_feature_proj.in_features != width:
            self.visual_feature_proj = nn.Linear(width, self.iqm_hidden_size).to(query_tokens[0].device)
        visual = FactoredMemory(query_tokens, self.visual_feature_proj)
        if not self.iqm_factored:
            visual = visual.materialise()
        te = text_embeddings
        if te.dim() == 2:                                   # [E,2] -> [B,2,E]
            te = te.transpose(0, 1).unsqueeze(0).repeat(B, 1, 1)
        elif te.dim() != 3:
            te = te.unsqueeze(0).repeat(B, 1, 1)
        # (a [B,E,2] stack, what test_last / train pass, goes in as is: 768 "tokens" of width 2 -- adapter.py:228-243)
        if self.text_feature_proj is None or self.text_feature_proj.in_features != te.shape[-1]:
            self.text_feature_proj = nn.Linear(int(te.shape[-1]), 768).to(te.device)
        out = self.iqm(query_embeds=query_embeds, query_length=query_embeds.shape[1], encoder_hidden_states=visual,
                       text_encoder_hidden_states=self.text_feature_proj(te))
        out.last_hidden_state = self.iqm_dropout(self.iqm_layer_norm(out.last_hidden_state))
        return out

    # ------------------------------------------------------------------------------------------ reference surface
    def forward(self, x, text_embeddings=None, iqm_hidden_size=None):
        """x [B,3,S,S] fp32 -> (seg_tokens: list of len(levels) x [B,P,768] unit-norm, det_token [B,768],
        iqm_outputs | None). model/adapter.py:137-271. ``iqm_outputs.last_hidden_state`` is [B,2,iqm_hidden_size]."""
        want_iqm = text_embeddings is not None
        if self._training_route():
            seg, det, tokens, cls, query = self._tower_autograd(x)
        else:
            seg, det, tokens, cls, query = self._tower_engine(x, want_iqm)
        iqm_outputs = self._iqm_branch(tokens, cls, text_embeddings, query_tokens=query) if want_iqm else None
        return seg, det, iqm_outputs

    def encode_text(self, text, adapt_text=True):
        """int token ids [n,77] -> [n,768] (un-normalised), model/adapter.py:273-304."""
        if not adapt_text:
            return self.clipmodel.encode_text(text)
        if self._training_route():
            P = _tp.named(self.clipmodel)
            eot = _tp.text_tower(text, P, list(self.text_adapter)[:self.text_adapt_until], self.t_w,
                                 bool(getattr(self.clipmodel, "quick_gelu", False)))
            return self.text_adapter[-1](eot)
        self._need_cuda()
        return self._text_engine().encode_text(text)

    def forward_original(self, x, modality="visual"):
        """Un-adapted CLIP features (model/adapter.py:125-135; training script only)."""
        if modality != "visual":
            raise ValueError("modality must be visual")
        cls_features, patch_features = self.clipmodel.encode_image(x, [24])
        patch_features = [t[:, 1:] for t in patch_features]
        ln = self.clipmodel.visual
        W = patch_features[0].shape[-1]
        patch_features = [F.layer_norm(t, (W,), ln.ln_post.weight, ln.ln_post.bias, 1e-5) @ ln.proj
                          for t in patch_features]
        return patch_features, cls_features


def _sinusoid(max_len: int, d_model: int) -> torch.Tensor:
    """[1, max_len, d_model] sine / cosine table (model/adapter.py:98-105)."""
    pos = torch.arange(max_len).unsqueeze(1)
    freq = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * freq)
    pe[:, 1::2] = torch.cos(pos * freq)
    return pe.unsqueeze(0)
