"""Mirror of the reference's ``AdaptedCLIP`` (model/adapter.py:10-304) for inference: same constructor arguments,
sub-module names and state_dict keys (``image_adapter.{layer_adapters,seg_proj,det_proj}``, ``text_adapter``,
``clipmodel``, ``image_encoder``), same ``forward`` / ``encode_text`` signatures and return values -- so that the
reference's evaluation loop (test_last.get_predictions, test_last.py:53-158) and its checkpoint loading
(test_last.py:231-247) run unchanged against this class -- with every tensor-core contraction and fused op executed
by libaaclip_b200.so (sm_100a) through the C ABI.

What is NOT here: the IQM Q-Former branch (random, un-checkpointed weights at test time) -- its two
inputs are exported instead and an optional user callable ``iqm_branch`` may consume them -- and the training path.
There is no PyTorch / CPU fallback: without a CUDA sm_100a device or the built library, forward raises.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
from torch import nn

from .. import _lib
from .. import engine as _engine
from .adapter_modules import SimpleAdapter, SimpleProj


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
        for mod in (self.image_adapter, self.text_adapter):          # model/adapter.py:107-113
            for p in mod.parameters():
                if p.dim() > 1:
                    nn.init.xavier_uniform_(p)
        # IQM branch: hyper-parameters kept for signature compatibility; see module docstring
        self.iqm_hidden_size = iqm_hidden_size
        self.iqm_num_layers, self.iqm_num_heads = iqm_num_layers, iqm_num_heads
        self.iqm_branch: Optional[Callable] = None   # callable(tokens[4x[B,P,W]], cls[B,W], text_embeddings) -> outputs
        self._operand_dtype, self._max_batch = operand_dtype, max_batch
        self._vis: Optional[_engine.AaclipEngine] = None
        self._txt: Optional[_engine.AaclipTextEngine] = None
        self._vis_stamp = self._txt_stamp = None

    # ------------------------------------------------------------------------------------------ engine plumbing
    def _device(self) -> torch.device:
        return self.image_encoder.conv1.weight.device

    @staticmethod
    def _stamp(tensors):
        return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)

    def _check_inference(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("AdaptedCLIP (aaclip_b200) is the inference hot path: call it under torch.no_grad(). "
                               "The training path (train.py) is outside this build (SURVEY.md section 8f rank 4).")
        if self._device().type != "cuda":
            raise RuntimeError("AdaptedCLIP (aaclip_b200) needs its parameters on a CUDA sm_100a device "
                               "(model.to('cuda')); there is no CPU path")

    def _visual_tensors(self):
        sd = {"image_encoder." + k: v for k, v in self.image_encoder.state_dict().items()
              if k != "proj" and "attn_pool" not in k}
        sd.update({"image_adapter." + k: v for k, v in self.image_adapter.state_dict().items()})
        return sd

    def _text_tensors(self):
        keep = ("token_embedding.", "positional_embedding", "transformer.", "ln_final.")
        sd = {"clipmodel." + k: v for k, v in self.clipmodel.state_dict().items() if k.startswith(keep)}
        sd.update({"text_adapter." + k: v for k, v in self.text_adapter.state_dict().items()})
        return sd

    def _visual_engine(self) -> _engine.AaclipEngine:
        sd = self._visual_tensors()
        stamp = self._stamp(sd.values())
        if self._vis is None or stamp != self._vis_stamp:     # first call, or weights were (re)loaded / moved
            W = self.image_encoder.conv1.weight.shape[0]
            ps = self.image_encoder.conv1.weight.shape[-1]
            n_tok = self.image_encoder.positional_embedding.shape[0]
            grid = int(round((n_tok - 1) ** 0.5))
            layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("image_encoder.transformer.resblocks."))
            if self._vis is not None:
                self._vis.close()
            self._vis = _engine.AaclipEngine(
                self._device(), img_size=grid * ps, patch=ps, width=W, heads=W // 64, layers=layers,
                mlp_dim=sd["image_encoder.transformer.resblocks.0.mlp.c_fc.weight"].shape[0], embed_dim=768,
                adapt_until=self.image_adapt_until, levels=tuple(self.levels),
                act=_lib.ACT_QUICK_GELU if getattr(self.clipmodel, "quick_gelu", False) else _lib.ACT_GELU_ERF,
                proj_relu=int(self.relu), adapt_weight=self.i_w, max_batch=self._max_batch,
                operand_dtype=self._operand_dtype)
            self._vis.load_state_dict(sd)
            if self._vis.missing:
                raise RuntimeError(f"{self._vis.missing} visual-path tensors missing from the wrapped CLIP model")
            self._vis_stamp = stamp
        return self._vis

    def _text_engine(self) -> _engine.AaclipTextEngine:
        sd = self._text_tensors()
        stamp = self._stamp(sd.values())
        if self._txt is None or stamp != self._txt_stamp:
            vocab, Wt = sd["clipmodel.token_embedding.weight"].shape
            layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("clipmodel.transformer.resblocks."))
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

    # ------------------------------------------------------------------------------------------ reference surface
    def forward(self, x, text_embeddings=None, iqm_hidden_size=None):
        """x [B,3,S,S] fp32 -> (seg_tokens: list of len(levels) x [B,P,768] unit-norm, det_token [B,768],
        iqm_outputs). model/adapter.py:137-271. iqm_outputs is None unless `self.iqm_branch` is set (then it receives
        the two tensors the reference's IQM branch consumes: ln_post'ed tokens per level and the final CLS row)."""
        self._check_inference()
        eng = self._visual_engine()
        want_iqm = text_embeddings is not None and self.iqm_branch is not None
        outs_seg, outs_det, outs_tok, outs_cls = [], [], [], []
        for i in range(0, x.shape[0], self._max_batch):              # larger batches run in max_batch chunks
            seg, det, tokens, cls = eng.forward(x[i:i + self._max_batch], want_tokens=want_iqm, want_cls=want_iqm)
            outs_seg.append(seg), outs_det.append(det), outs_tok.append(tokens), outs_cls.append(cls)
        if len(outs_seg) == 1:
            seg, det, tokens, cls = outs_seg[0], outs_det[0], outs_tok[0], outs_cls[0]
        else:
            seg = [torch.cat([s[l] for s in outs_seg]) for l in range(len(self.levels))]
            det = torch.cat(outs_det)
            tokens = [torch.cat([t[l] for t in outs_tok]) for l in range(len(self.levels))] if want_iqm else None
            cls = torch.cat(outs_cls) if want_iqm else None
        iqm_outputs = self.iqm_branch(tokens, cls, text_embeddings) if want_iqm else None
        return seg, det, iqm_outputs

    def encode_text(self, text, adapt_text=True):
        """int token ids [n,77] -> [n,768] (un-normalised), model/adapter.py:273-304."""
        if not adapt_text:
            raise NotImplementedError("encode_text(adapt_text=False) = the plain CLIP text tower with text_projection "
                                      "(model/model.py:190-201): not on the AA-CLIP hot path")
        self._check_inference()
        return self._text_engine().encode_text(text)

    def forward_original(self, x, modality="visual"):
        raise NotImplementedError("forward_original (un-adapted CLIP features, model/adapter.py:125-135) is only used "
                                  "by the training script and is outside the inference hot path")
