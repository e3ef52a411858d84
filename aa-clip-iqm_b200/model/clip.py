"""Mirror of the reference's model factory for the hot path (model/clip.py:84-202, model/model.py:396-427).

``create_model`` returns a ``ClipWeights`` object: an ``nn.Module`` that HOLDS the CLIP parameters under the
reference's state_dict keys (``visual.conv1.weight``, ``visual.transformer.resblocks.<i>.attn.in_proj_weight``,
``transformer.resblocks.<i>...``, ``token_embedding.weight``, ``ln_final.weight`` ...) and exposes the attributes the
callers read (``.visual``, ``.visual.image_size / grid_size``, ``.context_length``, ``.vocab_size``). The adapted towers
run inside the sm_100a engine (AdaptedCLIP); ``encode_image`` / ``encode_text`` (the UN-adapted CLIP towers, used by
the reference's scripts outside the hot path) are plain PyTorch ops on the held parameters. A reference ``CLIP``
instance can be passed to AdaptedCLIP instead -- only its parameters and those attributes are used.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Union

import torch
import torch.nn.functional as F
from torch import nn

# model/model_configs/ViT-L-14-336.json (the one configuration of the hot path)
MODEL_CONFIGS = {
    "ViT-L-14-336": dict(
        embed_dim=768,
        vision_cfg=dict(image_size=336, layers=24, width=1024, patch_size=14, heads=16, mlp_ratio=4.0),
        text_cfg=dict(context_length=77, vocab_size=49408, width=768, heads=12, layers=12, mlp_ratio=4.0),
    ),
}
# where pretrained="openai" looks for the checkpoint (model/clip.py:16); overridable, like the reference's table
_MODEL_CKPT_PATHS = {"ViT-L-14-336": os.path.join("model", "ViT-L-14-336px.pt")}


def _tree_set(root: nn.Module, dotted: str, tensor: torch.Tensor):
    """Registers `tensor` as parameter `dotted` (e.g. 'transformer.resblocks.3.ln_1.weight') under nested Modules."""
    *path, leaf = dotted.split(".")
    m = root
    for p in path:
        if not hasattr(m, p):
            m.add_module(p, nn.Module())
        m = getattr(m, p)
    m.register_parameter(leaf, nn.Parameter(tensor, requires_grad=False))


def _block_shapes(prefix: str, W: int, F_: int):
    return {
        prefix + "ln_1.weight": (W,), prefix + "ln_1.bias": (W,),
        prefix + "attn.in_proj_weight": (3 * W, W), prefix + "attn.in_proj_bias": (3 * W,),
        prefix + "attn.out_proj.weight": (W, W), prefix + "attn.out_proj.bias": (W,),
        prefix + "ln_2.weight": (W,), prefix + "ln_2.bias": (W,),
        prefix + "mlp.c_fc.weight": (F_, W), prefix + "mlp.c_fc.bias": (F_,),
        prefix + "mlp.c_proj.weight": (W, F_), prefix + "mlp.c_proj.bias": (W,),
    }


def clip_param_shapes(cfg: dict) -> Dict[str, tuple]:
    """state_dict key -> shape of the reference CLIP module for `cfg` (model/model.py:149-174)."""
    v, t, E = cfg["vision_cfg"], cfg["text_cfg"], cfg["embed_dim"]
    W, ps, S = v["width"], v["patch_size"], v["image_size"]
    n_tok = (S // ps) ** 2 + 1
    shapes = {
        "visual.class_embedding": (W,), "visual.positional_embedding": (n_tok, W), "visual.proj": (W, E),
        "visual.conv1.weight": (W, 3, ps, ps),
        "visual.ln_pre.weight": (W,), "visual.ln_pre.bias": (W,),
        "visual.ln_post.weight": (W,), "visual.ln_post.bias": (W,),
    }
    for i in range(v["layers"]):
        shapes.update(_block_shapes(f"visual.transformer.resblocks.{i}.", W, int(W * v["mlp_ratio"])))
    Wt = t["width"]
    for i in range(t["layers"]):
        shapes.update(_block_shapes(f"transformer.resblocks.{i}.", Wt, int(Wt * t["mlp_ratio"])))
    shapes.update({
        "token_embedding.weight": (t["vocab_size"], Wt), "positional_embedding": (t["context_length"], Wt),
        "ln_final.weight": (Wt,), "ln_final.bias": (Wt,), "text_projection": (Wt, E), "logit_scale": (),
    })
    return shapes


class ClipWeights(nn.Module):
    """Holds CLIP parameters under the reference's names. `init`: 'random' (reference-like scales) or 'empty'."""

    def __init__(self, cfg: dict, quick_gelu: bool = False, init: str = "random", seed: Optional[int] = None):
        super().__init__()
        self.cfg = cfg
        self.quick_gelu = bool(quick_gelu)
        v, t = cfg["vision_cfg"], cfg["text_cfg"]
        g = None
        if seed is not None:
            g = torch.Generator().manual_seed(seed)
        for key, shape in clip_param_shapes(cfg).items():
            if init == "empty":
                ten = torch.empty(shape)
            elif key.endswith(("ln_1.weight", "ln_2.weight", "ln_pre.weight", "ln_post.weight", "ln_final.weight")):
                ten = torch.ones(shape)
            elif key.endswith(".bias"):
                ten = torch.zeros(shape)
            elif key == "logit_scale":
                ten = torch.tensor(math.log(1 / 0.07))
            else:
                fan_in = shape[-1] if len(shape) == 2 else (int(torch.tensor(shape[1:]).prod()) if len(shape) > 2 else shape[0])
                std = {"token_embedding.weight": 0.02, "positional_embedding": 0.01}.get(key, fan_in ** -0.5)
                ten = torch.randn(shape, generator=g) * std
            _tree_set(self, key, ten)
        # attributes the reference's callers read
        self.visual.image_size = (v["image_size"], v["image_size"])
        self.visual.patch_size = (v["patch_size"], v["patch_size"])
        self.visual.grid_size = (v["image_size"] // v["patch_size"],) * 2
        self.visual.image_mean = (0.48145466, 0.4578275, 0.40821073)     # model/clip.py:187-188
        self.visual.image_std = (0.26862954, 0.26130258, 0.27577711)
        self.context_length = t["context_length"]
        self.vocab_size = t["vocab_size"]

    def forward(self, *a, **k):
        raise RuntimeError("ClipWeights holds the CLIP parameters for AdaptedCLIP (sm_100a engine); the joint "
                           "image/text forward of the reference's CLIP class is not part of the hot path.")

    # The two un-adapted CLIP entry points the reference's scripts call OUTSIDE the hot path: train.py stage 1 and
    # forward_original (encode_image), test_last.py:264-271 without a text-adapter checkpoint (encode_text). Plain
    # PyTorch ops on these parameters (model/torch_path.py); differentiable.
    def encode_image(self, image, out_layers, normalize: bool = False):
        """(ln_post(CLS) @ proj [B,E], [tokens [B,N,W] after block l for l in out_layers]) -- model/model.py:185-188."""
        from . import torch_path as tp
        pooled, tokens = tp.clip_encode_image(image, tp.named(self.visual), list(out_layers), self.quick_gelu)
        return (F.normalize(pooled, dim=-1) if normalize else pooled), tokens

    def encode_text(self, text, normalize: bool = False):
        """token ids [n,77] -> ln_final(EOT row) @ text_projection [n,E] -- model/model.py:190-201."""
        from . import torch_path as tp
        x = tp.text_tower(text, tp.named(self), (), 0.0, self.quick_gelu) @ self.text_projection
        return F.normalize(x, dim=-1) if normalize else x


def resize_pos_embed(state_dict: Dict[str, torch.Tensor], model, interpolation: str = "bicubic",
                     antialias: bool = True) -> None:
    """Rescales ``state_dict['visual.positional_embedding']`` to the grid of `model` in place, as the reference does
    when a 336-px checkpoint is loaded at 518 px (model/model.py:396-427): the class-token row is kept, the patch
    rows are viewed as a [1,W,g,g] image and resampled (bicubic, antialias, align_corners=False)."""
    old = state_dict.get("visual.positional_embedding", None)
    if old is None or not hasattr(model.visual, "grid_size"):
        return
    gs = model.visual.grid_size
    gh, gw = (gs, gs) if isinstance(gs, int) else tuple(gs)
    if gh * gw + 1 == old.shape[0]:
        return
    tok, img = old[:1], old[1:]
    og = int(math.sqrt(img.shape[0]))
    img = img.reshape(1, og, og, -1).permute(0, 3, 1, 2)
    img = F.interpolate(img, size=(gh, gw), mode=interpolation, antialias=antialias, align_corners=False)
    img = img.permute(0, 2, 3, 1).reshape(gh * gw, -1)
    state_dict["visual.positional_embedding"] = torch.cat([tok, img], dim=0)


def _load_checkpoint_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """A plain state_dict file, or a TorchScript archive of the OpenAI release (model/openai.py:56-65 tries jit first
    and falls back to a state dict); fp16 tensors are promoted to fp32 (create_model's default precision)."""
    try:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        sd = torch.load(path, map_location="cpu")
    if isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    return {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()
            if k not in ("input_resolution", "context_length", "vocab_size")}


def create_model(model_name: str, img_size: int, pretrained: Optional[str] = None, precision: str = "fp32",
                 device: Union[str, torch.device] = "cpu", jit: bool = False, force_quick_gelu: bool = False,
                 force_custom_text: bool = False, force_patch_dropout: Optional[float] = None,
                 force_image_size: Optional[int] = None, output_dict: Optional[bool] = None,
                 require_pretrained: bool = False, adapter: bool = False) -> ClipWeights:
    """Same signature and control flow as the reference's create_model (model/clip.py:84-202), for the hot-path
    configuration. pretrained='openai' loads ``_MODEL_CKPT_PATHS[model_name]`` at `img_size` (positional embedding
    resampled); otherwise random init at the config's image size unless `force_image_size` is given -- `img_size`
    alone is ignored on that route, exactly as in the reference (model/clip.py:159-161)."""
    import copy
    model_name = model_name.replace("/", "-")
    if model_name not in MODEL_CONFIGS:
        raise RuntimeError(f"Model config for {model_name} not found.")
    if precision != "fp32":
        raise ValueError("the reference scripts run fp32 parameters (model/clip.py:88); the engine picks its own "
                         "tensor-core operand dtype")
    if jit or force_custom_text:
        raise ValueError("jit / custom_text models are outside the hot path")
    cfg = copy.deepcopy(MODEL_CONFIGS[model_name])
    if pretrained and pretrained.lower() == "openai":
        cfg["vision_cfg"]["image_size"] = img_size
        path = _MODEL_CKPT_PATHS[model_name]
        if not path or not os.path.exists(path):
            raise RuntimeError(f"Pretrained weights (openai) not found for model {model_name}: {path}")
        sd = _load_checkpoint_state_dict(path)
        model = ClipWeights(cfg, quick_gelu=False, init="empty")   # erf-GELU: SURVEY.md section 0.1
        resize_pos_embed(sd, model)
        model.load_state_dict(sd, strict=True)
    else:
        if force_image_size is not None:
            cfg["vision_cfg"]["image_size"] = force_image_size
        model = ClipWeights(cfg, quick_gelu=force_quick_gelu)
        if pretrained:
            path = _MODEL_CKPT_PATHS[model_name]
            if not path or not os.path.exists(path):
                raise RuntimeError(f"Pretrained weights ({pretrained}) not found for model {model_name}.")
            sd = _load_checkpoint_state_dict(path)
            resize_pos_embed(sd, model)
            model.load_state_dict(sd, strict=True)
        elif require_pretrained:
            raise RuntimeError(f"Pretrained weights were required for (model: {model_name}, pretrained: "
                               f"{pretrained}) but not loaded.")
    return model.to(device)
