"""Reference import path ``model.clip`` (model/clip.py) -> aa-clip-iqm_b200/model/clip.py."""
from aaclip_b200.model.clip import (  # noqa: F401
    MODEL_CONFIGS, _MODEL_CKPT_PATHS, ClipWeights, create_model, resize_pos_embed)
