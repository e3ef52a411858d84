"""Reference import path ``utils`` (test_last.py:13, train.py:13): the two helpers the evaluation / training scripts
use. The augmentation helpers of the reference's utils.py (kornia flips, affine warps, noise) belong to its training
data pipeline and are not provided."""
import os
import random

import numpy as np
import torch


def setup_seed(seed: int) -> None:
    """Seeds python / numpy / torch (CPU and every GPU) and asks cuDNN for deterministic kernels (utils.py:10-20)."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    os.environ["CUBLAS_WORKSPACE_CONFIG"] = ":4096:8"
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def cos_sim(a_norm: torch.Tensor, b_norm: torch.Tensor) -> torch.Tensor:
    """b @ a^T for a matrix of unit rows, b @ a for a single unit vector (utils.py:88-95)."""
    if a_norm.dim() == 2:
        return b_norm @ a_norm.transpose(1, 0)
    if a_norm.dim() == 1:
        return b_norm @ a_norm
    raise NotImplementedError
