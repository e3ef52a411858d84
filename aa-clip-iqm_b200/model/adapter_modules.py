"""Parameter holders with the reference's names and state_dict keys (model/adapter_modules.py:6-26):
``SimpleAdapter.fc = Sequential(Linear(bias=False), LeakyReLU)`` -> key ``fc.0.weight``;
``SimpleProj.fc = Linear(bias=False)`` -> ``fc.weight``, or the Sequential form (``fc.0.weight``) when relu=True.
They only HOLD the weights (so reference checkpoints load unchanged); the matmul + LeakyReLU run inside the engine's
GEMM epilogues (EPI_LEAKY_F32 / EPI_F32). Calling one directly raises: there is no PyTorch compute path."""
from torch import nn


class _EngineOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} only holds weights; its math runs inside the sm_100a engine "
                           "(AdaptedCLIP.forward / encode_text). There is no PyTorch fallback.")


class SimpleAdapter(_EngineOnly):
    def __init__(self, c_in, c_out=768):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c_in, c_out, bias=False), nn.LeakyReLU())


class SimpleProj(_EngineOnly):
    def __init__(self, c_in, c_out=768, relu=True):
        super().__init__()
        if relu:
            self.fc = nn.Sequential(nn.Linear(c_in, c_out, bias=False), nn.LeakyReLU())
        else:
            self.fc = nn.Linear(c_in, c_out, bias=False)
