"""The two adapter building blocks with the reference's names and state_dict keys (model/adapter_modules.py:6-26):
``SimpleAdapter.fc = Sequential(Linear(bias=False), LeakyReLU)`` -> key ``fc.0.weight``;
``SimpleProj.fc = Linear(bias=False)`` -> ``fc.weight``, or the Sequential form (``fc.0.weight``) when relu=True.

At inference the engine reads their weights and runs the matmul + LeakyReLU inside its GEMM epilogues
(EPI_LEAKY_F32 / EPI_F32); ``forward`` below is what the autograd (training) route and the host-side IQM branch call."""
from torch import nn


class SimpleAdapter(nn.Module):
    def __init__(self, c_in, c_out=768):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c_in, c_out, bias=False), nn.LeakyReLU())

    def forward(self, x):
        return self.fc(x)


class SimpleProj(nn.Module):
    def __init__(self, c_in, c_out=768, relu=True):
        super().__init__()
        layer = nn.Linear(c_in, c_out, bias=False)
        self.fc = nn.Sequential(layer, nn.LeakyReLU()) if relu else layer

    def forward(self, x):
        return self.fc(x)
