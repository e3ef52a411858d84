"""gaussian_blur2d restated from kornia 0.6.9's published behaviour: normalised taps exp(-x^2/(2 sigma^2)) on
x = -(k-1)/2 .. (k-1)/2, 2-D kernel = outer product, reflect border of k//2, depthwise correlation.
PARITY UNPINNED against real kornia (not installable here); cross-checked against torchvision's gaussian_blur."""
import torch
import torch.nn.functional as F


def _taps(k: int, sigma: float, device, dtype):
    x = torch.arange(k, device=device, dtype=dtype) - k // 2
    if k % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * float(sigma) ** 2))
    return g / g.sum()


def gaussian_blur2d(x, kernel_size, sigma, border_type="reflect", separable=False):
    assert border_type == "reflect"
    ky, kx = kernel_size
    wy = _taps(ky, sigma[0], x.device, x.dtype)
    wx = _taps(kx, sigma[1], x.device, x.dtype)
    k2 = torch.matmul(wy.unsqueeze(-1), wx.unsqueeze(-1).t())
    c = x.shape[1]
    xp = F.pad(x, (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect")
    return F.conv2d(xp, k2.expand(c, 1, ky, kx), groups=c)
