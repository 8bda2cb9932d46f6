"""Loss / metric specification (device agnostic PyTorch).

Reference: ``network/ssim.py`` (11x11 Gaussian SSIM, zero padding), ``network/layers.py``
(``psnr``, ``edge_aware_loss`` v1 built on kornia's Sobel ``spatial_gradient``,
``edge_aware_loss_v2`` monodepth2-style), and the loss assembly in
``synthesis_task.py:295-351``.  kornia is not a dependency here: the normalised Sobel operator
with replicate padding is written out explicitly.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import torch
import torch.nn.functional as F

SSIM_WINDOW = 11
SSIM_SIGMA = 1.5
SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2


@lru_cache(maxsize=8)
def _gauss_1d(size: int, sigma: float) -> Tuple[float, ...]:
    g = [math.exp(-((i - size // 2) ** 2) / (2.0 * sigma * sigma)) for i in range(size)]
    s = sum(g)
    return tuple(v / s for v in g)


def gaussian_window(channels: int, size: int = SSIM_WINDOW, sigma: float = SSIM_SIGMA,
                    dtype=torch.float32, device=None) -> torch.Tensor:
    key = ("gauss", str(device), dtype, channels, size, sigma)
    if key not in _CONST_CACHE:
        g = torch.tensor(_gauss_1d(size, sigma), dtype=dtype)
        _CONST_CACHE[key] = torch.outer(g, g)[None, None].expand(channels, 1, size, size).contiguous().to(device)
    return _CONST_CACHE[key]


def ssim_map(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    c = a.shape[1]
    win = gaussian_window(c, dtype=a.dtype, device=a.device)
    pad = SSIM_WINDOW // 2
    blur = lambda t: F.conv2d(t, win, padding=pad, groups=c)
    mu_a, mu_b = blur(a), blur(b)
    var_a = blur(a * a) - mu_a * mu_a
    var_b = blur(b * b) - mu_b * mu_b
    cov = blur(a * b) - mu_a * mu_b
    num = (2 * mu_a * mu_b + SSIM_C1) * (2 * cov + SSIM_C2)
    den = (mu_a * mu_a + mu_b * mu_b + SSIM_C1) * (var_a + var_b + SSIM_C2)
    return num / den


def ssim(a: torch.Tensor, b: torch.Tensor, size_average: bool = True) -> torch.Tensor:
    m = ssim_map(a, b)
    return m.mean() if size_average else m.mean(dim=(1, 2, 3))


class SSIM(torch.nn.Module):
    """Module form with the reference signature ``SSIM(window_size=11, size_average=True)``."""

    def __init__(self, window_size: int = SSIM_WINDOW, size_average: bool = True):
        super().__init__()
        if window_size != SSIM_WINDOW:
            raise ValueError("only the 11x11 window used by MINE is supported")
        self.window_size = window_size
        self.size_average = size_average

    def forward(self, img1, img2):
        return ssim(img1, img2, self.size_average)


def psnr(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    mse = ((a - b) ** 2).mean(dim=(1, 2, 3))
    return (20.0 * torch.log10(1.0 / torch.sqrt(mse))).mean()


_CONST_CACHE = {}


def _sobel_kernels(device: str, dtype, normalized: bool) -> torch.Tensor:
    """Cached per device: creating constants from Python lists is a host->device copy, which is illegal
    inside CUDA-graph capture."""
    key = ("sobel", device, dtype, normalized)
    if key not in _CONST_CACHE:
        kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=dtype)
        if normalized:
            kx = kx / 8.0
        _CONST_CACHE[key] = torch.stack([kx, kx.t()])[:, None].to(device)
    return _CONST_CACHE[key]


def sobel_gradients(x: torch.Tensor, normalized: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel Sobel d/dx, d/dy with replicate padding (what kornia's ``spatial_gradient``
    computes; ``normalized`` divides the kernel by its L1 norm = 8)."""
    b, c, h, w = x.shape
    k = _sobel_kernels(str(x.device), x.dtype, bool(normalized))                  # 2,1,3,3
    xp = F.pad(x.reshape(b * c, 1, h, w), (1, 1, 1, 1), mode="replicate")
    g = F.conv2d(xp, k).reshape(b, c, 2, h, w)
    return g[:, :, 0], g[:, :, 1]


def edge_aware_loss(img: torch.Tensor, disp: torch.Tensor, gmin: float, grad_ratio: float) -> torch.Tensor:
    """Smoothness v1: hinge on instance-normalised |Sobel(disp)| away from image edges."""
    gx, gy = sobel_gradients(img, normalized=True)
    gx = gx.abs().sum(1, keepdim=True)
    gy = gy.abs().sum(1, keepdim=True)
    ex = (gx / (gx.amax(dim=(1, 2, 3), keepdim=True) * grad_ratio)).clamp(max=1.0)
    ey = (gy / (gy.amax(dim=(1, 2, 3), keepdim=True) * grad_ratio)).clamp(max=1.0)
    dx, dy = sobel_gradients(disp, normalized=False)
    dx = F.instance_norm(dx.abs()) - gmin
    dy = F.instance_norm(dy.abs()) - gmin
    return (dx.clamp(min=0) * (1.0 - ex) + dy.clamp(min=0) * (1.0 - ey)).mean()


def edge_aware_loss_v2(img: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """Smoothness v2: mean-normalised disparity forward differences weighted by exp(-|dI|)."""
    d = disp / (disp.mean(dim=(2, 3), keepdim=True) + 1e-7)
    ddx = (d[..., :, :-1] - d[..., :, 1:]).abs()
    ddy = (d[..., :-1, :] - d[..., 1:, :]).abs()
    idx = (img[..., :, :-1] - img[..., :, 1:]).abs().mean(1, keepdim=True)
    idy = (img[..., :-1, :] - img[..., 1:, :]).abs().mean(1, keepdim=True)
    return (ddx * torch.exp(-idx)).mean() + (ddy * torch.exp(-idy)).mean()


def masked_l1(syn: torch.Tensor, gt: torch.Tensor, mask_count: torch.Tensor, threshold: float) -> torch.Tensor:
    """|syn-gt| where at least ``threshold`` planes were visible; mean over *all* elements
    (reference ``synthesis_task.py:326-328``)."""
    valid = (mask_count >= threshold).to(syn.dtype)
    return ((syn - gt).abs() * valid).mean()


def log_disparity_l1(disp_syn: torch.Tensor, disp_gt: torch.Tensor, scale_factor: torch.Tensor) -> torch.Tensor:
    """mean |log(d_syn / scale) - log(d_gt)| over sparse points ``[B,1,N]``."""
    return (torch.log(disp_syn / scale_factor.reshape(-1, 1, 1)) - torch.log(disp_gt)).abs().mean()


def scale_factor_from_points(disp_syn: torch.Tensor, disp_gt: torch.Tensor) -> torch.Tensor:
    """``exp(mean(log d_syn - log d_gt))`` per image (reference ``compute_scale_factor``)."""
    return torch.exp((torch.log(disp_syn) - torch.log(disp_gt)).mean(dim=2)).squeeze(1)


def nearest_downsample(img: torch.Tensor, scale: int) -> torch.Tensor:
    """Image pyramid level: nearest-neighbour to (H/2^s, W/2^s) == strided slicing
    (``nn.Upsample(size=...)`` default mode, reference ``synthesis_task.py:129-133``)."""
    if scale == 0:
        return img
    f = 2 ** scale
    h, w = img.shape[-2] // f, img.shape[-1] // f
    return F.interpolate(img, size=(h, w), mode="nearest")
