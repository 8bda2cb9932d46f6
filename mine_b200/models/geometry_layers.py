"""Depth / pose utility layers of the monodepth2 family.

The reference vendors these in ``network/monodepth2/layers.py`` (``disp_to_depth``,
``transformation_from_parameters``, ``get_translation_matrix``, ``rot_from_axisangle``, ``BackprojectDepth``,
``Project3D``, ``get_smooth_loss``, a 3x3 ``SSIM``, ``compute_depth_errors``) without using them on the MINE
path.  They are provided for users who built on them; written from the maths, batched and
device-agnostic.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import geometry as geo


def disp_to_depth(disp: torch.Tensor, min_depth: float, max_depth: float):
    """Sigmoid disparity in [0,1] -> (scaled disparity, depth) with depth in [min_depth, max_depth]."""
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1.0 / scaled


def get_translation_matrix(t: torch.Tensor) -> torch.Tensor:
    """``[B,1,3]`` or ``[B,3]`` translation -> ``[B,4,4]``."""
    t = t.reshape(-1, 3)
    m = torch.eye(4, dtype=t.dtype, device=t.device).repeat(t.shape[0], 1, 1)
    m[:, :3, 3] = t
    return m


def rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """Axis-angle ``[B,1,3]`` -> ``[B,4,4]`` rotation (Rodrigues)."""
    v = vec.reshape(-1, 3)
    angle = v.norm(dim=1, keepdim=True)
    axis = v / (angle + 1e-7)
    c, s = torch.cos(angle)[:, 0], torch.sin(angle)[:, 0]
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    k = 1 - c
    rot = torch.zeros(v.shape[0], 4, 4, dtype=v.dtype, device=v.device)
    rot[:, 0, 0] = x * x * k + c; rot[:, 0, 1] = x * y * k - z * s; rot[:, 0, 2] = x * z * k + y * s
    rot[:, 1, 0] = x * y * k + z * s; rot[:, 1, 1] = y * y * k + c; rot[:, 1, 2] = y * z * k - x * s
    rot[:, 2, 0] = x * z * k - y * s; rot[:, 2, 1] = y * z * k + x * s; rot[:, 2, 2] = z * z * k + c
    rot[:, 3, 3] = 1
    return rot


def transformation_from_parameters(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool = False) -> torch.Tensor:
    r = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        r = r.transpose(1, 2)
        t = -t
    tm = get_translation_matrix(t)
    return r @ tm if invert else tm @ r


class BackprojectDepth(nn.Module):
    """Depth map -> homogeneous camera-frame points ``[B,4,H*W]``."""

    def __init__(self, batch_size: int, height: int, width: int):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        self.register_buffer("pix", geo.pixel_grid(height, width).reshape(1, 3, -1), persistent=False)

    def forward(self, depth: torch.Tensor, inv_K: torch.Tensor) -> torch.Tensor:
        b = depth.shape[0]
        cam = inv_K[:, :3, :3] @ self.pix.to(depth.dtype).expand(b, -1, -1)
        cam = depth.reshape(b, 1, -1) * cam
        return torch.cat([cam, torch.ones_like(cam[:, :1])], dim=1)


class Project3D(nn.Module):
    """Points + K + T -> normalised sampling grid ``[B,H,W,2]`` in [-1,1]."""

    def __init__(self, batch_size: int, height: int, width: int, eps: float = 1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points: torch.Tensor, K: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
        p = (K @ T)[:, :3, :] @ points
        pix = p[:, :2] / (p[:, 2:3] + self.eps)
        pix = pix.reshape(points.shape[0], 2, self.height, self.width).permute(0, 2, 3, 1)
        scale = torch.tensor([self.width - 1, self.height - 1], dtype=pix.dtype, device=pix.device)
        return (pix / scale - 0.5) * 2


def get_smooth_loss(disp: torch.Tensor, img: torch.Tensor) -> torch.Tensor:
    """Edge-aware first-order smoothness (no mean normalisation; see ``spec.losses.edge_aware_loss_v2``)."""
    ddx = (disp[..., :, :-1] - disp[..., :, 1:]).abs()
    ddy = (disp[..., :-1, :] - disp[..., 1:, :]).abs()
    idx = (img[..., :, :-1] - img[..., :, 1:]).abs().mean(1, keepdim=True)
    idy = (img[..., :-1, :] - img[..., 1:, :]).abs().mean(1, keepdim=True)
    return (ddx * torch.exp(-idx)).mean() + (ddy * torch.exp(-idy)).mean()


class SSIM3x3(nn.Module):
    """monodepth2's 3x3 average-pool SSIM *loss* map ``clamp((1 - SSIM) / 2, 0, 1)`` (reflection padded)."""

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        x, y = F.pad(x, (1, 1, 1, 1), mode="reflect"), F.pad(y, (1, 1, 1, 1), mode="reflect")
        mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
        sx = F.avg_pool2d(x * x, 3, 1) - mu_x ** 2
        sy = F.avg_pool2d(y * y, 3, 1) - mu_y ** 2
        sxy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
        n = (2 * mu_x * mu_y + c1) * (2 * sxy + c2)
        d = (mu_x ** 2 + mu_y ** 2 + c1) * (sx + sy + c2)
        return ((1 - n / d) / 2).clamp(0, 1)


def compute_depth_errors(gt: torch.Tensor, pred: torch.Tensor):
    """abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 between positive depth tensors."""
    thresh = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = [(thresh < 1.25 ** k).float().mean() for k in (1, 2, 3)]
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = ((gt - pred).abs() / gt).mean()
    sq_rel = ((gt - pred) ** 2 / gt).mean()
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


class VDRPredictor(nn.Module):
    """View-dependent radiance predictor placeholder: upstream ships it as an identity (its ``forward`` returns
    the input, ``view_dependent_radiance_predictor.py:45``) and never imports it."""

    def forward(self, mpi_rgb, *args, **kwargs):
        return mpi_rgb


class VGGPerceptualLoss(nn.Module):
    """L1 distance between VGG16 feature maps at 4 depths (upstream ``network/layers.py:9-45``, unused by the
    training loss).  Weights are read from ``MINE_VGG16_WEIGHTS`` (torchvision ``vgg16`` state dict); without
    them the module raises on construction instead of silently using random features."""

    def __init__(self, resize: bool = True, weights_path: str | None = None):
        super().__init__()
        import os
        import torchvision
        path = weights_path or os.environ.get("MINE_VGG16_WEIGHTS")
        if not path or not os.path.exists(path):
            raise FileNotFoundError("VGGPerceptualLoss needs local VGG16 weights (MINE_VGG16_WEIGHTS)")
        vgg = torchvision.models.vgg16(weights=None)
        vgg.load_state_dict(torch.load(path, map_location="cpu"))
        f = vgg.features
        self.blocks = nn.ModuleList([f[:4].eval(), f[4:9].eval(), f[9:16].eval(), f[16:23].eval()])
        for p in self.parameters():
            p.requires_grad_(False)
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        self.resize = resize

    def forward(self, syn: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        x, y = (syn - self.mean) / self.std, (gt - self.mean) / self.std
        if self.resize:
            x = F.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False)
            y = F.interpolate(y, size=(224, 224), mode="bilinear", align_corners=False)
        loss = 0.0
        for blk in self.blocks:
            x, y = blk(x), blk(y)
            loss = loss + F.l1_loss(x, y)
        return loss
