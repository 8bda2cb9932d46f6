"""Op front door used by the task/engine.

Every op has exactly two executors: the PyTorch *specification* (``mine_b200.spec``; CPU plumbing
path and test oracle) and the sm_100a kernel (``mine_b200.ops.cuda``).  CUDA tensors always go
to the kernels - if the extension is not built the call raises (no silent fallback) - unless
``MINE_B200_FORCE_SPEC=1`` is set for A/B debugging.

Layouts: MPIs travel *packed* as ``[B,S,H,W,4]`` (r,g,b,sigma interleaved: one 16-byte texel per
bilinear tap); the public ``B,S,4,H,W`` form is a zero-copy permuted view of it.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch

from ..spec import losses as L
from ..spec import render as R
from ..spec import sampling as S


def _use_kernels(t: torch.Tensor) -> bool:
    return t.is_cuda and os.environ.get("MINE_B200_FORCE_SPEC", "0") != "1"


def _cuda():
    from . import cuda as C           # raises with a clear message if the extension is missing
    return C


# ---- packing ---------------------------------------------------------------------------------
def pack_mpi(mpi_bs4hw: torch.Tensor) -> torch.Tensor:
    """``[B,S,4,H,W]`` -> packed ``[B,S,H,W,4]`` (a view when the input already is one)."""
    return mpi_bs4hw.permute(0, 1, 3, 4, 2)


def unpack_mpi(mpi_packed: torch.Tensor) -> torch.Tensor:
    return mpi_packed.permute(0, 1, 4, 2, 3)


def pack_rgb_sigma(rgb: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """Packed MPI from separate ``[B,S,3,H,W]`` / ``[B,S,1,H,W]`` tensors; zero-copy when both are
    channel slices of one packed tensor."""
    if (rgb.dim() == 5 and rgb.stride(2) == 1 and sigma.stride(2) in (1, 0) and rgb.stride(4) == 4
            and sigma.data_ptr() == rgb.data_ptr() + 3 * rgb.element_size() and rgb.stride() == sigma.stride()):
        b, s, _, h, w = rgb.shape
        return torch.as_strided(rgb, (b, s, h, w, 4), (rgb.stride(0), rgb.stride(1), rgb.stride(3), rgb.stride(4), 1))
    return torch.cat([rgb, sigma], dim=2).permute(0, 1, 3, 4, 2).contiguous()


# ---- rendering -------------------------------------------------------------------------------
def render_src(mpi: torch.Tensor, disparity: torch.Tensor, k_src_inv: torch.Tensor,
               src_img: Optional[torch.Tensor], use_alpha: bool = False, is_bg_depth_inf: bool = False,
               blend: bool = True) -> Dict[str, torch.Tensor]:
    """Source-view pass on a packed MPI.  Returns ``rgb, depth, disparity, mpi`` where ``mpi`` is the
    packed MPI the target pass must warp (source-blended colours when ``blend``)."""
    if _use_kernels(mpi):
        return _cuda().render_src(mpi, disparity, k_src_inv, src_img, use_alpha, is_bg_depth_inf, blend)
    u = unpack_mpi(mpi)
    out = R.render_src(u[:, :, :3], u[:, :, 3:], disparity, k_src_inv, src_img, use_alpha, is_bg_depth_inf, blend)
    blended = torch.cat([out["mpi_rgb"], u[:, :, 3:]], dim=2).permute(0, 1, 3, 4, 2)
    return {"rgb": out["rgb"], "depth": out["depth"], "disparity": out["disparity"], "mpi": blended,
            "weights": out["weights"], "t_acc": out["t_acc"]}


def render_tgt(mpi: torch.Tensor, disparity: torch.Tensor, g_tgt_src: torch.Tensor, k_src_inv: torch.Tensor,
               k_tgt: torch.Tensor, use_alpha: bool = False, is_bg_depth_inf: bool = False
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Target-view pass on a packed MPI -> ``(rgb B,3,H,W ; depth B,1,H,W ; mask B,1,H,W)``."""
    if _use_kernels(mpi):
        return _cuda().render_tgt(mpi, disparity, g_tgt_src, k_src_inv, k_tgt, use_alpha, is_bg_depth_inf)
    u = unpack_mpi(mpi)
    return R.render_tgt(u[:, :, :3], u[:, :, 3:], disparity, g_tgt_src, k_src_inv, k_tgt, use_alpha, is_bg_depth_inf)


def plane_weights_mean(mpi: torch.Tensor, disparity: torch.Tensor, k_src_inv: torch.Tensor,
                       is_bg_depth_inf: bool = False) -> torch.Tensor:
    """Mean compositing weight of every plane ``[B,S]`` (coarse-to-fine importance, no grad)."""
    with torch.no_grad():
        u = unpack_mpi(mpi)
        b, s, _, h, w = u.shape
        xyz = R.src_plane_points(k_src_inv, disparity, h, w)
        _, wts = R.sigma_to_weights(u[:, :, 3:], xyz)
        return wts.mean(dim=(2, 3, 4))


# ---- losses ----------------------------------------------------------------------------------
def ssim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if _use_kernels(a):
        return _cuda().ssim(a, b)
    return L.ssim(a, b)


def masked_l1(syn, gt, mask_count, threshold: float) -> torch.Tensor:
    if _use_kernels(syn):
        return _cuda().masked_l1(syn, gt, mask_count, float(threshold))
    return L.masked_l1(syn, gt, mask_count, threshold)


def edge_aware_loss(img, disp, gmin: float, grad_ratio: float) -> torch.Tensor:
    if _use_kernels(disp):
        return _cuda().edge_aware_loss(img, disp, float(gmin), float(grad_ratio))
    return L.edge_aware_loss(img, disp, gmin, grad_ratio)


def edge_aware_loss_v2(img, disp) -> torch.Tensor:
    if _use_kernels(disp):
        return _cuda().edge_aware_loss_v2(img, disp)
    return L.edge_aware_loss_v2(img, disp)


def psnr(a, b) -> torch.Tensor:
    return L.psnr(a, b)


def sparse_disparity(disp_map: torch.Tensor, k: torch.Tensor, xyz: torch.Tensor) -> torch.Tensor:
    """Synthesised disparity at the projections of sparse camera-frame points -> ``[B,1,N]``."""
    return S.gather_nearest(disp_map, S.project_points(k, xyz))


def sparse_point_loss(disp_map: torch.Tensor, k: torch.Tensor, xyz: torch.Tensor,
                      scale: Optional[torch.Tensor] = None, calibrate: bool = True):
    """``(mean |log(d / scale) - log(1/z)|, scale)`` at the projections of sparse camera-frame points; the scale is
    calibrated from these points when none is given (``calibrate=False``: ones - datasets with metric poses).
    On CUDA: one kernel per direction (``csrc/sparse.cu``; ``MINE_B200_SPARSE=spec`` selects the ~150-op composition of
    specification ops instead); off-GPU: the specification ops."""
    if scale is None and not calibrate:
        scale = torch.ones(xyz.shape[0], dtype=torch.float32, device=xyz.device)
    if _use_kernels(disp_map) and os.environ.get("MINE_B200_SPARSE", "fused") == "fused":
        from .sparse import sparse_point_loss as fused
        return fused(disp_map, k, xyz, scale)
    disp_gt = torch.reciprocal(xyz[:, 2:, :])
    disp_syn = sparse_disparity(disp_map, k, xyz)
    if scale is None:
        scale = L.scale_factor_from_points(disp_syn, disp_gt)
    return L.log_disparity_l1(disp_syn, disp_gt, scale), scale


def image_pyramid(img: torch.Tensor, levels: int = 4):
    return [L.nearest_downsample(img, s) for s in range(levels)]
