"""Camera / pose algebra without host synchronisation.

The reference inverts 3x3 / 4x4 matrices with ``torch.inverse`` wrapped in a retry loop that
calls ``torch.cuda.synchronize()`` and reads ``isnan().any()`` back to the host every time
(reference ``utils.py:96-117``, used at ``synthesis_task.py:208,244`` and
``operations/homography_sampler.py:112-113``).  Everything here is closed form, batched,
device-agnostic and sync-free, so a training step can be captured in a CUDA graph.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def inv3x3(m: torch.Tensor) -> torch.Tensor:
    """Adjugate inverse of ``[...,3,3]`` matrices: the columns of adj(M) are the cross products of its rows
    (6 small kernels on the GPU instead of a cuSOLVER call with a host sync)."""
    r0, r1, r2 = m[..., 0, :], m[..., 1, :], m[..., 2, :]
    c0 = torch.linalg.cross(r1, r2, dim=-1)
    c1 = torch.linalg.cross(r2, r0, dim=-1)
    c2 = torch.linalg.cross(r0, r1, dim=-1)
    det = (r0 * c0).sum(dim=-1)
    return torch.stack([c0, c1, c2], dim=-1) / det[..., None, None]


def inv_rigid(g: torch.Tensor) -> torch.Tensor:
    """Inverse of a rigid ``[...,4,4]`` transform ``[R t; 0 1]`` -> ``[R^T  -R^T t; 0 1]``."""
    r = g[..., :3, :3]
    t = g[..., :3, 3:]
    rt = r.transpose(-1, -2)
    top = torch.cat([rt, -rt @ t], dim=-1)
    bottom = torch.zeros_like(g[..., 3:, :])
    bottom[..., 0, 3] = 1
    return torch.cat([top, bottom], dim=-2)


def inv_affine4x4(g: torch.Tensor) -> torch.Tensor:
    """Inverse of ``[A t; 0 1]`` with a general (not necessarily orthonormal) 3x3 block.

    This is what the reference's ``inverse(G_src_tgt)`` computes for pose matrices
    (``synthesis_task.py:208``); COLMAP poses are rigid up to rounding, so this and
    :func:`inv_rigid` agree to ~1e-7, but we keep the general form for exact parity.
    """
    a_inv = inv3x3(g[..., :3, :3])
    t = g[..., :3, 3:]
    top = torch.cat([a_inv, -a_inv @ t], dim=-1)
    bottom = torch.zeros_like(g[..., 3:, :])
    bottom[..., 0, 3] = 1
    return torch.cat([top, bottom], dim=-2)


def scale_intrinsics(k: torch.Tensor, scale: int) -> torch.Tensor:
    """``K / 2**scale`` with ``K[2,2]`` reset to 1 (reference ``synthesis_task.py:238-241``;
    deliberately *no* half-pixel correction - checkpoints were trained with this)."""
    ks = k / float(2 ** scale)
    ks = ks.clone()
    ks[..., 2, 2] = 1.0
    return ks


_PYR_SCALES = {}


def intrinsics_pyramid(k: torch.Tensor, levels: int = 4) -> torch.Tensor:
    """``[levels,B,3,3]``: ``K / 2**s`` with ``K[2,2] = 1`` for every pyramid level, in two kernels."""
    key = (str(k.device), k.dtype, levels)
    if key not in _PYR_SCALES:      # cached per device: a host->device copy is illegal inside CUDA-graph capture
        _PYR_SCALES[key] = torch.tensor([1.0 / 2 ** s for s in range(levels)], dtype=k.dtype).to(k.device)
    ks = k[None] * _PYR_SCALES[key][:, None, None, None]
    ks[..., 2, 2] = 1.0
    return ks


def inv_intrinsics(k: torch.Tensor) -> torch.Tensor:
    """Inverse of an upper-triangular pinhole matrix (falls back to the general adjugate)."""
    return inv3x3(k)


def fov_intrinsics(h: int, w: int, fov_deg: float = 90.0, dtype=torch.float32, device=None) -> torch.Tensor:
    """Pinhole K for a horizontal field of view (video generator preset,
    reference ``visualizations/image_to_video.py:192-202``)."""
    f = w * 0.5 / math.tan(math.radians(fov_deg) * 0.5)
    return torch.tensor([[f, 0.0, w * 0.5], [0.0, f, h * 0.5], [0.0, 0.0, 1.0]], dtype=dtype, device=device)


def pixel_grid(h: int, w: int, dtype=torch.float32, device=None) -> torch.Tensor:
    """Homogeneous integer pixel coordinates ``[3,H,W]`` = (u, v, 1), u in [0,W-1]
    (reference ``operations/homography_sampler.py:24-33``)."""
    v, u = torch.meshgrid(torch.arange(h, dtype=dtype, device=device),
                          torch.arange(w, dtype=dtype, device=device), indexing="ij")
    return torch.stack([u, v, torch.ones_like(u)], dim=0)


def plane_homography(k_tgt: torch.Tensor, k_src_inv: torch.Tensor, g_tgt_src: torch.Tensor,
                     depth: torch.Tensor) -> torch.Tensor:
    """``H_tgt<-src = K_tgt (R + t n^T / d) K_src^-1`` for fronto-parallel planes n=(0,0,1).

    ``k_tgt``, ``k_src_inv``: ``[B,3,3]``; ``g_tgt_src``: ``[B,4,4]``; ``depth``: ``[B,S]``.
    Returns ``[B,S,3,3]``.  (reference ``operations/homography_sampler.py:101-108``; the
    reference's ``- t n^T / (-d)`` is ``+ t n^T / d``.)
    """
    r = g_tgt_src[:, None, :3, :3]                       # B,1,3,3
    t = g_tgt_src[:, None, :3, 3]                        # B,1,3
    tn = torch.zeros_like(r).expand(-1, depth.shape[1], -1, -1).clone()
    tn[..., :, 2] = t / depth[..., None]
    return k_tgt[:, None] @ (r + tn) @ k_src_inv[:, None]


def split_pose(g: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return g[..., :3, :3], g[..., :3, 3]


def rescale_translation(g_tgt_src: torch.Tensor, scale_factor: torch.Tensor) -> torch.Tensor:
    """Divide the translation by the per-image scale factor (``synthesis_task.py:439-442``)."""
    g = g_tgt_src.detach().clone()
    g[:, :3, 3] = g[:, :3, 3] / scale_factor.detach().reshape(-1, 1).to(g.dtype)
    return g
