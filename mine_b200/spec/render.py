"""Pure-PyTorch *specification* of MPI rendering (device agnostic, differentiable).

This is the numerical contract of the framework: the CPU plumbing path executes it and every
sm_100a kernel in ``mine_b200/ops`` is tested against it.  Semantics follow SURVEY 2.7 /
reference ``operations/mpi_rendering.py`` and ``operations/homography_sampler.py``; the
implementation is organised differently on purpose:

* camera-frame points are never materialised through batched GEMMs - rays are ``K^-1 (u,v,1)``
  and plane points are ``ray * depth`` (reference builds B*S copies of the meshgrid and calls
  ``torch.matmul``, ``mpi_rendering.py:140-163``);
* the target-frame points sampled by ``grid_sample`` in the reference (7-channel warp,
  ``mpi_rendering.py:206-219``) are evaluated analytically at the clamped sample position
  (bilinear interpolation of an affine field is exact), so only rgb+sigma are gathered;
* the homography inverse is closed form (no retry loop / host sync).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .. import geometry as geo

LAST_PLANE_THICKNESS = 1.0e3      # delta of the farthest plane (mpi_rendering.py:48-52)
TRANSMITTANCE_EPS = 1.0e-6        # added inside the running product (mpi_rendering.py:58)
WEIGHT_SUM_EPS = 1.0e-5           # depth normalisation (mpi_rendering.py:80)
BG_DEPTH = 1000.0                 # background depth when is_bg_depth_inf (mpi_rendering.py:77)


# ----------------------------------------------------------------------------------------------
# geometry of the plane sweep
# ----------------------------------------------------------------------------------------------
def src_rays(k_src_inv: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """``K^-1 (u,v,1)`` for every pixel: ``[B,3,H,W]`` (z component is 1 for pinhole K)."""
    grid = geo.pixel_grid(h, w, dtype=k_src_inv.dtype, device=k_src_inv.device)      # 3,H,W
    return torch.einsum("bij,jhw->bihw", k_src_inv, grid)


def src_plane_points(k_src_inv: torch.Tensor, disparity: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """``xyz_src[b,s] = rays[b] / disparity[b,s]`` -> ``[B,S,3,H,W]``
    (reference ``get_src_xyz_from_plane_disparity``)."""
    rays = src_rays(k_src_inv, h, w)
    depth = torch.reciprocal(disparity)
    return rays[:, None] * depth[:, :, None, None, None]


def transform_points(g: torch.Tensor, xyz: torch.Tensor) -> torch.Tensor:
    """Apply ``[B,4,4]`` rigid transforms to ``[B,S,3,H,W]`` points
    (reference ``get_tgt_xyz_from_plane_disparity`` / ``transform_G_xyz``)."""
    r = g[:, :3, :3]
    t = g[:, :3, 3]
    return torch.einsum("bij,bsjhw->bsihw", r, xyz) + t[:, None, :, None, None]


# ----------------------------------------------------------------------------------------------
# compositing
# ----------------------------------------------------------------------------------------------
def _exclusive_cumprod(x: torch.Tensor, dim: int) -> torch.Tensor:
    ones = torch.ones_like(x.narrow(dim, 0, 1))
    return torch.cat([ones, torch.cumprod(x, dim=dim).narrow(dim, 0, x.shape[dim] - 1)], dim=dim)


def plane_thickness(xyz: torch.Tensor) -> torch.Tensor:
    """Euclidean distance between consecutive planes along each pixel ray, ``[B,S,1,H,W]``;
    the last plane gets 1e3."""
    d = torch.linalg.vector_norm(xyz[:, 1:] - xyz[:, :-1], dim=2, keepdim=True)
    last = torch.full_like(d[:, :1], LAST_PLANE_THICKNESS)
    return torch.cat([d, last], dim=1)


def sigma_to_weights(sigma: torch.Tensor, xyz: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Volume-rendering weights.  Returns ``(T_acc, w)`` both ``[B,S,1,H,W]`` with
    ``T_s = exp(-sigma_s delta_s)``, ``T_acc_s = prod_{j<s}(T_j + 1e-6)``, ``w_s = T_acc_s (1-T_s)``."""
    trans = torch.exp(-sigma * plane_thickness(xyz))
    t_acc = _exclusive_cumprod(trans + TRANSMITTANCE_EPS, dim=1)
    return t_acc, t_acc * (1.0 - trans)


def alpha_to_weights(alpha: torch.Tensor) -> torch.Tensor:
    """Over-compositing weights for the ``mpi.use_alpha`` variant (reference ``alpha_composition``)."""
    return alpha * _exclusive_cumprod(1.0 - alpha, dim=1)


def composite(rgb: torch.Tensor, xyz: torch.Tensor, weights: torch.Tensor,
              is_bg_depth_inf: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(sum_s w c, depth)`` with ``depth = sum w z / (sum w + 1e-5)`` (or the bg-inf form).
    Reference ``weighted_sum_mpi``."""
    w_sum = weights.sum(dim=1)
    rgb_out = (weights * rgb).sum(dim=1)
    wz = (weights * xyz[:, :, 2:3]).sum(dim=1)
    if is_bg_depth_inf:
        depth = wz + (1.0 - w_sum) * BG_DEPTH
    else:
        depth = wz / (w_sum + WEIGHT_SUM_EPS)
    return rgb_out, depth


def render(rgb: torch.Tensor, sigma: torch.Tensor, xyz: torch.Tensor, use_alpha: bool = False,
           is_bg_depth_inf: bool = False):
    """Dispatcher with the reference's return convention ``(rgb, depth, blend_weights, weights)``
    (reference ``mpi_rendering.render``).  With ``use_alpha`` the blend weights are zero
    (no source blending) and depth is the alpha-composited z."""
    if not use_alpha:
        t_acc, w = sigma_to_weights(sigma, xyz)
        rgb_out, depth = composite(rgb, xyz, w, is_bg_depth_inf)
        return rgb_out, depth, t_acc, w
    w = alpha_to_weights(sigma)
    rgb_out = (w * rgb).sum(dim=1)
    depth = (w * xyz[:, :, 2:3]).sum(dim=1)
    return rgb_out, depth, torch.zeros_like(rgb), w


def blend_with_source(rgb: torch.Tensor, t_acc: torch.Tensor, src_img: torch.Tensor) -> torch.Tensor:
    """``c'_s = T_acc_s * I_src + (1 - T_acc_s) * c_s`` (reference ``synthesis_task.py:267-268``)."""
    return t_acc * src_img[:, None] + (1.0 - t_acc) * rgb


def render_src(mpi_rgb: torch.Tensor, mpi_sigma: torch.Tensor, disparity: torch.Tensor,
               k_src_inv: torch.Tensor, src_img: Optional[torch.Tensor] = None,
               use_alpha: bool = False, is_bg_depth_inf: bool = False, blend: bool = True):
    """Source-view pass of the training graph (reference ``synthesis_task.py:249-275``).

    Returns dict with ``rgb`` (B,3,H,W), ``depth`` and ``disparity`` (B,1,H,W), ``mpi_rgb``
    (the possibly source-blended colours that the target pass must warp), ``t_acc``, ``weights``.
    """
    b, s, _, h, w = mpi_rgb.shape
    xyz = src_plane_points(k_src_inv, disparity, h, w)
    rgb_syn, depth, t_acc, weights = render(mpi_rgb, mpi_sigma, xyz, use_alpha, is_bg_depth_inf)
    out_rgb = mpi_rgb
    if blend and src_img is not None:
        out_rgb = blend_with_source(mpi_rgb, t_acc, src_img)
        rgb_syn, depth = composite(out_rgb, xyz, weights, is_bg_depth_inf)
    return {"rgb": rgb_syn, "depth": depth, "disparity": torch.reciprocal(depth),
            "mpi_rgb": out_rgb, "t_acc": t_acc, "weights": weights, "xyz": xyz}


# ----------------------------------------------------------------------------------------------
# target-view warp
# ----------------------------------------------------------------------------------------------
def tgt_sample_coords(disparity: torch.Tensor, g_tgt_src: torch.Tensor, k_src_inv: torch.Tensor,
                      k_tgt: torch.Tensor, h: int, w: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """For every target pixel and plane: continuous source pixel coordinates ``[B,S,H,W,2]`` and the
    in-FoV mask ``[B,S,H,W]`` (``-1 < x < W`` and ``-1 < y < H``).  No gradient flows through the
    homography (the reference inverts it under ``no_grad``)."""
    with torch.no_grad():
        depth = torch.reciprocal(disparity)
        h_ts = geo.plane_homography(k_tgt, k_src_inv, g_tgt_src, depth)          # B,S,3,3
        h_st = geo.inv3x3(h_ts)
        grid = geo.pixel_grid(h, w, dtype=h_st.dtype, device=h_st.device)         # 3,H,W
        p = torch.einsum("bsij,jhw->bshwi", h_st, grid)
        xy = p[..., :2] / p[..., 2:3]
        valid = (xy[..., 0] > -1) & (xy[..., 0] < w) & (xy[..., 1] > -1) & (xy[..., 1] < h)
    return xy, valid


def bilinear_border(src: torch.Tensor, xy: torch.Tensor) -> torch.Tensor:
    """``grid_sample(bilinear, padding_mode='border', align_corners=False)`` expressed in pixel
    coordinates: sample ``src [N,C,H,W]`` at continuous pixel-centre positions ``xy [N,h,w,2]``
    clamped to the image (reference ``homography_sampler.py:134-139``)."""
    n, c, hh, ww = src.shape
    gx = (xy[..., 0] + 0.5) / (ww * 0.5) - 1.0
    gy = (xy[..., 1] + 0.5) / (hh * 0.5) - 1.0
    return F.grid_sample(src, torch.stack([gx, gy], dim=-1), mode="bilinear",
                         padding_mode="border", align_corners=False)


def render_tgt(mpi_rgb: torch.Tensor, mpi_sigma: torch.Tensor, disparity: torch.Tensor,
               g_tgt_src: torch.Tensor, k_src_inv: torch.Tensor, k_tgt: torch.Tensor,
               use_alpha: bool = False, is_bg_depth_inf: bool = False):
    """Warp the S planes into the target camera and composite (reference
    ``render_tgt_rgb_depth``).  Returns ``(rgb B,3,H,W ; depth B,1,H,W ; mask B,1,H,W)`` where
    ``mask`` counts the planes whose sample fell inside the source image."""
    b, s, _, h, w = mpi_rgb.shape
    xy, valid = tgt_sample_coords(disparity, g_tgt_src, k_src_inv, k_tgt, h, w)

    planes = torch.cat([mpi_rgb, mpi_sigma], dim=2).reshape(b * s, 4, h, w)
    warped = bilinear_border(planes, xy.reshape(b * s, h, w, 2)).reshape(b, s, 4, h, w)
    rgb_t, sigma_t = warped[:, :, :3], warped[:, :, 3:4]

    # xyz in the target frame at the *clamped* sample position (== what bilinear sampling of the
    # affine xyz field returns, including under border padding)
    with torch.no_grad():
        xc = xy[..., 0].clamp(0, w - 1)
        yc = xy[..., 1].clamp(0, h - 1)
        pix = torch.stack([xc, yc, torch.ones_like(xc)], dim=2)                 # B,S,3,H,W
        rays = torch.einsum("bij,bsjhw->bsihw", k_src_inv, pix)
        xyz_src = rays * torch.reciprocal(disparity)[:, :, None, None, None]
        xyz_t = transform_points(g_tgt_src, xyz_src)

    sigma_t = torch.where(xyz_t[:, :, 2:3] >= 0, sigma_t, torch.zeros_like(sigma_t))
    rgb_syn, depth, _, _ = render(rgb_t, sigma_t, xyz_t, use_alpha, is_bg_depth_inf)
    mask = valid.to(mpi_rgb.dtype).sum(dim=1, keepdim=True)
    return rgb_syn, depth, mask


def render_novel_view(mpi_rgb, mpi_sigma, disparity, g_tgt_src, k_src_inv, k_tgt,
                      scale_factor: Optional[torch.Tensor] = None, use_alpha: bool = False,
                      is_bg_depth_inf: bool = False):
    """Reference ``SynthesisTask.render_novel_view`` semantics on top of :func:`render_tgt`."""
    if scale_factor is not None:
        g_tgt_src = geo.rescale_translation(g_tgt_src, scale_factor)
    rgb, depth, mask = render_tgt(mpi_rgb, mpi_sigma, disparity, g_tgt_src, k_src_inv, k_tgt,
                                  use_alpha, is_bg_depth_inf)
    return {"tgt_imgs_syn": rgb, "tgt_disparity_syn": torch.reciprocal(depth), "tgt_mask_syn": mask}
