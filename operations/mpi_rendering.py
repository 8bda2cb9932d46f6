"""Upstream-named functional API over ``mine_b200`` (B,S,C,H,W tensors)."""
import torch

from mine_b200.ops import api as _ops
from mine_b200.spec import render as _R
from mine_b200.spec import sampling as _S

render = _R.render
plane_volume_rendering = lambda rgb, sigma, xyz, is_bg_depth_inf=False: _R.render(rgb, sigma, xyz, False, is_bg_depth_inf)
weighted_sum_mpi = lambda rgb, xyz, weights, is_bg_depth_inf=False: _R.composite(rgb, xyz, weights, is_bg_depth_inf)
get_tgt_xyz_from_plane_disparity = lambda xyz_src, G_tgt_src: _R.transform_points(G_tgt_src, xyz_src)


def alpha_composition(alpha_BK1HW, value_BKCHW):
    w = _R.alpha_to_weights(alpha_BK1HW)
    return (value_BKCHW * w).sum(dim=1), w


def get_src_xyz_from_plane_disparity(meshgrid_src_homo, mpi_disparity_src, K_src_inv):
    return _R.src_plane_points(K_src_inv, mpi_disparity_src, meshgrid_src_homo.shape[1], meshgrid_src_homo.shape[2])


def get_xyz_from_depth(meshgrid_homo, depth, K_inv):
    return _R.src_rays(K_inv, meshgrid_homo.shape[1], meshgrid_homo.shape[2]) * depth


def render_tgt_rgb_depth(H_sampler, mpi_rgb_src, mpi_sigma_src, mpi_disparity_src, xyz_tgt_BS3HW, G_tgt_src,
                         K_src_inv, K_tgt, use_alpha=False, is_bg_depth_inf=False):
    """``xyz_tgt_BS3HW`` is accepted for signature parity; target-frame points are analytic here."""
    packed = _ops.pack_rgb_sigma(mpi_rgb_src, mpi_sigma_src)
    return _ops.render_tgt(packed, mpi_disparity_src, G_tgt_src, K_src_inv, K_tgt, use_alpha, is_bg_depth_inf)


def predict_mpi_coarse_to_fine(mpi_predictor, src_imgs, xyz_src_BS3HW_coarse, disparity_coarse_src, S_fine,
                               is_bg_depth_inf=False):
    if S_fine <= 0:
        return mpi_predictor(src_imgs, disparity_coarse_src), disparity_coarse_src
    with torch.no_grad():
        coarse = mpi_predictor(src_imgs, disparity_coarse_src)[0]
        _, w = _R.sigma_to_weights(coarse[:, :, 3:], xyz_src_BS3HW_coarse)
        disparity_all = _S.refine_disparity(disparity_coarse_src, w, S_fine)
    return mpi_predictor(src_imgs, disparity_all), disparity_all


def disparity_consistency_src_to_tgt(meshgrid_homo, K_src_inv, disparity_src, G_tgt_src, K_tgt, disparity_tgt):
    B, _, H, W = disparity_src.shape
    xyz_src = get_xyz_from_depth(meshgrid_homo, torch.reciprocal(disparity_src), K_src_inv)
    xyz_tgt = _R.transform_points(G_tgt_src, xyz_src[:, None])[:, 0].reshape(B, 3, -1)
    pxpy = _S.project_points(K_tgt, xyz_tgt)
    ok = (pxpy[:, 0:1] >= 0) & (pxpy[:, 0:1] <= W - 1) & (pxpy[:, 1:2] >= 0) & (pxpy[:, 1:2] <= H - 1)
    diff = (torch.reciprocal(xyz_tgt[:, 2:]) - _S.gather_nearest(disparity_tgt, pxpy)).abs()
    return diff[ok].mean()
