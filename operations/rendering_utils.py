"""Upstream-named sampling helpers."""
import torch

from mine_b200.spec import sampling as _S

gather_pixel_by_pxpy = _S.gather_nearest
sample_pdf = _S.sample_pdf


def transform_G_xyz(G, xyz, is_return_homo=False):
    squeeze = G.dim() == 2
    G_, xyz_ = (G[None], xyz[None]) if squeeze else (G, xyz)
    out = G_ @ torch.cat([xyz_, torch.ones_like(xyz_[:, :1])], dim=1)
    out = out if is_return_homo else out[:, :3]
    return out


def uniformly_sample_disparity_from_bins(batch_size, disparity_np, device):
    return _S.stratified_disparity(batch_size, torch.as_tensor(disparity_np, dtype=torch.float32, device=device))


def uniformly_sample_disparity_from_linspace_bins(batch_size, num_bins, start, end, device):
    return _S.stratified_disparity_linspace(batch_size, num_bins, start, end, device=device)
