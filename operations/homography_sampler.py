"""Upstream-named homography sampler facade."""
import numpy as np
import torch

from mine_b200 import geometry as _geo
from mine_b200.spec import render as _R


class HomographySample:
    def __init__(self, H_tgt, W_tgt, device=None):
        self.device = torch.device("cpu") if device is None else device
        self.Height_tgt, self.Width_tgt = H_tgt, W_tgt
        self.meshgrid = _geo.pixel_grid(H_tgt, W_tgt, device=self.device)
        self.n = torch.tensor([0.0, 0.0, 1.0], device=self.device)

    @staticmethod
    def euler_to_rotation_matrix(x_angle, y_angle, z_angle, seq="xyz", degrees=False):
        from scipy.spatial.transform import Rotation
        return Rotation.from_euler(seq, [-x_angle, -y_angle, -z_angle], degrees=degrees).as_matrix().astype(np.float32)

    def sample(self, src_BCHW, d_src_B, G_tgt_src, K_src_inv, K_tgt):
        """Warp ``src_BCHW`` lying on the fronto-parallel plane at depth ``d_src_B`` into the target
        camera; returns ``(tgt_BCHW, valid_mask BxHxW)``."""
        disparity = torch.reciprocal(d_src_B.reshape(-1, 1).to(src_BCHW.dtype))
        xy, valid = _R.tgt_sample_coords(disparity, G_tgt_src, K_src_inv, K_tgt, self.Height_tgt, self.Width_tgt)
        return _R.bilinear_border(src_BCHW, xy[:, 0]), valid[:, 0]
