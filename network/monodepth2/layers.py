"""The subset of monodepth2 layers MINE uses, re-expressed on our modules (reference ``network/monodepth2/layers.py``:
``ConvBlock`` :106-120, ``Conv3x3`` :123-138, ``upsample`` :198-201; the unused geometry utilities live in
``mine_b200/models/geometry_layers.py``)."""
import torch.nn as nn
import torch.nn.functional as F

from mine_b200.models.norm import BatchNorm


class Conv3x3(nn.Module):
    def __init__(self, in_channels, out_channels, use_refl=True):
        super().__init__()
        self.mode = "reflect" if use_refl else "constant"
        self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3)

    def forward(self, x):
        return self.conv(F.pad(x, (1, 1, 1, 1), mode=self.mode))


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.bn = BatchNorm(out_channels)

    def forward(self, x):
        return F.elu(self.bn(self.conv(x)))


def upsample(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")

from mine_b200.models.geometry_layers import (BackprojectDepth, Project3D, SSIM3x3 as SSIM, compute_depth_errors,  # noqa: E402,F401
                                              disp_to_depth, get_smooth_loss, get_translation_matrix,
                                              rot_from_axisangle, transformation_from_parameters)
