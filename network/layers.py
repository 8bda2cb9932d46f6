"""Upstream path of the smoothness / PSNR / VGG losses (reference ``network/layers.py``): ``mine_b200/spec/losses.py``
(CUDA: ``ops/csrc/smooth.cu``) and ``mine_b200/models/geometry_layers.py``."""
from mine_b200.spec.losses import edge_aware_loss, edge_aware_loss_v2, psnr  # noqa: F401
from mine_b200.models.geometry_layers import VGGPerceptualLoss  # noqa: E402,F401
