"""Upstream path of the SSIM loss (reference ``network/ssim.py:7-76``): ``mine_b200/spec/losses.py`` (CUDA: ``ops/csrc/losses.cu``)."""
from mine_b200.spec.losses import SSIM, gaussian_window, ssim  # noqa: F401
