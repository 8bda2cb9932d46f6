from mine_b200.spec.losses import SSIM, gaussian_window, ssim  # noqa: F401
