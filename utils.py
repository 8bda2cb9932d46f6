"""Upstream-named helpers (``from utils import restore_model, get_embedder, inverse, ...``)."""
from mine_b200.geometry import inv3x3, inv_affine4x4
from mine_b200.models.checkpoint import restore_model  # noqa: F401
from mine_b200.spec.embedder import get_embedder  # noqa: F401
from mine_b200.utils import (AverageMeter, disparity_normalization_vis, linspace_batch, run_shell_cmd,  # noqa: F401
                             run_shell_cmd_shell)


def inverse(matrices):
    """Batched inverse of 3x3 or rigid/affine 4x4 matrices - closed form, no host sync, no retries."""
    return inv3x3(matrices) if matrices.shape[-1] == 3 else inv_affine4x4(matrices)
