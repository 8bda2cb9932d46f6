"""Autograd wrappers over the sm_100a extension (``_mine_b200_cuda``).

Importing this module loads (building in-tree if necessary) the extension; there is no Python
fallback - on a GPU box a missing/broken extension is an error, not a silent slow path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import build as _build

try:
    _ext = _build.load()
except Exception as e:  # pragma: no cover
    raise RuntimeError(
        "mine_b200 CUDA extension is not available (python -m mine_b200.ops.build): %r" % (e,)) from e

from .counters import LAUNCHES  # noqa: E402  (kernels launched by this package; bench.py reports it)


def _count(n: int = 1) -> None:
    LAUNCHES["count"] += n


def _depth_mode(use_alpha: bool, bg_inf: bool, normalise_alpha: bool = False) -> int:
    if use_alpha and not normalise_alpha:
        return 2
    return 1 if bg_inf else 0


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.detach().to(torch.float32).contiguous()


# ---- source view ---------------------------------------------------------------------------------
class _RenderSrc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mpi, disparity, k_src_inv, src_img, use_alpha, blend, depth_mode):
        mpi_c, disp, kinv, img = _f32c(mpi), _f32c(disparity), _f32c(k_src_inv), _f32c(src_img)
        do_blend = bool(blend and (img is not None) and not use_alpha)
        rgb, depth, wsum, out_mpi = _ext.render_src_fwd(mpi_c, disp, kinv, img, use_alpha, do_blend, depth_mode, do_blend)
        _count()
        if not do_blend:
            out_mpi = mpi_c
        ctx.save_for_backward(mpi_c, disp, kinv, img if img is not None else torch.empty(0, device=mpi.device), depth, wsum)
        ctx.cfg = (use_alpha, do_blend, depth_mode, img is not None)
        ctx.mark_non_differentiable(wsum)
        return rgb, depth, out_mpi, wsum

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_out_mpi, _g_wsum):
        mpi, disp, kinv, img, depth, wsum = ctx.saved_tensors
        use_alpha, do_blend, depth_mode, has_img = ctx.cfg
        g_blend = _f32c(g_out_mpi)
        g_mpi = _ext.render_src_bwd(mpi, disp, kinv, img if has_img else None, depth, wsum, _f32c(g_rgb),
                                    _f32c(g_depth), g_blend if do_blend else None, use_alpha, do_blend, depth_mode)
        _count()
        if not do_blend and g_blend is not None:       # identity pass-through of the MPI
            g_mpi = g_mpi + g_blend
        return g_mpi, None, None, None, None, None, None


def render_src(mpi, disparity, k_src_inv, src_img, use_alpha=False, is_bg_depth_inf=False, blend=True) -> Dict:
    # reference quirk: with source blending enabled the source depth is always the normalised form,
    # also in alpha mode (synthesis_task.py:267-274 re-composites with weighted_sum_mpi)
    mode = _depth_mode(use_alpha, is_bg_depth_inf, normalise_alpha=bool(blend and src_img is not None))
    rgb, depth, out_mpi, _ = _RenderSrc.apply(mpi, disparity, k_src_inv, src_img if blend else None, bool(use_alpha),
                                              bool(blend), mode)
    return {"rgb": rgb, "depth": depth, "disparity": torch.reciprocal(depth), "mpi": out_mpi}


# ---- target view ---------------------------------------------------------------------------------
class _RenderTgt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mpi, disparity, g_tgt_src, k_src_inv, k_tgt, use_alpha, depth_mode):
        mpi_c, disp, g, kinv, kt = _f32c(mpi), _f32c(disparity), _f32c(g_tgt_src), _f32c(k_src_inv), _f32c(k_tgt)
        rgb, depth, mask, wsum = _ext.render_tgt_fwd(mpi_c, disp, g, kinv, kt, use_alpha, depth_mode)
        _count()
        ctx.save_for_backward(mpi_c, disp, g, kinv, kt, rgb, depth, wsum)
        ctx.cfg = (use_alpha, depth_mode)
        ctx.mark_non_differentiable(mask)
        return rgb, depth, mask

    @staticmethod
    def backward(ctx, g_rgb, g_depth, _g_mask):
        mpi, disp, g, kinv, kt, rgb, depth, wsum = ctx.saved_tensors
        use_alpha, depth_mode = ctx.cfg
        g_mpi = _ext.render_tgt_bwd(mpi, disp, g, kinv, kt, rgb, depth, wsum, _f32c(g_rgb), _f32c(g_depth), use_alpha,
                                    depth_mode)
        _count(2)        # zero-fill + scatter kernel
        return g_mpi, None, None, None, None, None, None


def render_tgt(mpi, disparity, g_tgt_src, k_src_inv, k_tgt, use_alpha=False, is_bg_depth_inf=False
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return _RenderTgt.apply(mpi, disparity, g_tgt_src, k_src_inv, k_tgt, bool(use_alpha),
                            _depth_mode(use_alpha, is_bg_depth_inf))


# ---- losses --------------------------------------------------------------------------------------
class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a_c, b_c = _f32c(a), _f32c(b)
        need = a.requires_grad
        total, partials = _ext.ssim_fwd(a_c, b_c, need)
        _count()
        if need:
            ctx.save_for_backward(a_c, b_c, partials)
        ctx.inv_n = 1.0 / a_c.numel()
        return total * ctx.inv_n

    @staticmethod
    def backward(ctx, g):
        a, b, partials = ctx.saved_tensors
        grad = _ext.ssim_bwd(a, b, partials, _f32c(g).reshape(1), ctx.inv_n)
        _count()
        return grad, None


def ssim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _SSIM.apply(a, b)


class _MaskedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, syn, gt, mask, thr):
        need = syn.requires_grad
        total, sign = _ext.masked_l1_fwd(_f32c(syn), _f32c(gt), _f32c(mask), float(thr), need)
        _count()
        if need:
            ctx.save_for_backward(sign)
        ctx.inv_n = 1.0 / syn.numel()
        return total * ctx.inv_n

    @staticmethod
    def backward(ctx, g):
        (sign,) = ctx.saved_tensors
        return sign * (g * ctx.inv_n), None, None, None


def masked_l1(syn, gt, mask_count, threshold: float) -> torch.Tensor:
    return _MaskedL1.apply(syn, gt, mask_count, threshold)


class _SmoothV1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, disp, gmin, ratio):
        need = disp.requires_grad
        out, stats, sob, hmap, hs = _ext.smooth_v1_fwd(_f32c(img), _f32c(disp), float(gmin), float(ratio), need)
        _count(2)
        if need:
            ctx.save_for_backward(sob, stats, hmap, hs)
        return out

    @staticmethod
    def backward(ctx, g):
        sob, stats, hmap, hs = ctx.saved_tensors
        grad = _ext.smooth_v1_bwd(sob, stats, hmap, hs, _f32c(g).reshape(1))
        _count()
        return None, grad, None, None


class _SmoothV2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, disp):
        need = disp.requires_grad
        out, sums, g, gd = _ext.smooth_v2_fwd(_f32c(img), _f32c(disp), need)
        _count(2)
        if need:
            ctx.save_for_backward(g, sums, gd)
        return out

    @staticmethod
    def backward(ctx, gout):
        g, sums, gd = ctx.saved_tensors
        grad = _ext.smooth_v2_bwd(g, sums, gd, _f32c(gout).reshape(1))
        _count()
        return None, grad


def edge_aware_loss(img, disp, gmin: float, grad_ratio: float) -> torch.Tensor:
    """Smoothness v1 (Sobel edge mask x hinge on instance-normalised |Sobel(disp)|): 2 launches (+1 backward)."""
    return _SmoothV1.apply(img, disp, gmin, grad_ratio)


def edge_aware_loss_v2(img, disp) -> torch.Tensor:
    """Smoothness v2 (mean-normalised first differences x exp(-|dI|)): 2 launches (+1 backward)."""
    return _SmoothV2.apply(img, disp)


# ---- optimizer -----------------------------------------------------------------------------------
def fused_adam_(p, g, m, v, hyper, beta1, beta2, eps, weight_decay) -> None:
    """``hyper``: device fp32 tensor ``[lr, step]`` (bias corrections are derived in-kernel)."""
    _ext.fused_adam(p, g, m, v, hyper, float(beta1), float(beta2), float(eps), float(weight_decay))
    _count()
