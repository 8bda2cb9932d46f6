"""Sparse-point disparity supervision as one autograd node (kernels: ``csrc/sparse.cu``).

``sparse_point_loss(disp_map, K, xyz, scale=None) -> (loss, scale)`` is the reference's

    d      = gather_nearest(disp_map, project(K, xyz))                     synthesis_task.py:276-281, 316-318
    scale  = exp(mean_n(log d - log(1/z)))        (only where none is given) synthesis_task.py:211-220
    loss   = mean_{b,n} |log(d / scale_b) - log(1/z)|                       synthesis_task.py:310-323

with both gradient paths of the differentiable scale factor handled in the backward kernel.  The composition of
PyTorch ops in ``mine_b200.spec`` stays the default; ``MINE_B200_SPARSE=fused`` selects this node on CUDA tensors
(opt-in until it has been measured on hardware).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .conv_engine import _count, ext


class SparsePointLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp_map, k, xyz, scale):
        disp = disp_map.detach().float().contiguous()
        sc = scale.detach().float().contiguous() if scale is not None else None
        loss, scale_out, idx, d_syn, sgn = ext().sparse_point_fwd(disp, k.detach().float().contiguous(),
                                                                  xyz.detach().float().contiguous(), sc)
        _count()
        ctx.save_for_backward(idx, d_syn, sgn, scale_out)
        ctx.cfg = (tuple(disp_map.shape), scale is None, disp_map.dtype)
        if scale is not None:
            ctx.mark_non_differentiable(scale_out)          # the caller keeps using its own (differentiable) tensor
        return loss, scale_out

    @staticmethod
    def backward(ctx, g_loss, g_scale):
        idx, d_syn, sgn, scale = ctx.saved_tensors
        shape, computed, dtype = ctx.cfg
        gl = g_loss.detach().float().reshape(1).contiguous()
        gs = g_scale.detach().float().contiguous() if (computed and g_scale is not None) else None
        grad_disp, grad_scale = ext().sparse_point_bwd(gl, gs, idx, d_syn, sgn, scale, list(shape), computed)
        _count(2)
        return grad_disp.to(dtype), None, None, (None if computed else grad_scale)


def sparse_point_loss(disp_map: torch.Tensor, k: torch.Tensor, xyz: torch.Tensor,
                      scale: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    loss, scale_out = SparsePointLoss.apply(disp_map, k, xyz, scale)
    return loss, (scale if scale is not None else scale_out)
