"""Batch normalisation with cross-replica statistics (SyncBN semantics) and a pluggable reducer.

The reference converts every BatchNorm to ``nn.SyncBatchNorm`` over all ranks
(``synthesis_task.py:106-113``): 67 layers -> 67 forward all_gathers + 67 backward all_reduces
of <=16 KiB per step.  Here one module owns both behaviours:

* statistics are reduced through ``reducer(tensor)`` - an in-place SUM all-reduce supplied by
  ``mine_b200.parallel`` (own NVLink one-shot kernel on CUDA, ``torch.distributed`` on CPU/gloo)
  - a single fused vector ``[sum(C), sumsq(C), count]`` forward and ``[sum_dy(C), sum_dy_xhat(C)]``
  backward;
* parameter / buffer names equal ``nn.BatchNorm2d`` (``weight, bias, running_mean, running_var,
  num_batches_tracked``) so reference checkpoints load unchanged.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn as nn

Reducer = Optional[Callable[[torch.Tensor], torch.Tensor]]


def _acc_dtype(t: torch.Tensor) -> torch.dtype:
    return torch.float64 if t.dtype == torch.float64 else torch.float32


class _BatchNormTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, reducer):
        c = x.shape[1]
        dims = [0] + list(range(2, x.dim()))
        acc = _acc_dtype(x)
        xf = x.to(acc)
        stats = torch.empty(2 * c + 1, dtype=acc, device=x.device)
        stats[:c] = xf.sum(dim=dims)
        stats[c:2 * c] = (xf * xf).sum(dim=dims)
        stats[2 * c] = float(x.numel() // c)
        if reducer is not None:
            stats = reducer(stats)
        n = stats[2 * c]
        mean = stats[:c] / n
        var = (stats[c:2 * c] / n - mean * mean).clamp_min(0.0)
        invstd = torch.rsqrt(var + eps)
        if running_mean is not None:
            with torch.no_grad():
                unbiased = var * (n / (n - 1.0).clamp_min(1.0))
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1 - momentum).add_(unbiased.to(running_var.dtype), alpha=momentum)
        shape = [1, c] + [1] * (x.dim() - 2)
        xhat = (xf - mean.view(shape)) * invstd.view(shape)
        y = xhat * weight.to(acc).view(shape) + bias.to(acc).view(shape)
        ctx.save_for_backward(x, weight, mean, invstd, n)
        ctx.reducer = reducer
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, invstd, n = ctx.saved_tensors
        c = x.shape[1]
        dims = [0] + list(range(2, x.dim()))
        shape = [1, c] + [1] * (x.dim() - 2)
        acc = _acc_dtype(x)
        dyf = dy.to(acc)
        xhat = (x.to(acc) - mean.view(shape)) * invstd.view(shape)
        red = torch.empty(2 * c, dtype=acc, device=x.device)
        red[:c] = dyf.sum(dim=dims)
        red[c:] = (dyf * xhat).sum(dim=dims)
        dbias = red[:c].clone()
        dweight = red[c:].clone()
        if ctx.reducer is not None:
            red = ctx.reducer(red)
        m_dy = (red[:c] / n).view(shape)
        m_dyx = (red[c:] / n).view(shape)
        dx = (dyf - m_dy - xhat * m_dyx) * (invstd * weight.to(acc)).view(shape)
        return dx.to(x.dtype), dweight.to(weight.dtype), dbias.to(weight.dtype), None, None, None, None, None


class _SyncBatchNormCUDA(torch.autograd.Function):
    """Cross-replica BN on CUDA built from ATen's fused batch_norm_* kernels (the ones
    ``nn.SyncBatchNorm`` uses) with OUR reducer in place of the all_gather / all_reduce collectives:
    one [2C] SUM forward (sum, sum of squares) and one [2C] SUM backward.  Used by the encoder
    (library convs, 4 % of the FLOPs); the decoder's BN lives inside the tcgen05 conv engine."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, reducer, world):
        c = x.shape[1]
        x = x.contiguous(memory_format=torch.channels_last) if x.dim() == 4 and not x.is_contiguous() else x
        mean_l, invstd_l = torch.batch_norm_stats(x, eps)
        n_l = float(x.numel() // c)
        var_l = (1.0 / (invstd_l * invstd_l) - eps).clamp_min(0.0)
        packed = torch.cat([mean_l * n_l, (var_l + mean_l * mean_l) * n_l])
        packed = reducer(packed)
        n = n_l * world
        mean = packed[:c] / n
        var = (packed[c:] / n - mean * mean).clamp_min(0.0)
        invstd = torch.rsqrt(var + eps)
        if running_mean is not None:
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1 - momentum).add_((var * (n / max(n - 1.0, 1.0))).to(running_var.dtype), alpha=momentum)
        y = torch.batch_norm_elemt(x, weight, bias, mean, invstd, eps)
        ctx.save_for_backward(x, weight, mean, invstd)
        ctx.reducer, ctx.n = reducer, n
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, invstd = ctx.saved_tensors
        dy = dy.contiguous(memory_format=torch.channels_last) if dy.dim() == 4 and not dy.is_contiguous() else dy
        sum_dy, sum_dy_xmu, gw, gb = torch.batch_norm_backward_reduce(dy, x, mean, invstd, weight, True, True, True)
        c = sum_dy.numel()
        packed = ctx.reducer(torch.cat([sum_dy, sum_dy_xmu]))
        count = torch.full((1,), ctx.n, dtype=torch.int32, device=x.device)
        dx = torch.batch_norm_backward_elemt(dy, x, mean, invstd, weight, packed[:c], packed[c:], count)
        return dx, gw, gb, None, None, None, None, None, None


class BatchNorm(nn.Module):
    """Drop-in for ``nn.BatchNorm2d`` / ``nn.SyncBatchNorm`` (same state-dict keys)."""

    defer_counters = False      # set by the trainer: ``num_batches_tracked`` of all layers advance in one launch per step

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.reducer: Reducer = None

    def scale_shift(self):
        """Eval-mode affine ``y = x * a + b`` (used by fused kernels and weight folding)."""
        acc = _acc_dtype(self.weight)
        a = self.weight.to(acc) * torch.rsqrt(self.running_var.to(acc) + self.eps)
        return a, self.bias.to(acc) - self.running_mean.to(acc) * a

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            if BatchNorm.defer_counters:             # the trainer bumps all counters with ONE multi-tensor op per step
                self._nbt_pending = True
            else:
                with torch.no_grad():
                    self.num_batches_tracked += 1
            if x.is_cuda and x.dtype != torch.float64:
                if self.reducer is None:        # single replica: ATen/cuDNN fused training BN
                    return torch.nn.functional.batch_norm(x, self.running_mean, self.running_var, self.weight,
                                                          self.bias, True, self.momentum, self.eps)
                world = int(getattr(getattr(self.reducer, "__self__", None), "world_size", 1))
                return _SyncBatchNormCUDA.apply(x, self.weight, self.bias, self.running_mean, self.running_var,
                                                self.momentum, self.eps, self.reducer, world)
            return _BatchNormTrain.apply(x, self.weight, self.bias, self.running_mean, self.running_var,
                                         self.momentum, self.eps, self.reducer)
        if x.is_cuda and x.dtype != torch.float64:
            return torch.nn.functional.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                                                  False, self.momentum, self.eps)
        a, b = self.scale_shift()
        shape = [1, self.num_features] + [1] * (x.dim() - 2)
        acc = _acc_dtype(x)
        return (x.to(acc) * a.to(acc).view(shape) + b.to(acc).view(shape)).to(x.dtype)

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}"


def flush_batch_counters(*modules: nn.Module) -> None:
    """``num_batches_tracked += 1`` for every layer whose forward ran since the last flush (deferred mode): one
    multi-tensor launch instead of one tiny kernel per BatchNorm layer (57 in the ResNet-50 encoder)."""
    pend = []
    for mod in modules:
        for m in mod.modules():
            if isinstance(m, BatchNorm) and getattr(m, "_nbt_pending", False):
                m._nbt_pending = False
                pend.append(m.num_batches_tracked)
    if pend:
        with torch.no_grad():
            torch._foreach_add_(pend, 1)


def set_stat_reducer(module: nn.Module, reducer: Reducer) -> None:
    for m in module.modules():
        if isinstance(m, BatchNorm):
            m.reducer = reducer
