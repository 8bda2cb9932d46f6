"""Executable specification of the conv-engine entry points (``csrc/conv_bindings.cpp``) in plain PyTorch.

Every function here has the signature and the semantics of the sm_100a kernel of the same name - tap tables,
sub-pixel groups, TMA out-of-bounds zero fill, strided gathers, epilogue fusions, BatchNorm partial sums - but runs
anywhere (CPU included).  Two uses:

* ``tests/test_engine_emulated.py`` drives the Python orchestration of the engine (``conv_engine.py``,
  ``encoder_engine.py``: packing, tap tables, autograd wiring, BN algebra) against the ``nn.Module`` models on the
  CPU test tier, where the kernels themselves cannot run;
* it is the written-down contract the GPU tests hold the kernels to (``tests/test_conv_engine_gpu.py`` compares
  kernel and emulator outputs for the same arguments).

It is NOT a fallback: ``conv_engine.ext()`` only returns it when ``conv_engine.use_emulator(True)`` was called
explicitly (tests) - on a GPU the real extension is used or an error is raised.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

ACT_DTYPE = torch.bfloat16          # ``conv_engine.use_emulator(True, torch.float32)`` switches to fp32 for tight tests
# ``True``: fp32 operands behave like the tf32 kernels - the MMA reads only the 10 high mantissa bits (truncation) and
# every operand PRODUCER rounds to the nearest TF32 value (``cvt.rna``).  ``False`` (default): exact fp32, the
# device-independent contract the CPU tier tests the orchestration against.
TF32_OPERANDS = False


def _trunc(x: torch.Tensor) -> torch.Tensor:
    """What a kind::tf32 MMA sees of an fp32 operand."""
    if TF32_OPERANDS and x.dtype == torch.float32:
        return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    return x


def _round_op(x: torch.Tensor) -> torch.Tensor:
    """Operand store of a producer kernel (``st8_op`` / ``store_operand``)."""
    if TF32_OPERANDS and x.dtype == torch.float32:
        return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return x


def _gather(x: torch.Tensor, iy: torch.Tensor, ix: torch.Tensor) -> torch.Tensor:
    """``x[:, iy, ix, :]`` with zero fill outside the tensor (what a TMA box returns)."""
    hi, wi = x.shape[1], x.shape[2]
    vy, vx = (iy >= 0) & (iy < hi), (ix >= 0) & (ix < wi)
    patch = x[:, iy.clamp(0, hi - 1)][:, :, ix.clamp(0, wi - 1)]
    return patch * (vy[:, None] & vx[None, :])[None, :, :, None].to(patch.dtype)


def conv_taps(x, wpack, out, Hg, Wg, G, T, tap_y, tap_x, in_stride, Co, out_sy, out_sx, out_oy, out_ox, accumulate,
              chan_bias, plane_bias, shared_map, planes_per_image, stats, act, head_alpha, raw_out, TH, TW) -> None:
    """``out[n, oy*out_sy + out_oy[g], ox*out_sx + out_ox[g], :Co] = sum_t x[n, oy*in_stride + tap_y[g,t],
    ox*in_stride + tap_x[g,t], :] @ wpack[g*T + t, :Co, :]^T`` (+ epilogue terms), for oy < Hg, ox < Wg."""
    n = x.shape[0]
    dev = x.device
    xf, wf = _trunc(x.float()), _trunc(wpack.float())
    oy, ox = torch.arange(Hg, device=dev), torch.arange(Wg, device=dev)
    planes = max(int(planes_per_image), 1)
    for g in range(G):
        acc = torch.zeros((n, Hg, Wg, Co), dtype=torch.float32, device=dev)
        for t in range(T):
            patch = _gather(xf, oy * in_stride + tap_y[g * T + t], ox * in_stride + tap_x[g * T + t])
            acc += patch @ wf[g * T + t, :Co].t()
        yy, xx = oy * out_sy + out_oy[g], ox * out_sx + out_ox[g]
        if chan_bias is not None:
            acc += chan_bias.float()
        if plane_bias is not None:
            acc += plane_bias.float().reshape(n, 1, 1, Co)
        if shared_map is not None:
            sm = shared_map.float().reshape(n // planes, shared_map.shape[-3], shared_map.shape[-2], Co)
            acc += sm[:, yy][:, :, xx].repeat_interleave(planes, dim=0)
        if stats is not None:
            stats[0] += acc.sum(dim=(0, 1, 2))
            stats[1] += (acc * acc).sum(dim=(0, 1, 2))
        if act == 1:
            o4 = out.reshape(n, out.shape[-3], out.shape[-2], 4)
            rgb = torch.sigmoid(acc[..., :3])
            last = torch.sigmoid(acc[..., 3:4]) if head_alpha else acc[..., 3:4].abs() + 1e-4
            o4[:, yy[:, None], xx[None, :]] = torch.cat([rgb, last], dim=-1)
            if raw_out is not None:
                sgn = torch.where(acc[..., 3] >= 0, 1, -1).to(raw_out.dtype)
                raw_out.reshape(n, out.shape[-3], out.shape[-2])[:, yy[:, None], xx[None, :]] = sgn
        else:
            if accumulate:
                acc = acc + out[:, yy[:, None], xx[None, :]].float()
            out[:, yy[:, None], xx[None, :]] = acc.to(out.dtype)


def wgrad_taps(dy, x, dw, Hg, Wg, G, T, tap_y, tap_x, dy_stride, dy_oy, dy_ox, TH, TW, x_stride: int = 1) -> None:
    """``dw[g*T + t, co, ci] += sum_{n, oy < Hg, ox < Wg} dy[n, oy*dy_stride + dy_oy[g], ox*dy_stride + dy_ox[g], co] *
    x[n, oy*x_stride + tap_y[g,t], ox*x_stride + tap_x[g,t], ci]`` (zero outside either tensor)."""
    dev = x.device
    dyf, xf = _trunc(dy.float()), _trunc(x.float())
    oy, ox = torch.arange(Hg, device=dev), torch.arange(Wg, device=dev)
    for g in range(G):
        dyg = _gather(dyf, oy * dy_stride + dy_oy[g], ox * dy_stride + dy_ox[g])
        for t in range(T):
            xg = _gather(xf, oy * x_stride + tap_y[g * T + t], ox * x_stride + tap_x[g * T + t])
            dw[g * T + t] += torch.einsum("nhwo,nhwi->oi", dyg, xg)


def _phase_has(p: int, a: int, k: int) -> bool:
    return (k == 0 if a == 0 else k >= 1) if p == 0 else (k <= 1 if a == 0 else k == 2)


def pack_weights(w: torch.Tensor, mode: int) -> torch.Tensor:
    """fp32 ``[Co,Ci,3,3]`` -> ``[9 | 16, rows_pad, cols]`` GEMM operand pack (modes: see ``conv_tcgen05.cu``)."""
    co, ci = w.shape[:2]
    dgrad, up = mode >= 2, bool(mode & 1)
    if not up:
        taps = [w[:, :, ky, kx] for ky in range(3) for kx in range(3)]
    else:
        taps = []
        for py in range(2):
            for px in range(2):
                for a in range(2):
                    for b in range(2):
                        acc = torch.zeros_like(w[:, :, 0, 0])
                        for ky in range(3):
                            for kx in range(3):
                                if _phase_has(py, a, ky) and _phase_has(px, b, kx):
                                    acc = acc + w[:, :, ky, kx]
                        taps.append(acc)
    mats = [m.t() if dgrad else m for m in taps]                         # rows: Co (fprop) or Ci (dgrad)
    rows = mats[0].shape[0]
    rows_pad = (rows + 15) // 16 * 16
    out = torch.zeros((len(mats), rows_pad, mats[0].shape[1]), dtype=ACT_DTYPE, device=w.device)
    for i, m in enumerate(mats):
        out[i, :rows] = _round_op(m.to(ACT_DTYPE))
    return out


def _pad_src(n: int, mode: int, device) -> torch.Tensor:
    s = torch.arange(-1, n + 1, device=device)
    if mode == 0:
        s = s.abs()
        s = torch.where(s >= n, 2 * n - 2 - s, s)
    else:
        s = s.clamp(0, n - 1)
    return s


def _bn_coef(stats, gamma, beta, count, eps):
    mean = stats[0] / count
    var = (stats[1] / count - mean * mean).clamp_min(0)
    invstd = torch.rsqrt(var + eps)
    a = gamma * invstd
    return mean, invstd, a, beta - mean * a


def bn_act_pad_fwd(y, stats, gamma, beta, pad_mode, count, eps):
    """``pad(ELU(BN(y)))``: reflection (0) or replication (1) border of one pixel, NHWC."""
    _, _, a, b = _bn_coef(stats.float(), gamma.float(), beta.float(), count, eps)
    u = F.elu(y.float() * a + b)
    sy, sx = _pad_src(y.shape[1], pad_mode, y.device), _pad_src(y.shape[2], pad_mode, y.device)
    return _round_op(u[:, sy][:, :, sx].to(y.dtype).contiguous())


def _bn_act_g(dapad, y, stats, gamma, beta, pad_mode, count, eps):
    """fp32 gradient w.r.t. the BatchNorm output: adjoint of the pad times ELU' (shared by the two forms below)."""
    n, h, w, c = y.shape
    mean, invstd, a, b = _bn_coef(stats.float(), gamma.float(), beta.float(), count, eps)
    sy, sx = _pad_src(h, pad_mode, y.device), _pad_src(w, pad_mode, y.device)
    d = torch.zeros((n, h, w, c), dtype=torch.float32, device=y.device)
    d.index_put_((torch.arange(n, device=y.device)[:, None, None], sy[None, :, None], sx[None, None, :]),
                 dapad.float(), accumulate=True)
    yf = y.float()
    u = yf * a + b
    g = d * torch.where(u > 0, torch.ones_like(u), torch.exp(u))
    xhat = (yf - mean) * invstd
    return g, torch.stack([g.sum(dim=(0, 1, 2)), (g * xhat).sum(dim=(0, 1, 2))])


def bn_act_bwd_reduce(dapad, y, stats, gamma, beta, pad_mode, count, eps):
    """Adjoint of the pad, times ELU', plus the two BatchNorm backward sums ``[sum g, sum g*xhat]``."""
    g, sums = _bn_act_g(dapad, y, stats, gamma, beta, pad_mode, count, eps)
    return [g.to(y.dtype), sums]


def bn_act_bwd_sums(dapad, y, stats, gamma, beta, pad_mode, count, eps):
    """The two sums of :func:`bn_act_bwd_reduce` without materialising ``g`` (default backward path)."""
    return _bn_act_g(dapad, y, stats, gamma, beta, pad_mode, count, eps)[1]


def bn_bwd_apply_fused(dapad, y, stats, gamma, beta, sums, planes_per_image, want_shared, want_plane_bias, count, eps,
                       pad_mode):
    """:func:`bn_bwd_apply` with ``g`` recomputed from the padded upstream gradient (kept in fp32: never stored)."""
    g = _bn_act_g(dapad, y, stats, gamma, beta, pad_mode, count, eps)[0]
    return bn_bwd_apply(g, y, stats, gamma, sums, planes_per_image, want_shared, want_plane_bias, count, eps)


def bn_bwd_apply(g, y, stats, gamma, sums, planes_per_image, want_shared, want_plane_bias, count, eps):
    n, h, w, c = y.shape
    s = max(int(planes_per_image), 1)
    mean, invstd, _, _ = _bn_coef(stats.float(), gamma.float(), torch.zeros_like(gamma.float()), count, eps)
    xhat = (y.float() - mean) * invstd
    dy = gamma.float() * invstd * (g.float() - sums[0] / count - xhat * (sums[1] / count))
    dshared = dy.reshape(n // s, s, h, w, c).sum(dim=1) if want_shared else None
    dpb = dy.sum(dim=(1, 2)) if want_plane_bias else None
    return [_round_op(dy.to(y.dtype)), dshared, dpb]


def head_bwd(g_mpi, mpi, sign, use_alpha):
    """Gradient of ``(sigmoid rgb, |x| + 1e-4 or sigmoid)`` as the 16-channel tensor the GEMMs consume."""
    gm, o = g_mpi.reshape(-1, 4).float(), mpi.reshape(-1, 4).float()
    d = gm * o * (1 - o)
    if not use_alpha:
        d = torch.cat([d[:, :3], gm[:, 3:] * sign.reshape(-1, 1).float()], dim=1)
    dz = torch.zeros((*sign.shape, 16), dtype=ACT_DTYPE, device=mpi.device)
    dz[..., :4] = _round_op(d.reshape(*sign.shape, 4).to(ACT_DTYPE))
    return [dz, d.sum(dim=0)]



# ---- encoder companions (csrc/encoder_elem.cu) ------------------------------------------------------------------
ROUND_ENCODER_OUT = True          # conv_engine.output_rounding: False while the hybrid encoder feeds library convolutions


def bn_res_act_fwd(y, stats, gamma, beta, residual, slope, count, eps):
    """``act(BN(y) [+ residual])`` on an unpadded NHWC tensor; ``act(v) = v if v > 0 else slope * v``
    (0: ReLU, 0.1: LeakyReLU, 1: identity)."""
    _, _, a, b = _bn_coef(stats.float(), gamma.float(), beta.float(), count, eps)
    u = y.float() * a + b
    if residual is not None:
        u = u + residual.float()
    u = torch.where(u > 0, u, u * slope)
    return _round_op(u.to(y.dtype)) if ROUND_ENCODER_OUT else u.to(y.dtype)


def bn_res_act_bwd_reduce(dout, out, y, stats, gamma, beta, slope, count, eps):
    """``g = dout * act'`` (also the gradient of the residual branch) and ``[sum g, sum g*xhat]``."""
    mean, invstd, _, _ = _bn_coef(stats.float(), gamma.float(), beta.float(), count, eps)
    g = dout.float()
    if slope != 1.0:
        g = torch.where(out.float() > 0, g, g * slope)
    xhat = (y.float() - mean) * invstd
    sums = torch.stack([g.sum(dim=(0, 1, 2)), (g * xhat).sum(dim=(0, 1, 2))])
    return [g.to(y.dtype), sums]


def bn_update_running(stats, running_mean, running_var, num_batches_tracked, count, momentum):
    """In-place momentum update of the BatchNorm buffers from the reduced batch sums (unbiased variance)."""
    mean = stats[0].float() / count
    var = (stats[1].float() / count - mean * mean).clamp_min(0)
    running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
    running_var.mul_(1 - momentum).add_((var * (count / max(count - 1.0, 1.0))).to(running_var.dtype), alpha=momentum)
    num_batches_tracked += 1


def bn_update_running_multi(stats, running_mean, running_var, num_batches_tracked, count, momentum):
    for i in range(len(stats)):
        bn_update_running(stats[i], running_mean[i], running_var[i], num_batches_tracked[i], count[i], momentum[i])


def channel_stats(y):
    """``[sum, sum of squares]`` per channel of an NHWC tensor (for convolutions that ran outside the engine)."""
    yf = y.float()
    return torch.stack([yf.sum(dim=(0, 1, 2)), (yf * yf).sum(dim=(0, 1, 2))])


def head_conv_direct(apad, wpk, bias, use_alpha):
    """MPI head on a pre-padded input: ``apad [N,H+2,W+2,C]``, ``wpk [9,C,4]`` (tap-major, output channel last),
    returns the packed fp32 MPI ``[N,H,W,4]`` and the sign of the sigma pre-activation (``csrc/head_direct.cu``)."""
    n, hp, wp_, c = apad.shape
    h, w = hp - 2, wp_ - 2
    xf, wf = apad.float(), wpk.float().reshape(9, c, 4)
    z = bias.float().reshape(1, 1, 1, 4).expand(n, h, w, 4).clone()
    for ky in range(3):
        for kx in range(3):
            z = z + xf[:, ky:ky + h, kx:kx + w] @ wf[ky * 3 + kx]
    last = torch.sigmoid(z[..., 3:]) if use_alpha else z[..., 3:].abs() + 1e-4
    mpi = torch.cat([torch.sigmoid(z[..., :3]), last], dim=-1).contiguous()
    sign = torch.where(z[..., 3] >= 0, 1, -1).to(torch.int8)
    return [mpi, sign]


# ---- sparse-point supervision (csrc/sparse.cu) -------------------------------------------------------------------
def sparse_point_fwd(disp, k, xyz, scale_in):
    b, n = xyz.shape[0], xyz.shape[2]
    h, w = disp.shape[-2], disp.shape[-1]
    p = k.float() @ xyz.float()
    u, v = p[:, 0] / p[:, 2], p[:, 1] / p[:, 2]
    ix = torch.round(u).long().clamp(0, w - 1)
    iy = torch.round(v).long().clamp(0, h - 1)
    idx = iy * w + ix                                                          # [B,N]
    d = torch.gather(disp.float().reshape(b, h * w), 1, idx)
    g = torch.reciprocal(xyz[:, 2].float())
    scale = scale_in.float().clone() if scale_in is not None else torch.exp((torch.log(d) - torch.log(g)).mean(dim=1))
    t = torch.log(d / scale[:, None]) - torch.log(g)
    return [t.abs().sum() / (b * n), scale, idx.to(torch.int32), d, torch.sign(t)]


def sparse_point_bwd(g_loss, g_scale, idx, d_syn, sgn, scale, disp_shape, computed_scale):
    b, n = d_syn.shape
    h, w = disp_shape[-2], disp_shape[-1]
    c = g_loss.reshape(()) / (b * n)
    dl_ds = -c * sgn.sum(dim=1) / scale
    through = torch.zeros(b, dtype=torch.float32, device=d_syn.device)
    grad_scale = torch.zeros(b, dtype=torch.float32, device=d_syn.device)
    if computed_scale:
        gs = dl_ds + (g_scale if g_scale is not None else 0.0)
        through = gs * scale / n
    else:
        grad_scale = dl_ds
    gd = (c * sgn + through[:, None]) / d_syn
    grad = torch.zeros((b, h * w), dtype=torch.float32, device=d_syn.device)
    grad.scatter_add_(1, idx.long(), gd)
    return [grad.reshape(disp_shape), grad_scale]


# ---- split-K convolution for small maps (csrc/conv_splitk.cu) ----------------------------------------------------
def conv_taps_splitk(x, wpack, out32, Hg, Wg, T, tap_y, tap_x, in_stride, Co, TH, TW, ksplit):
    """``out32 += conv_taps(...)`` with one tap group, accumulated in fp32 (the K range is split over CTAs)."""
    tmp = torch.zeros_like(out32)
    conv_taps(x, wpack, tmp, Hg, Wg, 1, T, list(tap_y), list(tap_x), in_stride, Co, 1, 1, [0], [0], False, None, None,
              None, 1, None, 0, False, None, TH, TW)
    out32 += tmp


def splitk_finalize(acc, want_stats):
    y = acc.to(ACT_DTYPE)
    if not want_stats:
        return [y, torch.empty(0, device=acc.device)]
    a = acc.float()
    return [y, torch.stack([a.sum(dim=(0, 1, 2)), (a * a).sum(dim=(0, 1, 2))])]
