"""ResNet encoder on the tcgen05 conv engine (Python orchestration).

Same kernels as the decoder (``csrc/conv_tcgen05.cu``): every 1x1 and 3x3 convolution of the trunk - stride 1 or 2,
forward, data gradient and weight gradient - is a tap table for ``conv_taps`` / ``wgrad_taps``:

    fprop 3x3 /s : 9 taps (ky-1, kx-1), TMA element stride s, zero padding = TMA out-of-bounds fill
    fprop 1x1 /s : 1 tap, element stride s
    dgrad  /1    : mirrored taps, transposed weight pack
    dgrad 3x3 /2 : 4 output-parity phases x (1|2)x(1|2) taps (zero-weight taps pad every phase to 4), strided store
    dgrad 1x1 /2 : 1 tap, strided store into a zeroed tensor
    wgrad        : taps as fprop, x gathered with element stride s

BatchNorm statistics come out of the conv epilogue (per-channel sum / sum of squares); normalise + residual add +
ReLU is one elementwise kernel, its backward two (reduce, apply) with the cross-GPU reduction of the 2C sums in
between - the same structure as the decoder layers.  The 7x7 stem runs as a single-tap GEMM on an unfolded image,
and the decoder's per-image convolutions (receptive-field extension, shared skip maps) use the same entry points,
so in this mode no library convolution is left on the prediction path (max-pool / nearest upsampling / unfold are
framework data-movement ops).

Reference semantics: ``network/monodepth2/resnet_encoder.py:88-108`` (torchvision ResNet trunk, five outputs).
Status: validated against the ``nn.Module`` encoder through the kernel specification (``ops/emu.py``,
``tests/test_engine_emulated.py``) and on B200 (``tests/test_encoder_engine_gpu.py``); selected with
``MINE_B200_ENCODER=tcgen05``.  ``MINE_B200_ENCODER=hybrid`` keeps the library's implicit GEMMs for the convolutions
(they are small and latency bound) and uses only the fused BatchNorm / residual / ReLU kernels of this path.
Default: ``cudnn`` (library convolutions and ATen BatchNorm under autocast).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import conv_engine as E
from .conv_engine import BN_EPS, _count, ctx_world, ext, pick_tile


def _taps(k: int) -> Tuple[List[int], List[int]]:
    p = k // 2
    return [ky - p for ky in range(k) for kx in range(k)], [kx - p for ky in range(k) for kx in range(k)]


def _out_size(n: int, k: int, stride: int) -> int:
    return (n + 2 * (k // 2) - k) // stride + 1


def _pack_fprop(w: torch.Tensor) -> torch.Tensor:
    """[Co,Ci,k,k] -> [k*k, Co, Ci] in the operand dtype."""
    co, ci, k, _ = w.shape
    return E.to_operand(w.detach().permute(2, 3, 0, 1).reshape(k * k, co, ci)).contiguous()


def _pack_dgrad(w: torch.Tensor) -> torch.Tensor:
    """[Co,Ci,k,k] -> [k*k, Ci, Co]."""
    co, ci, k, _ = w.shape
    return E.to_operand(w.detach().permute(2, 3, 1, 0).reshape(k * k, ci, co)).contiguous()


# per-axis decomposition of the stride-2 3x3 data gradient: dx[2q + p] = sum_a dy[q + OFF[p][a]] * W[K[p][a]]
_S2_OFF = ((0, 0), (1, 0))
_S2_K = ((1, -1), (0, 2))            # -1: zero-weight filler tap (keeps T uniform across the four phases)
_S2_INDEX_CACHE = {}


def _pack_dgrad_s2(w: torch.Tensor) -> torch.Tensor:
    """[Co,Ci,3,3] -> [16 (phase*4 + tap), Ci, Co] for the four output-parity phases."""
    co, ci = w.shape[:2]
    key = str(w.device)
    if key not in _S2_INDEX_CACHE:
        idx = []
        for py in range(2):
            for px in range(2):
                for a in range(2):
                    for b in range(2):
                        ky, kx = _S2_K[py][a], _S2_K[px][b]
                        idx.append(9 if (ky < 0 or kx < 0) else ky * 3 + kx)
        _S2_INDEX_CACHE[key] = torch.tensor(idx, dtype=torch.long).to(w.device)
    taps = torch.cat([_pack_dgrad(w), torch.zeros((1, ci, co), dtype=E.ACT_DTYPE, device=w.device)], dim=0)
    return taps.index_select(0, _S2_INDEX_CACHE[key]).contiguous()


_S2_TY = [_S2_OFF[py][a] for py in range(2) for px in range(2) for a in range(2) for b in range(2)]
_S2_TX = [_S2_OFF[px][b] for py in range(2) for px in range(2) for a in range(2) for b in range(2)]
_S2_OY = [py for py in range(2) for px in range(2)]
_S2_OX = [px for py in range(2) for px in range(2)]


# ---------------------------------------------------------------------------------------------
# raw wrappers
# ---------------------------------------------------------------------------------------------
NUM_SMS = 148


def split_factor(work_items: int, iters: int) -> int:
    """K-split for layers whose output tiles cannot fill the machine (``MINE_B200_SPLITK=1``, opt-in): enough CTAs
    for ~2 per SM, at least 4 (tap, channel-block) iterations per CTA, at most 16 partial sums per output."""
    if os.environ.get("MINE_B200_SPLITK", "0") != "1" or work_items >= NUM_SMS:
        return 1
    ks = min((2 * NUM_SMS + work_items - 1) // work_items, iters // 4, 16)
    return max(ks, 1)


def _taps_conv(x, wpack, hg: int, wg: int, ty, tx, stride: int, co: int, stats) -> torch.Tensor:
    """One tap group through ``conv_taps`` - or, when the output has too few tiles to fill the GPU and
    ``MINE_B200_SPLITK=1``, through the split-K kernel (fp32 partial sums) + finalize (bf16 + BatchNorm sums)."""
    n, ci, t = x.shape[0], x.shape[3], len(ty)
    th, tw = pick_tile(hg, wg)
    work = ((hg + th - 1) // th) * ((wg + tw - 1) // tw) * n * (1 if co <= 256 else co // 128)
    ks = split_factor(work, t * max(1, ci // 64))
    if ks > 1 and co >= 16 and (co & (co - 1)) == 0:
        acc = torch.zeros((n, hg, wg, co), dtype=torch.float32, device=x.device)
        ext().conv_taps_splitk(x, wpack, acc, hg, wg, t, ty, tx, stride, co, th, tw, ks)
        out, st = ext().splitk_finalize(acc, stats is not None)
        if stats is not None:
            stats += st
        _count(3)
        return out
    out = torch.empty((n, hg, wg, co), dtype=E.ACT_DTYPE, device=x.device)
    ext().conv_taps(x, wpack, out, hg, wg, 1, t, ty, tx, stride, co, 1, 1, [0], [0], False, None, None, None, 1, stats, 0,
                    False, None, th, tw)
    _count()
    return out


def conv_fprop(x: torch.Tensor, w: torch.Tensor, stride: int, stats: Optional[torch.Tensor]) -> torch.Tensor:
    """``x [N,H,W,Ci]`` -> ``conv(x, w, stride, padding=k//2) [N,Ho,Wo,Co]``; ``stats [2,Co]`` accumulates BN sums."""
    _, h, w_, _ = x.shape
    co, _, k, _ = w.shape
    ty, tx = _taps(k)
    _count()
    return _taps_conv(x, _pack_fprop(w), _out_size(h, k, stride), _out_size(w_, k, stride), ty, tx, stride, co, stats)


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, stride: int, h: int, w_: int) -> torch.Tensor:
    """Gradient w.r.t. the ``[N,h,w_,Ci]`` input of :func:`conv_fprop`."""
    n, ho, wo, co = dy.shape
    ci, k = w.shape[1], w.shape[2]
    if stride == 1:
        ty, tx = _taps(k)
        _count()
        return _taps_conv(dy, _pack_dgrad(w), h, w_, [-t for t in ty], [-t for t in tx], 1, ci, None)
    if stride != 2:
        raise ValueError("stride must be 1 or 2")
    if k == 1:
        out = torch.zeros((n, h, w_, ci), dtype=E.ACT_DTYPE, device=dy.device)        # odd rows / columns stay zero
        th, tw = pick_tile(ho, wo)
        ext().conv_taps(dy, _pack_dgrad(w), out, ho, wo, 1, 1, [0], [0], 1, ci, 2, 2, [0], [0], False,
                        None, None, None, 1, None, 0, False, None, th, tw)
        _count(3)
        return out
    if k != 3 or h % 2 or w_ % 2:
        raise ValueError("stride-2 data gradient needs a 3x3 kernel and even input size")
    out = torch.empty((n, h, w_, ci), dtype=E.ACT_DTYPE, device=dy.device)
    th, tw = pick_tile(h // 2, w_ // 2)
    ext().conv_taps(dy, _pack_dgrad_s2(w), out, h // 2, w_ // 2, 4, 4, _S2_TY, _S2_TX, 1, ci, 2, 2, _S2_OY, _S2_OX,
                    False, None, None, None, 1, None, 0, False, None, th, tw)
    _count(4)
    return out


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, k: int, stride: int) -> torch.Tensor:
    """fp32 ``[Co,Ci,k,k]`` weight gradient of :func:`conv_fprop`."""
    n, ho, wo, co = dy.shape
    ci = x.shape[3]
    if ci > 128 and ci % 128:
        raise ValueError("wgrad_taps needs Ci <= 128 or a multiple of 128")
    ty, tx = _taps(k)
    dw = torch.zeros((k * k, co, ci), dtype=torch.float32, device=dy.device)
    pix = E._wgrad_pixels(co, ci, ho, wo)
    if stride > 1:
        pix = min(pix, 128)                        # strided x box: stride * TW <= 256
    th, tw = pick_tile(ho, wo, pix)
    ext().wgrad_taps(dy, x, dw, ho, wo, 1, k * k, ty, tx, 1, [0], [0], th, tw, stride)
    _count(2)
    return dw.reshape(k, k, co, ci).permute(2, 3, 0, 1)


# ---------------------------------------------------------------------------------------------
# autograd nodes
# ---------------------------------------------------------------------------------------------
class Conv(torch.autograd.Function):
    """``y, stats = conv(x, w)``: NHWC activations in the operand dtype, fp32 master weights ``[Co,Ci,k,k]``."""

    @staticmethod
    def forward(ctx, x, w, stride, want_stats):
        stats = torch.zeros((2, w.shape[0]), dtype=torch.float32, device=x.device) if want_stats else None
        y = conv_fprop(x, w, stride, stats)
        ctx.save_for_backward(x, w)
        ctx.stride = int(stride)
        if stats is None:
            stats = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dw = conv_wgrad(dy, x, w.shape[2], ctx.stride).to(w.dtype) if ctx.needs_input_grad[1] else None
        dx = conv_dgrad(dy, w, ctx.stride, x.shape[1], x.shape[2]) if ctx.needs_input_grad[0] else None
        return dx, dw, None, None


RELU, LEAKY, IDENTITY = 0.0, 0.1, 1.0        # activation slopes: act(v) = v if v > 0 else slope * v


class BNAct(torch.autograd.Function):
    """``a = act(BN(y) [+ residual])`` from the conv epilogue's batch sums; training or running statistics.
    ``relu``: ``True`` / ``False`` (ReLU / none) or the negative slope of a LeakyReLU."""

    @staticmethod
    def forward(ctx, y, stats, gamma, beta, residual, relu, bn, reducer):
        slope = (RELU if relu else IDENTITY) if isinstance(relu, bool) else float(relu)
        training = bn is None or bn.training
        count = float(y.shape[0] * y.shape[1] * y.shape[2])
        fx = None
        if training:
            if stats.numel() == 0:
                stats = ext().channel_stats(y)
                _count()
            if reducer is not None:
                count *= ctx_world(reducer)
                fx = E.fused_exchange(reducer, stats.numel())
                if fx is None:
                    stats = reducer(stats.reshape(-1)).reshape(2, -1).contiguous()
        else:
            rm, rv = bn.running_mean.float(), bn.running_var.float()
            stats = torch.stack([rm * count, (rv + rm * rm) * count]).contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        if fx is not None:                   # statistic exchange inside the normalise kernel (csrc/ll_exchange.cuh)
            a, stats = ext().bn_res_act_fwd_x(y, stats, g32, b32, residual, slope, count, BN_EPS, *fx)
        else:
            a = ext().bn_res_act_fwd(y, stats, g32, b32, residual, slope, count, BN_EPS)
        _count()
        if bn is not None and training:
            E.update_running_stats(bn, stats, count)
        ctx.save_for_backward(y, a, stats, g32, b32)
        ctx.cfg = (slope, count, reducer, residual is not None, training)
        return a

    @staticmethod
    def backward(ctx, da):
        y, a, stats, g32, b32 = ctx.saved_tensors
        slope, count, reducer, has_res, training = ctx.cfg
        g, sums = ext().bn_res_act_bwd_reduce(da.contiguous(), a, y, stats, g32, b32, slope, count, BN_EPS)
        dgamma, dbeta = sums[1], sums[0]
        fx = E.fused_exchange(reducer, sums.numel()) if training else None
        if not training:             # frozen statistics: BN is a per-channel affine map
            sums = torch.zeros_like(sums)
        elif reducer is not None and fx is None:    # separate all-reduce, in place: keep the local sums (= gradients)
            dgamma, dbeta = sums[1].clone(), sums[0].clone()
            sums = reducer(sums.reshape(-1)).reshape(2, -1).contiguous()
        if fx is not None:
            dy = ext().bn_bwd_apply_x(g, y, stats, g32, sums, 1, False, False, count, BN_EPS, *fx)[0]
        else:
            dy = ext().bn_bwd_apply(g, y, stats, g32, sums, 1, False, False, count, BN_EPS)[0]
        _count(2)
        return dy, None, dgamma.to(g32.dtype), dbeta.to(b32.dtype), (g if has_res else None), None, None, None


class ConvLib(torch.autograd.Function):
    """Library convolution on the engine's NHWC operand layout (``hybrid`` mode: cuDNN implicit GEMMs for the small
    encoder layers, this package's fused BatchNorm / residual / ReLU kernels around them).  The NHWC tensor is
    handed over as a channels-last NCHW view, so no copies are made in either direction."""

    @staticmethod
    def forward(ctx, x, w, stride):
        k = w.shape[2]
        w_op = w.detach().to(E.ACT_DTYPE)                               # keeps the (channels-last) weight strides
        y = torch.ops.aten.convolution(x.permute(0, 3, 1, 2), w_op, None, [stride, stride], [k // 2, k // 2], [1, 1],
                                       False, [0, 0], 1)
        _count(2)
        ctx.save_for_backward(x, w_op)
        ctx.stride = int(stride)
        return y.permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def backward(ctx, dy):
        x, w_op = ctx.saved_tensors
        k, s = w_op.shape[2], ctx.stride
        dx, dw, _ = torch.ops.aten.convolution_backward(
            dy.permute(0, 3, 1, 2), x.permute(0, 3, 1, 2), w_op, None, [s, s], [k // 2, k // 2], [1, 1], False, [0, 0], 1,
            [bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1]), False])
        _count(2)
        if dx is not None:
            dx = dx.permute(0, 2, 3, 1).contiguous()
        return dx, (dw.float() if dw is not None else None), None


def conv_bn_act(x, conv, bn, relu=True, residual=None, reducer=None, library_conv=False):
    stride = conv.stride[0]
    if library_conv:
        y = ConvLib.apply(x, conv.weight, stride)
        stats = torch.empty(0, device=x.device)                         # BNAct computes them (channel_stats kernel)
    else:
        y, stats = Conv.apply(x, conv.weight, stride, bn.training)
    return BNAct.apply(y, stats, bn.weight, bn.bias, residual, relu, bn, reducer)


# ---------------------------------------------------------------------------------------------
# pieces that make the whole prediction path library-convolution free (MINE_B200_ENCODER=tcgen05)
# ---------------------------------------------------------------------------------------------
STEM_K = 256        # 7 * 7 * 3 = 147 im2col columns, zero padded: fprop needs a multiple of the 64-channel K block,
                    # wgrad_taps a multiple of its 128-column accumulator block


class StemConv(torch.autograd.Function):
    """7x7 / stride 2 / pad 3 convolution of the 3-channel image as ONE GEMM tap: the image is unfolded into
    ``[N, H/2, W/2, 147 -> 256]`` rows (a 25 MB gather at 384x256, no arithmetic) and multiplied with the
    ``[Co, 256]`` weight matrix by ``conv_taps`` (BatchNorm sums in the epilogue).  The image needs no gradient."""

    @staticmethod
    def forward(ctx, x_nchw, w, want_stats):
        n, _, h, w_ = x_nchw.shape
        co = w.shape[0]
        ho, wo = _out_size(h, 7, 2), _out_size(w_, 7, 2)
        cols = F.unfold(x_nchw.float(), 7, padding=3, stride=2)                       # [N, 147, Ho*Wo], (c, ky, kx) order
        a = torch.zeros((n, ho, wo, STEM_K), dtype=E.ACT_DTYPE, device=x_nchw.device)
        a[..., :147] = E.to_operand(cols.transpose(1, 2).reshape(n, ho, wo, 147))
        wp = torch.zeros((1, co, STEM_K), dtype=E.ACT_DTYPE, device=w.device)
        wp[0, :, :147] = E.to_operand(w.detach().reshape(co, 147))
        stats = torch.zeros((2, co), dtype=torch.float32, device=w.device) if want_stats else None
        y = torch.empty((n, ho, wo, co), dtype=E.ACT_DTYPE, device=w.device)
        th, tw = pick_tile(ho, wo)
        ext().conv_taps(a, wp, y, ho, wo, 1, 1, [0], [0], 1, co, 1, 1, [0], [0], False, None, None, None, 1, stats, 0,
                        False, None, th, tw)
        _count(6)
        ctx.save_for_backward(a)
        ctx.wshape = tuple(w.shape)
        if stats is None:
            stats = torch.empty(0, device=w.device)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        (a,) = ctx.saved_tensors
        dy = dy.contiguous()
        n, ho, wo, co = dy.shape
        dw = torch.zeros((1, co, STEM_K), dtype=torch.float32, device=dy.device)
        th, tw = pick_tile(ho, wo, E._wgrad_pixels(co, STEM_K, ho, wo))
        ext().wgrad_taps(dy, a, dw, ho, wo, 1, 1, [0], [0], 1, [0], [0], th, tw, 1)
        _count(2)
        return None, dw[0, :, :147].reshape(ctx.wshape), None


class SharedConv(torch.autograd.Function):
    """Per-image skip-feature convolution of the factorised decoder: ``conv3x3(xpad) -> fp32 [B,H,W,Co]`` (the map the
    per-plane conv epilogue adds to every plane); ``xpad`` is the reflection-padded NHWC feature map."""

    @staticmethod
    def forward(ctx, xpad, w):
        n, hp, wp_, _ = xpad.shape
        out = torch.empty((n, hp - 2, wp_ - 2, w.shape[0]), dtype=torch.float32, device=xpad.device)
        E.conv_same_raw(xpad, w, out=out)
        ctx.save_for_backward(xpad, w)
        return out

    @staticmethod
    def backward(ctx, dmap):
        xpad, w = ctx.saved_tensors
        dy = E.to_operand(dmap).contiguous()
        dw = E.wgrad_same_raw(dy, xpad).to(w.dtype) if ctx.needs_input_grad[1] else None
        dx = E.dgrad_same_raw(dy, w) if ctx.needs_input_grad[0] else None
        return dx, dw


def _pool_nhwc(x: torch.Tensor) -> torch.Tensor:
    return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()


def _upsample_nhwc(x: torch.Tensor, size) -> torch.Tensor:
    if tuple(x.shape[1:3]) == tuple(size):
        return x
    return F.interpolate(x.permute(0, 3, 1, 2), size=tuple(size), mode="nearest").permute(0, 2, 3, 1).contiguous()


def receptive_field_extension(dec, top_nchw: torch.Tensor, reducer=None, library_conv: bool = False) -> torch.Tensor:
    """``DepthDecoder.receptive_field_extension`` (reference ``depth_decoder.py:55-61,96-101``) on the engine:
    pool -> 1x1 -> pool -> 3x3 -> up -> 3x3 -> up -> 1x1, every conv followed by BN + LeakyReLU(0.1)."""
    with E.output_rounding(not library_conv):
        return _receptive_field_extension(dec, top_nchw, reducer, library_conv)


def _receptive_field_extension(dec, top_nchw, reducer, library_conv):
    top = E.to_operand(top_nchw.permute(0, 2, 3, 1)).contiguous()

    def layer(x, blk):
        if library_conv:                           # "hybrid": library convolution, fused BN / LeakyReLU kernels
            y, stats = ConvLib.apply(x, blk[0].weight, 1), torch.empty(0, device=x.device)
        else:
            y, stats = Conv.apply(x, blk[0].weight, 1, blk[1].training)
        return BNAct.apply(y, stats, blk[1].weight, blk[1].bias, None, LEAKY, blk[1], reducer)
    d1 = layer(_pool_nhwc(top), dec.conv_down1)
    d2 = layer(_pool_nhwc(d1), dec.conv_down2)
    u1 = layer(_upsample_nhwc(d2, d1.shape[1:3]), dec.conv_up1)
    return layer(_upsample_nhwc(u1, top.shape[1:3]), dec.conv_up2).permute(0, 3, 1, 2)


def shared_skip_map(feat_nchw: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``conv3x3(reflect_pad(feat))`` as the fp32 NHWC map the plane-conv epilogue consumes."""
    x = E.to_operand(F.pad(feat_nchw.float(), (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)).contiguous()
    return SharedConv.apply(x, w)


# ---------------------------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------------------------
class EncoderEngine:
    """Runs a :class:`mine_b200.models.encoder.ResnetEncoder` (any depth) on the engine; returns the five NCHW
    (channels-last strided) feature maps ``ResnetEncoder.forward`` returns."""

    def __init__(self, backbone, library_conv: bool = False):
        self.backbone = backbone
        self.library_conv = bool(library_conv)       # "hybrid": library convolutions + our fused BN kernels

    def _reducer(self):
        return self.backbone.encoder.bn1.reducer

    def _block(self, blk, x, reducer):
        kw = dict(reducer=reducer, library_conv=self.library_conv)
        if blk.downsample is not None:
            idt = conv_bn_act(x, blk.downsample[0], blk.downsample[1], relu=False, **kw)
        else:
            idt = x
        if hasattr(blk, "conv3"):                                   # bottleneck
            out = conv_bn_act(x, blk.conv1, blk.bn1, **kw)
            out = conv_bn_act(out, blk.conv2, blk.bn2, **kw)
            return conv_bn_act(out, blk.conv3, blk.bn3, relu=True, residual=idt, **kw)
        out = conv_bn_act(x, blk.conv1, blk.bn1, **kw)
        return conv_bn_act(out, blk.conv2, blk.bn2, relu=True, residual=idt, **kw)

    def __call__(self, img: torch.Tensor):
        # hybrid: every activation produced here feeds a library convolution -> plain fp32, no TF32 rounding on store
        with E.output_rounding(not self.library_conv):
            return self._forward(img)

    def _forward(self, img: torch.Tensor):
        bb, e = self.backbone, self.backbone.encoder
        reducer = self._reducer()
        x = (img - bb.img_mean.to(img.dtype)) / bb.img_std.to(img.dtype)
        if self.library_conv or e.conv1.weight.shape[1] != 3:
            amp = dict(device_type=img.device.type, dtype=torch.bfloat16,
                       enabled=img.is_cuda and E.ACT_DTYPE == torch.bfloat16)
            with torch.autocast(**amp):
                y = F.conv2d(x.contiguous(memory_format=torch.channels_last), e.conv1.weight, None, 2, 3)
            y, stats = E.to_operand(y.permute(0, 2, 3, 1)).contiguous(), torch.empty(0, device=img.device)
        else:                                                      # stem as one im2col GEMM tap on the engine
            y, stats = StemConv.apply(x, e.conv1.weight, e.bn1.training)
        c1 = BNAct.apply(y, stats, e.bn1.weight, e.bn1.bias, None, True, e.bn1, reducer)
        x = F.max_pool2d(c1.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()
        feats = [c1]
        for li in range(1, 5):
            for blk in getattr(e, f"layer{li}"):
                x = self._block(blk, x, reducer)
            feats.append(x)
        return tuple(f.permute(0, 3, 1, 2) for f in feats)
