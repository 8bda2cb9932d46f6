"""sm_100a convolution engine for the per-plane MPI decoder (Python orchestration).

Kernels: ``csrc/conv_tcgen05.cu`` (tcgen05/TMEM/TMA implicit GEMM: fprop, sub-pixel-phase upsample
fprop, dgrad, wgrad) and ``csrc/decoder_elem.cu`` (BN-apply + ELU + pad, BN backward, head backward).

Layout: activations NHWC bf16 (``[B*S, H, W, C]``); every activation that feeds a 3x3 conv is stored
once, already 1-pixel padded (reflection for same-resolution convs, replication for the fused
nearest-x2-upsample conv - the sub-pixel decomposition of ``conv3x3(reflect_pad(up2(x)))`` is an exact
2x2 conv per output phase on the replicate-padded low-resolution tensor).

Per fused layer (``PlaneConvBNAct``):
    forward : conv_taps (+shared-skip map +per-plane embedding bias, BN partial sums in the epilogue)
              -> [all-reduce of the 2C statistics across GPUs] -> bn_act_pad (normalise + ELU + pad)
    backward: bn_act_bwd_reduce (pad adjoint, ELU', BN reductions) -> [all-reduce] -> bn_bwd_apply
              (dy, shared-skip grad, embedding-bias grad) -> wgrad_taps + conv_taps (dgrad)
Reference semantics: ``network/monodepth2/depth_decoder.py:124-146`` + ``layers.py:106-138``
(ConvBlock = ReflPad + Conv3x3 + BN + ELU, nearest upsample, heads).
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .counters import LAUNCHES

AVAILABLE = True
BN_EPS = 1e-5
# Activation / operand storage type.  Two precisions (``engine.precision``):
#   "tf32": fp32 NHWC activations and weight packs, tcgen05 ``kind::tf32`` MMAs, fp32 accumulation - the numerics class
#           of the reference (fp32 tensors, cuDNN TF32 convolutions; it never autocasts).  DEFAULT.
#   "bf16": bf16 activations / packs, ``kind::f16`` MMAs, fp32 accumulation and fp32 master weights - the fast mode.
ACT_DTYPE = torch.float32
PRECISION = "tf32"

_ext = None
_emulated = False


def set_precision(precision: str) -> None:
    """Select the operand storage type of every engine kernel ("tf32" or "bf16"); process-wide."""
    global ACT_DTYPE, PRECISION
    if precision not in ("tf32", "bf16"):
        raise ValueError("engine.precision must be 'tf32' or 'bf16', got %r" % (precision,))
    PRECISION = precision
    ACT_DTYPE = torch.float32 if precision == "tf32" else torch.bfloat16
    from . import emu
    emu.ACT_DTYPE = ACT_DTYPE
    if not _emulated and _ext is not None:
        _ext.set_operand_size(4 if precision == "tf32" else 2)
    if "_PACK_STATE" in globals():
        _PACK_STATE["plan"] = _PACK_STATE["seen"] = None                 # packs of the other operand type are stale


def set_deterministic(on: bool) -> None:
    """``engine.deterministic``: bitwise reproducible BatchNorm reductions in the encoder kernels of the "hybrid" /
    "tcgen05" encoder modes (two-level fixed-order sums instead of fp32 atomics); process-wide."""
    if not _emulated and ext() is not None and hasattr(ext(), "set_deterministic"):
        ext().set_deterministic(bool(on))


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest TF32-representable fp32 (10 explicit mantissa bits, ties away from zero = ``cvt.rna.tf32.f32``).
    ``kind::tf32`` MMAs ignore the 13 low mantissa bits, i.e. truncate; every tensor-core operand is therefore rounded
    once where it is produced (kernels: ``st8_op`` / ``store_operand``; framework-side producers: this function)."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


_OUTPUT_ROUNDING = True


class output_rounding:
    """``with output_rounding(False):`` - activations produced for LIBRARY convolutions (hybrid encoder) stay plain fp32
    (the library rounds its own operands); the default rounds every produced activation to TF32 because the next
    consumer is a tcgen05 kernel, whose MMA would otherwise truncate."""

    def __init__(self, on: bool):
        self.on = bool(on)

    def __enter__(self):
        global _OUTPUT_ROUNDING
        self.prev, _OUTPUT_ROUNDING = _OUTPUT_ROUNDING, self.on
        from . import emu
        emu.ROUND_ENCODER_OUT = self.on
        if not _emulated and _ext is not None:
            _ext.set_output_rounding(self.on)
        return self

    def __exit__(self, *exc):
        global _OUTPUT_ROUNDING
        _OUTPUT_ROUNDING = self.prev
        from . import emu
        emu.ROUND_ENCODER_OUT = self.prev
        if not _emulated and _ext is not None:
            _ext.set_output_rounding(self.prev)
        return False


def to_operand(t: torch.Tensor) -> torch.Tensor:
    """Cast a framework tensor to the engine's operand storage type (bf16, or fp32 rounded to TF32)."""
    if ACT_DTYPE == torch.float32:
        t = t.float()
        if not _OUTPUT_ROUNDING:
            return t
        if _emulated:
            from . import emu
            if not emu.TF32_OPERANDS:          # exact-fp32 specification mode (CPU tier)
                return t
        r = round_tf32(t.detach())
        return t + (r - t.detach()) if t.requires_grad else r      # straight-through: rounding has unit derivative
    return t.to(ACT_DTYPE)


def use_emulator(flag: bool, dtype: torch.dtype = torch.bfloat16, tf32_operands: bool = False) -> None:
    """Route the engine's kernel calls to the PyTorch specification in ``emu.py`` (any device, tests only).
    ``tf32_operands``: fp32 operands are truncated / rounded exactly like the tf32 kernels do (GPU comparisons)."""
    global _emulated, ACT_DTYPE
    from . import emu
    _emulated = bool(flag)
    emu.TF32_OPERANDS = bool(flag and tf32_operands and dtype == torch.float32)
    if flag:
        ACT_DTYPE = dtype
    else:
        ACT_DTYPE = torch.float32 if PRECISION == "tf32" else torch.bfloat16
    emu.ACT_DTYPE = ACT_DTYPE


def ext():
    global _ext
    if _emulated:
        from . import emu
        return emu
    if _ext is None:
        from . import cuda as C
        _ext = C._ext
        _ext.set_operand_size(4 if ACT_DTYPE == torch.float32 else 2)
    return _ext


def _count(n=1):
    LAUNCHES["count"] += n


# ---------------------------------------------------------------------------------------------
# tiling / packing helpers
# ---------------------------------------------------------------------------------------------
def pick_tile(h: int, w: int, pixels: int = 128) -> Tuple[int, int]:
    """(TH, TW) with TH*TW == pixels maximising the useful fraction of the tiles that cover h x w."""
    best, best_u = None, -1.0
    tw = pixels
    while tw >= 1:
        th = pixels // tw
        if tw <= 256 and th <= 256:
            u = (h * w) / (((h + th - 1) // th) * th * ((w + tw - 1) // tw) * tw)
            if u > best_u + 1e-9 or (abs(u - best_u) < 1e-9 and best is not None and tw > best[1]):
                best, best_u = (th, tw), u
        tw //= 2
    return best


_PHASE = torch.tensor([[[1.0, 0.0, 0.0], [0.0, 1.0, 1.0]],      # phase 0: taps {k0}, {k1 + k2}
                       [[1.0, 1.0, 0.0], [0.0, 0.0, 1.0]]])     # phase 1: taps {k0 + k1}, {k2}
_PHASE_CACHE: Dict = {}


def _phase(device, dtype) -> torch.Tensor:
    """Device-resident copy (a host->device copy is illegal inside CUDA-graph capture)."""
    key = (str(device), dtype)
    if key not in _PHASE_CACHE:
        _PHASE_CACHE[key] = _PHASE.to(device, dtype)
    return _PHASE_CACHE[key]


def pack_up(w: torch.Tensor) -> torch.Tensor:
    """[Co, Ci, 3, 3] -> phase pack [4 (py*2+px), 4 (a*2+b), Co, Ci] of the 2x2 sub-pixel kernels (torch form of
    ``pack(w, 1)``; used by tests and as the definition the CUDA packer is checked against)."""
    m = _phase(w.device, w.dtype)
    wp = torch.einsum("pak,qbl,oikl->pqaboi", m, m, w)
    return wp.reshape(4, 4, w.shape[0], w.shape[1])


_FOLD_CACHE: Dict = {}


def _fold_matrix(device, dtype) -> torch.Tensor:
    """``[9, 16]``: which of the 16 (phase, tap) gradients of the sub-pixel form add up to each 3x3 tap."""
    key = (str(device), dtype)
    if key not in _FOLD_CACHE:
        m = _PHASE                                                     # [p, a, k]
        f = torch.einsum("pak,qbl->klpqab", m, m).reshape(9, 16)
        _FOLD_CACHE[key] = f.to(device, dtype)
    return _FOLD_CACHE[key]


def unpack_up_grad(dwp: torch.Tensor) -> torch.Tensor:
    """adjoint of :func:`pack_up`: [4,4,Co,Ci] -> [Co,Ci,3,3] (one small GEMM with a constant 9x16 matrix)."""
    co, ci = dwp.shape[-2:]
    dw9 = _fold_matrix(dwp.device, dwp.dtype) @ dwp.reshape(16, co * ci)
    return dw9.reshape(3, 3, co, ci).permute(2, 3, 0, 1)


class PackPlan:
    """The operand packs of a fixed set of ``(weight, mode)`` pairs re-made by ONE kernel launch per step
    (``pack_weights_multi_kernel``; 26 launches per step before).  Outputs and the device-resident job table persist; the
    weights are addressed by pointer, so they must keep their storage (they live in the flat parameter arena)."""

    def __init__(self, items):
        self.items = list(items)
        self.es = ext().get_operand_size()
        self.outs, self.table, self.nblocks = ext().pack_plan_create([w for w, _ in self.items],
                                                                     [int(m) for _, m in self.items])
        self.index = {(_pack_key(w), int(m)): o for (w, m), o in zip(self.items, self.outs)}

    def run(self) -> None:
        ext().pack_plan_run(self.table, self.nblocks)
        _count()


_PACK_STATE = {"plan": None, "seen": None}


def _pack_key(w: torch.Tensor):
    return (w.data_ptr(), tuple(w.shape), tuple(w.stride()))


def pack_plan_enabled() -> bool:
    return not _emulated and os.environ.get("MINE_B200_PACK_PLAN", "1") == "1"


def pack(w: torch.Tensor, mode: int) -> torch.Tensor:
    """GEMM operand pack of a [Co,Ci,3,3] fp32 weight (any strides): served from the active :class:`PackPlan` (one launch
    per step for all layers) or made by its own kernel launch.
    mode 0/1: fprop same / upsample, 2/3: dgrad same / upsample (see csrc/conv_tcgen05.cu)."""
    w = w.detach()
    if w.dtype != torch.float32:
        w = w.float()
    plan = _PACK_STATE["plan"]
    if plan is not None:
        hit = plan.index.get((_pack_key(w), int(mode)))
        if hit is not None:
            return hit
    seen = _PACK_STATE["seen"]
    if seen is not None:
        seen.append((w, int(mode)))
    _count()
    return ext().pack_weights(w, mode)


SAME_TAPS_Y = [ky for ky in range(3) for kx in range(3)]
SAME_TAPS_X = [kx for ky in range(3) for kx in range(3)]
UP_TAPS_Y = [py + a for py in range(2) for px in range(2) for a in range(2) for b in range(2)]
UP_TAPS_X = [px + b for py in range(2) for px in range(2) for a in range(2) for b in range(2)]
UP_OY = [py for py in range(2) for px in range(2)]
UP_OX = [px for py in range(2) for px in range(2)]


# ---------------------------------------------------------------------------------------------
# raw kernel wrappers (no autograd)
# ---------------------------------------------------------------------------------------------
def conv_same_raw(xpad, w, out=None, chan_bias=None, plane_bias=None, shared_map=None, planes=1, stats=None,
                  head=False, head_alpha=False):
    """3x3 conv on a pre-padded input ``xpad [N,H+2,W+2,Ci]`` -> ``[N,H,W,Co]`` bf16 (or the packed fp32 MPI
    + sign tensor when ``head``)."""
    n, hp, wp_, ci = xpad.shape
    h, w_ = hp - 2, wp_ - 2
    co = w.shape[0]
    pack_ = pack(w, 0)
    th, tw = pick_tile(h, w_)
    if head:
        out = torch.empty((n, h, w_, 4), dtype=torch.float32, device=xpad.device)
        sign = torch.empty((n, h, w_), dtype=torch.int8, device=xpad.device)
        ext().conv_taps(xpad, pack_, out, h, w_, 1, 9, SAME_TAPS_Y, SAME_TAPS_X, 1, co, 1, 1, [0], [0], False,
                        chan_bias, None, None, 1, None, 1, head_alpha, sign, th, tw)
        _count()
        return out, sign
    if out is None:
        out = torch.empty((n, h, w_, co), dtype=ACT_DTYPE, device=xpad.device)
    ext().conv_taps(xpad, pack_, out, h, w_, 1, 9, SAME_TAPS_Y, SAME_TAPS_X, 1, co, 1, 1, [0], [0], False,
                    chan_bias, plane_bias, shared_map, planes, stats, 0, False, None, th, tw)
    _count()
    return out


def conv_up_raw(xpad_lo, w, chan_bias=None, plane_bias=None, shared_map=None, planes=1, stats=None):
    """``conv3x3(reflect_pad(nearest_up2(x)))`` from the replicate-padded low-res ``xpad_lo [N,h+2,w+2,Ci]``
    -> ``[N,2h,2w,Co]`` bf16 (4 sub-pixel phases x 4 taps)."""
    n, hp, wp_, ci = xpad_lo.shape
    h, w_ = hp - 2, wp_ - 2
    co = w.shape[0]
    pack_ = pack(w, 1)
    out = torch.empty((n, 2 * h, 2 * w_, co), dtype=ACT_DTYPE, device=xpad_lo.device)
    th, tw = pick_tile(h, w_)
    ext().conv_taps(xpad_lo, pack_, out, h, w_, 4, 4, UP_TAPS_Y, UP_TAPS_X, 1, co, 2, 2, UP_OY, UP_OX, False,
                    chan_bias, plane_bias, shared_map, planes, stats, 0, False, None, th, tw)
    _count()
    return out


def dgrad_same_raw(dy, w, accumulate_into=None):
    """Gradient w.r.t. the PADDED input of :func:`conv_same_raw`: ``dy [N,H,W,Co]`` -> ``[N,H+2,W+2,Ci]``
    (``accumulate_into``: add it onto an existing gradient of that shape inside the kernel epilogue)."""
    n, h, w_, co = dy.shape
    ci = w.shape[1]
    pack_ = pack(w, 2)                                                            # [tap][Ci][Co]
    acc = accumulate_into is not None
    out = accumulate_into if acc else torch.empty((n, h + 2, w_ + 2, ci), dtype=ACT_DTYPE, device=dy.device)
    th, tw = pick_tile(h + 2, w_ + 2)
    ext().conv_taps(dy, pack_, out, h + 2, w_ + 2, 1, 9, [-k for k in SAME_TAPS_Y], [-k for k in SAME_TAPS_X], 1, ci,
                    1, 1, [0], [0], acc, None, None, None, 1, None, 0, False, None, th, tw)
    _count()
    return out


class ReflectPadNHWC(torch.autograd.Function):
    """1-pixel reflection pad of a contiguous NHWC tensor, gradient by the adjoint kernel (``csrc/pad_nhwc.cu``): the padded
    skip feature and its gradient stay channels-last, so the library convolution around it needs no layout copies."""

    @staticmethod
    def forward(ctx, x):
        _count()
        return ext().pad_reflect_nhwc(x)

    @staticmethod
    def backward(ctx, gp):
        _count()
        return ext().pad_reflect_nhwc_bwd(gp.contiguous())


def skip_conv_own_pad() -> bool:
    return not _emulated and os.environ.get("MINE_B200_SKIP_PAD", "own") == "own"


class SplitWeight(torch.autograd.Function):
    """``w[:, :cp], w[:, cp:cp+cs], w[:, cp+cs:]`` (per-plane / shared-skip / embedding input channels of a factorised
    decoder conv) with ONE gradient buffer: backward writes the three parts into an ``empty_like(w)`` (3 strided copies).
    Plain slicing costs three zero-filled full-size tensors, three slice copies and two adds per block, and the summed
    gradient misses the parameter's channels-last layout (one more clone in AccumulateGrad)."""

    @staticmethod
    def forward(ctx, w, cp, cs):
        ctx.set_materialize_grads(False)
        ctx.split = (int(cp), int(cs))
        ctx.like = w
        return w[:, :cp], w[:, cp:cp + cs], w[:, cp + cs:]

    @staticmethod
    def backward(ctx, gp, gs, ge):
        cp, cs = ctx.split
        w = ctx.like
        g = torch.empty_like(w)                    # preserves the (channels-last) strides of the parameter
        for lo, hi, part in ((0, cp, gp), (cp, cp + cs, gs), (cp + cs, w.shape[1], ge)):
            if hi > lo:
                if part is None:
                    g[:, lo:hi].zero_()
                else:
                    g[:, lo:hi].copy_(part)
        return g, None, None


def split_block_weights(blk):
    """The three input-channel groups of a decoder block's conv weight (see :class:`SplitWeight`)."""
    w = blk.conv.conv.weight
    if blk.c_shared == 0 and blk.c_emb == 0:
        return w, w[:, :0], w[:, :0]
    if not (w.requires_grad and torch.is_grad_enabled()) or os.environ.get("MINE_B200_SPLIT_WEIGHT", "1") != "1":
        return blk.split_weights()
    return SplitWeight.apply(w, blk.c_plane, blk.c_shared)


class GradSlot:
    """Meeting point of the gradients of an activation with TWO consumers (a decoder level feeds its MPI head and the next
    level): the consumer whose backward runs first returns its data gradient to autograd and parks the tensor here; the
    second one ADDS its gradient onto that tensor inside the dgrad kernel epilogue and returns ``None`` - no separate
    ``add_`` pass over a [N, H+2, W+2, C] tensor (206 MB at level 1).  Both backward calls precede the producer's backward
    (autograd dependency order), so the in-place update is seen by it."""

    __slots__ = ("buf",)

    def __init__(self):
        self.buf = None

    def deliver(self, dgrad_fn):
        """``dgrad_fn(accumulate_into)`` -> gradient tensor; returns what the caller hands to autograd."""
        if self.buf is None:
            self.buf = dgrad_fn(None)
            return self.buf
        dgrad_fn(self.buf)
        self.buf = None
        return None


def grad_slots_enabled() -> bool:
    return os.environ.get("MINE_B200_GRAD_SLOT", "1") == "1"


def dgrad_up_raw(dy, w):
    """Gradient w.r.t. the replicate-padded LOW-res input of :func:`conv_up_raw`:
    ``dy [N,2h,2w,Co]`` -> ``[N,h+2,w+2,Ci]`` (16 strided taps, one accumulator)."""
    n, h2, w2, co = dy.shape
    h, w_ = h2 // 2, w2 // 2
    ci = w.shape[1]
    pack_ = pack(w, 3)                                                            # [g*4+t][Ci][Co]
    # forward: out[2y+py] += xpad[y + (py+a)]  =>  dxpad[q] += dy[2(q - py - a) + py] = dy[2q - py - 2a]
    ty = [-py - 2 * a for py in range(2) for px in range(2) for a in range(2) for b in range(2)]
    tx = [-px - 2 * b for py in range(2) for px in range(2) for a in range(2) for b in range(2)]
    out = torch.empty((n, h + 2, w_ + 2, ci), dtype=ACT_DTYPE, device=dy.device)
    th, tw = pick_tile(h + 2, w_ + 2)
    ext().conv_taps(dy, pack_, out, h + 2, w_ + 2, 1, 16, ty, tx, 2, ci, 1, 1, [0], [0], False, None, None, None, 1,
                    None, 0, False, None, th, tw)
    _count()
    return out


def _wgrad_pixels(co: int, ci: int, h: int, w: int) -> int:
    """Pixels per K step: aim at ~8 KB per TMA box (tiny boxes are issue-bound), keep the tile inside the map."""
    es = 4 if ACT_DTYPE == torch.float32 else 2
    kp = 8192 // (es * min(128 // es, max(co, ci)))
    kp = max(32, min(256, kp))
    while kp > 32 and kp > h * w:
        kp //= 2
    return kp


def wgrad_same_raw(dy, xpad):
    """``dW [Co,Ci,3,3]`` (fp32) of :func:`conv_same_raw`."""
    n, h, w_, co = dy.shape
    ci = xpad.shape[3]
    dw = torch.zeros((9, co, ci), dtype=torch.float32, device=dy.device)
    th, tw = pick_tile(h, w_, _wgrad_pixels(co, ci, h, w_))
    ext().wgrad_taps(dy, xpad, dw, h, w_, 1, 9, SAME_TAPS_Y, SAME_TAPS_X, 1, [0], [0], th, tw, 1)
    _count(2)
    return dw.reshape(3, 3, co, ci).permute(2, 3, 0, 1)


def wgrad_up_raw(dy, xpad_lo):
    """``dW [Co,Ci,3,3]`` (fp32) of :func:`conv_up_raw` (phase gradients folded back onto the 3x3 taps)."""
    n, h2, w2, co = dy.shape
    h, w_ = h2 // 2, w2 // 2
    ci = xpad_lo.shape[3]
    dwp = torch.zeros((16, co, ci), dtype=torch.float32, device=dy.device)
    th, tw = pick_tile(h, w_, min(128, _wgrad_pixels(co, ci, h, w_)))      # strided dy box: 2*TW <= 256
    ext().wgrad_taps(dy, xpad_lo, dwp, h, w_, 4, 4, UP_TAPS_Y, UP_TAPS_X, 2, UP_OY, UP_OX, th, tw, 1)
    _count(2)
    return unpack_up_grad(dwp.reshape(4, 4, co, ci))


# ---------------------------------------------------------------------------------------------
# fused layers with autograd
# ---------------------------------------------------------------------------------------------
class PlaneConvBNAct(torch.autograd.Function):
    """apad_out = pad(ELU(BN(conv(xpad) [+ shared_map] [+ plane_bias | chan_bias])))

    ``up``: conv is the fused upsample conv (xpad replicate-padded low-res).  ``pad_out``: 0 reflection,
    1 replication (what the consumer needs).  ``reducer``: cross-GPU SUM of small fp32 vectors.
    BatchNorm running statistics of ``bn`` are updated in place (training mode only).
    """

    @staticmethod
    def forward(ctx, xpad, w, chan_bias, plane_bias, shared_map, gamma, beta, up, planes, pad_out, bn, reducer,
                slot=None):
        co = w.shape[0]
        training = bn is None or bn.training
        stats = torch.zeros((2, co), dtype=torch.float32, device=xpad.device) if training else None
        cb = chan_bias.detach().float().contiguous() if chan_bias is not None else None
        pb = plane_bias.detach().float().contiguous() if plane_bias is not None else None
        sm = shared_map.detach().float().contiguous() if shared_map is not None else None
        fn = conv_up_raw if up else conv_same_raw
        y = fn(xpad, w, chan_bias=cb, plane_bias=pb, shared_map=sm, planes=planes, stats=stats)
        count = float(y.shape[0] * y.shape[1] * y.shape[2])
        fx = None
        if training:
            if reducer is not None:          # cross-replica statistics: one 2C-vector SUM ...
                count *= ctx_world(reducer)
                fx = fused_exchange(reducer, 2 * co)
                if fx is None:               # ... as a separate all-reduce launch
                    stats = reducer(stats.reshape(-1)).reshape(2, co).contiguous()
        else:                                # eval: normalise with the running statistics
            rm, rv = bn.running_mean.float(), bn.running_var.float()
            stats = torch.stack([rm * count, (rv + rm * rm) * count]).contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        if fx is not None:                   # ... or inside the normalise kernel (csrc/ll_exchange.cuh)
            apad, stats = ext().bn_act_pad_fwd_x(y, stats, g32, b32, int(pad_out), count, BN_EPS, *fx)
        else:
            apad = ext().bn_act_pad_fwd(y, stats, g32, b32, int(pad_out), count, BN_EPS)
        _count()
        if bn is not None and training:
            update_running_stats(bn, stats, count)
        ctx.save_for_backward(xpad, w, y, stats, g32, b32)
        ctx.cfg = (bool(up), int(planes), int(pad_out), count, reducer, chan_bias is not None, plane_bias is not None,
                   shared_map is not None)
        ctx.slot = slot if not up else None      # input shared with another consumer (GradSlot): same-resolution convs only
        return apad

    @staticmethod
    def backward(ctx, dapad):
        xpad, w, y, stats, g32, b32 = ctx.saved_tensors
        up, planes, pad_out, count, reducer, has_cb, has_pb, has_sm = ctx.cfg
        dapad = dapad.contiguous()
        dy, dshared, dpb, dgamma, dbeta = bn_act_backward(dapad, y, stats, g32, b32, pad_out, count, reducer, planes,
                                                          has_sm, has_pb)
        dcb = None
        if has_cb:
            # a bias in front of BatchNorm has an exactly-zero gradient (dy of a BatchNorm sums to zero over the batch);
            # autograd on the reference produces rounding noise around 0 here - return the exact value instead of
            # reducing a multi-hundred-MB tensor to get that noise
            dcb = torch.zeros(dy.shape[3], dtype=torch.float32, device=dy.device)
        dw = (wgrad_up_raw if up else wgrad_same_raw)(dy, xpad).to(w.dtype)
        dx = None
        if ctx.needs_input_grad[0]:
            if ctx.slot is not None:
                dx = ctx.slot.deliver(lambda acc: dgrad_same_raw(dy, w, accumulate_into=acc))
            else:
                dx = (dgrad_up_raw if up else dgrad_same_raw)(dy, w)
        return (dx, dw, dcb, dpb if has_pb else None, dshared if has_sm else None, dgamma.to(g32.dtype),
                dbeta.to(b32.dtype), None, None, None, None, None, None)


def bn_act_backward(dapad, y, stats, g32, b32, pad_out, count, reducer, planes, has_sm, has_pb):
    """Backward of ``pad(ELU(BN(y)))``: two kernels.  ``bn_act_bwd_reduce`` folds the pad adjoint, applies ELU', writes g
    and reduces ``[sum g, sum g * xhat]``; ``bn_bwd_apply`` writes dy (+ shared-map / plane-bias gradients), with the
    cross-replica SUM of the two reductions in its prologue when data parallel.  ``MINE_B200_BN_BWD=fused`` never writes g
    (``bn_act_bwd_sums`` + ``bn_bwd_apply_fused`` recompute it from the padded gradient): one tensor less through HBM,
    but measured SLOWER on B200 (level 0, tf32: 0.173 + 0.271 ms vs 0.218 + 0.189 ms - the halo-addressed read and the
    second exp() cost more than the saved write + read), so it stays opt-in.
    Returns ``(dy, dshared, dplane_bias, dgamma, dbeta)``."""
    fused_g = os.environ.get("MINE_B200_BN_BWD", "split") == "fused"
    if fused_g:
        g, sums = None, ext().bn_act_bwd_sums(dapad, y, stats, g32, b32, pad_out, count, BN_EPS)
    else:
        g, sums = ext().bn_act_bwd_reduce(dapad, y, stats, g32, b32, pad_out, count, BN_EPS)
    dgamma, dbeta = sums[1], sums[0]     # the LOCAL sums are the parameter gradients
    fx = fused_exchange(reducer, sums.numel())
    if fx is None and reducer is not None:   # separate all-reduce launch (in place: keep the local values first)
        dgamma, dbeta = sums[1].clone(), sums[0].clone()
        sums = reducer(sums.reshape(-1)).reshape(2, -1).contiguous()
    if fused_g:
        if fx is not None:
            out = ext().bn_bwd_apply_fused_x(dapad, y, stats, g32, b32, sums, planes, has_sm, has_pb, count, BN_EPS,
                                             pad_out, *fx)
        else:
            out = ext().bn_bwd_apply_fused(dapad, y, stats, g32, b32, sums, planes, has_sm, has_pb, count, BN_EPS, pad_out)
    elif fx is not None:                 # cross-replica SUM of the two reductions inside the apply kernel
        out = ext().bn_bwd_apply_x(g, y, stats, g32, sums, planes, has_sm, has_pb, count, BN_EPS, *fx)
    else:
        out = ext().bn_bwd_apply(g, y, stats, g32, sums, planes, has_sm, has_pb, count, BN_EPS)
    _count(2)
    return out[0], out[1], out[2], dgamma, dbeta


class BNActPad(torch.autograd.Function):
    """``apad = pad(ELU(BN(y)))`` for a pre-activation that did not come out of a conv kernel (decoder level 4_0: shared
    map + embedding bias): per-channel sums from ``channel_stats``, then the same kernels as :class:`PlaneConvBNAct`
    (statistic exchange fused into the normalise / apply kernels when data parallel)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, pad_out, bn, reducer):
        training = bn is None or bn.training
        co = y.shape[3]
        count = float(y.shape[0] * y.shape[1] * y.shape[2])
        fx = None
        if training:
            stats = ext().channel_stats(y)
            _count()
            if reducer is not None:
                count *= ctx_world(reducer)
                fx = fused_exchange(reducer, 2 * co)
                if fx is None:
                    stats = reducer(stats.reshape(-1)).reshape(2, co).contiguous()
        else:
            rm, rv = bn.running_mean.float(), bn.running_var.float()
            stats = torch.stack([rm * count, (rv + rm * rm) * count]).contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        if fx is not None:
            apad, stats = ext().bn_act_pad_fwd_x(y, stats, g32, b32, int(pad_out), count, BN_EPS, *fx)
        else:
            apad = ext().bn_act_pad_fwd(y, stats, g32, b32, int(pad_out), count, BN_EPS)
        _count()
        if bn is not None and training:
            update_running_stats(bn, stats, count)
        ctx.save_for_backward(y, stats, g32, b32)
        ctx.cfg = (int(pad_out), count, reducer)
        return apad

    @staticmethod
    def backward(ctx, dapad):
        y, stats, g32, b32 = ctx.saved_tensors
        pad_out, count, reducer = ctx.cfg
        dy, _, _, dgamma, dbeta = bn_act_backward(dapad.contiguous(), y, stats, g32, b32, pad_out, count, reducer, 1,
                                                  False, False)
        return dy, dgamma.to(g32.dtype), dbeta.to(b32.dtype), None, None, None


_RUNNING = {"defer": False, "pending": []}


def defer_running_stats(on: bool) -> None:
    """Trainer switch: collect the running-statistic updates of a step and apply them with one multi-layer launch
    (:func:`flush_running_stats`) instead of one small kernel per BatchNorm layer (67 per step with the hybrid encoder)."""
    if not on:
        flush_running_stats()
    _RUNNING["defer"] = bool(on)


def flush_running_stats() -> None:
    pend, _RUNNING["pending"] = _RUNNING["pending"], []
    if not pend:
        return
    with torch.no_grad():
        ext().bn_update_running_multi([p[1] for p in pend], [p[0].running_mean for p in pend],
                                      [p[0].running_var for p in pend], [p[0].num_batches_tracked for p in pend],
                                      [float(p[2]) for p in pend], [float(p[0].momentum) for p in pend])
    _count((len(pend) + 47) // 48)


def update_running_stats(bn, stats: torch.Tensor, count: float) -> None:
    """Momentum update of ``bn``'s buffers from the reduced batch sums ``[2, C]``: one kernel (``bn_update_running``), or
    queued for the multi-layer launch of :func:`flush_running_stats` when the trainer defers;
    ``MINE_B200_BN_RUNNING=aten`` selects the equivalent ~11 framework ops."""
    with torch.no_grad():
        if os.environ.get("MINE_B200_BN_RUNNING", "fused") == "fused" and (_emulated or stats.is_cuda):
            if _RUNNING["defer"] and stats.dtype == torch.float32 and stats.is_contiguous():
                _RUNNING["pending"].append((bn, stats, count))
                return
            ext().bn_update_running(stats, bn.running_mean, bn.running_var, bn.num_batches_tracked, float(count),
                                    float(bn.momentum))
            _count()
            return
        mean = stats[0] / count
        var = (stats[1] / count - mean * mean).clamp_min(0)
        m = bn.momentum
        bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
        bn.running_var.mul_(1 - m).add_(var * (count / max(count - 1.0, 1.0)), alpha=m)
        bn.num_batches_tracked += 1


def head_mode() -> str:
    """``MINE_B200_HEAD``: ``tcgen05`` (default: every head through conv_taps) or ``direct`` (CUDA-core kernel for the
    16 / 32-channel heads; opt-in until measured on hardware)."""
    return os.environ.get("MINE_B200_HEAD", "tcgen05")


def fused_exchange(reducer, numel: int):
    """Handle for running the cross-GPU statistic SUM inside the consuming kernel (P2P communicator with the
    low-latency protocol; ``MINE_B200_FUSED_BN=0`` disables), or ``None`` -> call ``reducer`` as a separate launch."""
    if reducer is None or _emulated:
        return None
    fn = getattr(getattr(reducer, "__self__", None), "fused_handle", None)
    h = fn() if fn is not None else None
    if h is None or numel > h[2]:
        return None
    return h


def ctx_world(reducer) -> int:
    return int(getattr(reducer, "world_size", None) or getattr(getattr(reducer, "__self__", None), "world_size", 1))


class HeadConv(torch.autograd.Function):
    """Packed fp32 MPI ``[N,H,W,4]`` = act(conv3x3(apad) + bias) with apad reflection padded."""

    @staticmethod
    def forward(ctx, apad, w, bias, use_alpha, slot=None):
        if head_mode() == "direct" and apad.shape[3] in (16, 32):
            # narrow full-resolution levels: bandwidth bound, CUDA-core kernel with one halo load (head_direct.cu)
            wpk = w.detach().float().permute(2, 3, 1, 0).contiguous()                 # [ky, kx, ci, co] = [9, C, 4]
            mpi, sign = ext().head_conv_direct(apad, wpk, bias.detach().float().contiguous(), bool(use_alpha))
            _count(2)
        else:
            mpi, sign = conv_same_raw(apad, w, chan_bias=bias.detach().float().contiguous(), head=True,
                                      head_alpha=use_alpha)
        ctx.save_for_backward(apad, w, mpi, sign)
        ctx.use_alpha = bool(use_alpha)
        ctx.slot = slot
        return mpi

    @staticmethod
    def backward(ctx, g_mpi):
        apad, w, mpi, sign = ctx.saved_tensors
        dz, dbias = ext().head_bwd(g_mpi.contiguous().float(), mpi, sign, ctx.use_alpha)
        _count()
        w16 = F.pad(w.detach(), (0, 0, 0, 0, 0, 0, 0, 16 - w.shape[0]))           # Co 4 -> 16 (zero rows)
        dw = wgrad_same_raw(dz, apad)[: w.shape[0]].to(w.dtype)
        dx = None
        if ctx.needs_input_grad[0]:
            if ctx.slot is not None:
                dx = ctx.slot.deliver(lambda acc: dgrad_same_raw(dz, w16, accumulate_into=acc))
            else:
                dx = dgrad_same_raw(dz, w16)
        return dx, dw, dbias.to(w.dtype), None, None


# ---------------------------------------------------------------------------------------------
# decoder driver
# ---------------------------------------------------------------------------------------------
def pad_nhwc(a: torch.Tensor, mode: str) -> torch.Tensor:
    """1-pixel pad of an NHWC tensor with torch ops (used for the tiny level-4 input only)."""
    x = a.permute(0, 3, 1, 2)
    x = F.pad(x.float(), (1, 1, 1, 1), mode="reflect" if mode == "reflect" else "replicate").to(a.dtype)
    return x.permute(0, 2, 3, 1).contiguous()


class ConvEngine:
    """Runs encoder (library convs, bf16 channels_last - 4 % of the FLOPs) and the factorised decoder
    with every per-plane convolution on the tcgen05 kernels."""

    def __init__(self, backbone, decoder, config, device, encoder_mode: str | None = None):
        self.backbone, self.decoder, self.config, self.device = backbone, decoder, config, device
        # encoder: "cudnn" (library convolutions + ATen BN under bf16 autocast), "hybrid" (library convolutions +
        # our fused BN kernels) or "tcgen05" (everything on the engine, encoder_engine.py)
        # default: one GPU -> "cudnn" (library convolutions + library BatchNorm: bitwise reproducible from run to run);
        # data parallel -> "hybrid" (BatchNorm / residual / ReLU of the encoder on our kernels: statistics kernel + one
        # fused normalise-residual-activation kernel per layer with the cross-replica exchange inside it; the library
        # sync-BN path costs +1.7 ms per step at 2 GPUs).  Measured on one B200, LLFF step: tf32 hybrid 11.52 ms vs cudnn
        # 11.59 ms, bf16 hybrid 9.81 ms vs cudnn 10.9 ms - hybrid is not the one-GPU default only because its fp32-atomic
        # statistics vary in the last bits between runs, and ~50 BatchNorm layers over a few dozen samples each amplify
        # that (scripts/graph_vs_eager_probe.py); ``engine.deterministic`` makes them reproducible at ~+1 ms per step.
        import torch.distributed as _dist
        multi = _dist.is_available() and _dist.is_initialized() and _dist.get_world_size() > 1
        self.encoder_mode = encoder_mode or os.environ.get("MINE_B200_ENCODER", "hybrid" if multi else "cudnn")
        if self.encoder_mode not in ("cudnn", "tcgen05", "hybrid"):
            raise ValueError("MINE_B200_ENCODER must be cudnn, hybrid or tcgen05, got %r" % self.encoder_mode)
        self.encoder_engine = None
        if self.encoder_mode != "cudnn":
            from .encoder_engine import EncoderEngine
            self.encoder_engine = EncoderEngine(backbone, library_conv=self.encoder_mode == "hybrid")

    def _reducer(self):
        from ..models.norm import BatchNorm
        for m in self.decoder.modules():
            if isinstance(m, BatchNorm):
                return m.reducer
        return None

    def _begin_packs(self, device) -> None:
        """Batched weight packing: the first training step records which ``(weight, mode)`` packs the model asks for (forward
        and backward); from the second step on one launch re-makes all of them before the encoder runs."""
        if not pack_plan_enabled() or device.type != "cuda":
            _PACK_STATE["plan"] = _PACK_STATE["seen"] = None
            return
        plan = getattr(self, "_pack_plan", None)
        if plan is not None and plan.es != ext().get_operand_size():
            plan = self._pack_plan = None                               # precision changed: other operand type
        if plan is None and not torch.cuda.is_current_stream_capturing():
            seen = getattr(self, "_pack_seen", None)
            if seen:
                # only weights that live in parameter storage (persistent addresses), not per-step temporaries
                homes = {q.untyped_storage().data_ptr() for mod in (self.backbone, self.decoder) for q in mod.parameters()}
                uniq = {}
                for w, m in seen:
                    if w.untyped_storage().data_ptr() in homes:
                        uniq.setdefault((_pack_key(w), m), (w, m))
                if uniq:
                    plan = self._pack_plan = PackPlan(uniq.values())
                self._pack_seen = None if uniq else []
            else:
                self._pack_seen = []
        _PACK_STATE["plan"] = plan
        _PACK_STATE["seen"] = getattr(self, "_pack_seen", None) if plan is None else None
        if plan is not None:
            plan.run()

    def predict(self, src_imgs: torch.Tensor, disparity: torch.Tensor) -> List[torch.Tensor]:
        dec = self.decoder
        b, s = disparity.shape
        n = b * s
        self._begin_packs(src_imgs.device)
        amp = dict(device_type=src_imgs.device.type, dtype=torch.bfloat16,
                   enabled=src_imgs.is_cuda and ACT_DTYPE == torch.bfloat16)
        if self.encoder_engine is not None:
            feats = self.encoder_engine(src_imgs)
        else:
            with torch.autocast(**amp):
                feats = self.backbone(src_imgs.contiguous(memory_format=torch.channels_last))
        own_convs = self.encoder_mode == "tcgen05"          # no library convolution anywhere on the prediction path
        if own_convs or self.encoder_mode == "hybrid":
            from . import encoder_engine as EE
            top = EE.receptive_field_extension(dec, feats[-1], self._reducer(), library_conv=not own_convs)
        else:
            with torch.autocast(**amp):
                top = dec.receptive_field_extension(feats[-1])
        emb = dec.embed(disparity).float()                                       # [N, E]
        reducer = self._reducer()
        use_alpha = bool(dec.use_alpha)

        def shared_and_bias(blk, feat):
            wp, ws, we = split_block_weights(blk)
            bias = blk.conv.conv.bias
            smap = None
            if feat is not None and own_convs:
                smap = EE.shared_skip_map(feat, ws)                               # [B,H,W,Co] fp32
            elif feat is not None:
                f_nhwc = feat.permute(0, 2, 3, 1)
                if (skip_conv_own_pad() and feat.is_cuda and f_nhwc.is_contiguous() and feat.shape[1] % 8 == 0
                        and feat.dtype in (torch.float32, torch.bfloat16) and min(feat.shape[2:]) >= 2):
                    # NHWC pad kernel + dense channels-last weight slice: the library convolution runs without a single
                    # layout copy in either direction (framework pad: 3 copies + 2 pad kernels per level and step)
                    padded = ReflectPadNHWC.apply(f_nhwc).permute(0, 3, 1, 2)
                    with torch.autocast(**amp):
                        smap = F.conv2d(padded, ws.contiguous(memory_format=torch.channels_last))
                else:
                    with torch.autocast(**amp):
                        smap = F.conv2d(F.pad(feat, (1, 1, 1, 1), mode="reflect"), ws)
                smap = smap.permute(0, 2, 3, 1).float().contiguous()            # [B,H,W,Co] fp32
            if blk.c_emb > 0:
                pbias = emb @ we.float().sum(dim=(2, 3)).t() + bias.float()[None]
                return wp, smap, None, pbias
            return wp, smap, bias, None

        # level 4_0: shared + embedding only (no per-plane input) - tiny, plain torch
        blk = dec.blocks["upconv_4_0"]
        _, smap, _, pbias = shared_and_bias(blk, top)
        y40 = smap[:, None] + pbias.reshape(b, s, 1, 1, -1)                        # [B,S,h,w,C]
        y40 = y40.reshape(n, *smap.shape[1:])
        c40 = y40.shape[3]
        if (src_imgs.is_cuda or _emulated) and c40 >= 16 and (c40 & (c40 - 1)) == 0:
            # BN + ELU + replication pad as one kernel pair (cross-GPU statistics exchanged inside)
            xpad = BNActPad.apply(y40.to(ACT_DTYPE).contiguous(), blk.bn.weight, blk.bn.bias, 1, blk.bn, reducer)
        else:
            a = F.elu(blk.bn(y40.permute(0, 3, 1, 2))).permute(0, 2, 3, 1)          # NHWC fp32
            xpad = pad_nhwc(to_operand(a), "replicate")                      # feeds the upsample conv

        outputs: Dict[int, torch.Tensor] = {}
        slot = None                                  # set when the current ``xpad`` also fed an MPI head (two consumers)
        for i in range(4, -1, -1):
            if i < 4:
                blk = dec.blocks[f"upconv_{i}_0"]
                wp, _, cb, _ = shared_and_bias(blk, None)
                xpad = PlaneConvBNAct.apply(xpad, wp, cb, None, None, blk.bn.weight, blk.bn.bias, False, s, 1,
                                            blk.bn, reducer, slot)
            slot = None
            blk = dec.blocks[f"upconv_{i}_1"]
            feat = feats[i - 1] if (dec.use_skips and i > 0) else None
            wp, smap, cb, pbias = shared_and_bias(blk, feat)
            xpad = PlaneConvBNAct.apply(xpad, wp, cb, pbias, smap, blk.bn.weight, blk.bn.bias, True, s, 0,
                                        blk.bn, reducer)
            if i in dec.scales:
                head = dec.heads[f"dispconv_{i}"]
                if i > 0 and grad_slots_enabled() and torch.is_grad_enabled() and xpad.requires_grad:
                    slot = GradSlot()                # head_i and upconv_{i-1}_0 both read xpad: gradients meet in the kernel
                mpi = HeadConv.apply(xpad, head.conv.weight, head.conv.bias, use_alpha, slot)
                outputs[i] = mpi.reshape(b, s, *mpi.shape[1:])
        return [outputs[k] for k in range(4)]
