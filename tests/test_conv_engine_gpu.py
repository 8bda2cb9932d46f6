"""tcgen05 conv engine vs plain PyTorch fp32 references (forward, dgrad, wgrad, fused layers), in both operand
precisions: ``tf32`` (fp32 storage, kind::tf32 MMAs - the default, reference numerics class) and ``bf16``.

Operands are pre-rounded to the operand format (bf16 rounding / TF32 truncation to 10 mantissa bits), so products are
exact in fp32 and the comparison isolates the kernel: in tf32 mode single-layer results must match the fp32 reference
to ~1e-4, bf16 results to the bf16 output rounding (~1e-2)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["tf32", "bf16"])
def precision(request):
    from mine_b200.ops import conv_engine as E
    old = E.PRECISION
    E.set_precision(request.param)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the reference side of every comparison is true fp32
    yield request.param
    torch.backends.cudnn.allow_tf32 = prev
    E.set_precision(old)


def _nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def _nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2)


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def _is_tf32():
    from mine_b200.ops import conv_engine as E
    return E.ACT_DTYPE == torch.float32


def _q(x):
    """Round to the operand format and return fp32: bf16 rounding, or TF32 (truncate to 10 explicit mantissa bits)."""
    if _is_tf32():
        return (x.float().contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    return x.to(torch.bfloat16).float()


def _act(x_nchw):
    """NCHW fp32 -> the engine's NHWC operand tensor."""
    from mine_b200.ops import conv_engine as E
    return _nhwc(x_nchw).to(E.ACT_DTYPE)


def _tol(bf16, tf32):
    return tf32 if _is_tf32() else bf16


_bf = _q


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


def _rel2(a, b):
    return (a.float() - b.float()).norm().item() / (b.float().norm().item() + 1e-12)


SHAPES = [(2, 16, 24, 64, 32), (3, 8, 12, 256, 256), (2, 40, 56, 16, 16), (1, 32, 48, 128, 64), (2, 20, 36, 32, 16)]


@pytest.mark.parametrize("n,h,w,ci,co", SHAPES)
def test_conv_same_fprop(n, h, w, ci, co):
    from mine_b200.ops import conv_engine as E
    x = _bf(_rand((n, ci, h, w), 0))
    wt = _bf(_rand((co, ci, 3, 3), 1, 0.1))
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    ref = F.conv2d(xp, wt)
    stats = torch.zeros(2, co, device="cuda")
    pb = _rand((n, co), 2)
    sm = _rand((1, h, w, co), 3)
    y = E.conv_same_raw(_act(xp), wt, plane_bias=pb, shared_map=sm, planes=n, stats=stats)
    ref = ref + pb[:, :, None, None] + _nchw(sm)
    assert _rel(_nchw(y), ref) < _tol(1e-2, 2e-4)
    assert torch.allclose(stats[0], ref.sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2 * ref.abs().max().item() * 10)
    assert torch.allclose(stats[1], (ref * ref).sum(dim=(0, 2, 3)), rtol=2e-3)


@pytest.mark.parametrize("n,h,w,ci,co", SHAPES)
def test_conv_up_fprop(n, h, w, ci, co):
    from mine_b200.ops import conv_engine as E
    x = _bf(_rand((n, ci, h, w), 0))
    wt = _bf(_rand((co, ci, 3, 3), 1, 0.1))
    up = F.interpolate(x, scale_factor=2, mode="nearest")
    ref = F.conv2d(F.pad(up, (1, 1, 1, 1), mode="reflect"), wt)
    xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
    y = E.conv_up_raw(_act(xp), wt)
    assert tuple(y.shape) == (n, 2 * h, 2 * w, co)
    assert _rel(_nchw(y), ref) < _tol(1.5e-2, 1e-3)     # phase weights are tap sums: not TF32-representable


@pytest.mark.parametrize("n,h,w,ci,co", SHAPES)
def test_dgrad_and_wgrad_same(n, h, w, ci, co):
    from mine_b200.ops import conv_engine as E
    xp = _bf(_rand((n, ci, h + 2, w + 2), 0)).requires_grad_(True)
    wt = _bf(_rand((co, ci, 3, 3), 1, 0.1)).requires_grad_(True)
    dy = _bf(_rand((n, co, h, w), 2))
    F.conv2d(xp, wt).backward(dy)
    dx = E.dgrad_same_raw(_act(dy), wt.detach())
    assert _rel(_nchw(dx), xp.grad) < _tol(1.5e-2, 2e-4)
    dw = E.wgrad_same_raw(_act(dy), _act(xp.detach()))
    assert _rel(dw, wt.grad) < _tol(1e-2, 2e-4)


@pytest.mark.parametrize("n,h,w,ci,co", SHAPES)
def test_dgrad_and_wgrad_up(n, h, w, ci, co):
    from mine_b200.ops import conv_engine as E
    xp = _bf(_rand((n, ci, h + 2, w + 2), 0)).requires_grad_(True)        # replicate-padded low-res (free variable)
    wt = _bf(_rand((co, ci, 3, 3), 1, 0.1)).requires_grad_(True)
    dy = _bf(_rand((n, co, 2 * h, 2 * w), 2))
    # reference through the explicit phase formulation (exact identity tested in test_conv_up_fprop)
    wp = E.pack_up(wt)                                                    # [4,4,Co,Ci]
    out = torch.zeros(n, co, 2 * h, 2 * w, device="cuda")
    for py in range(2):
        for px in range(2):
            acc = 0
            for a in range(2):
                for b in range(2):
                    win = xp[:, :, py + a: py + a + h, px + b: px + b + w]
                    acc = acc + torch.einsum("nihw,oi->nohw", win, wp[py * 2 + px, a * 2 + b])
            out[:, :, py::2, px::2] = acc
    out.backward(dy)
    dx = E.dgrad_up_raw(_act(dy), wt.detach())
    assert _rel(_nchw(dx), xp.grad) < _tol(1.5e-2, 1e-3)
    dw = E.wgrad_up_raw(_act(dy), _act(xp.detach()))
    assert _rel(dw, wt.grad) < _tol(1e-2, 2e-4)


@pytest.mark.parametrize("n,h,w,co,ci", [(2, 16, 24, 16, 32), (3, 9, 13, 16, 16), (2, 8, 16, 128, 64)])
def test_dgrad_accumulates_into_existing_gradient(n, h, w, co, ci):
    """Two consumers of one activation: the second data gradient is added onto the first inside the kernel epilogue
    (``GradSlot``) - must equal the separate sum."""
    from mine_b200.ops import conv_engine as E
    dy = _act(_q(_rand((n, co, h, w), 0)))
    wt = _q(_rand((co, ci, 3, 3), 1, 0.1))
    base = _act(_q(_rand((n, ci, h + 2, w + 2), 2)))
    want = base.float() + E.dgrad_same_raw(dy, wt).float()
    got = E.dgrad_same_raw(dy, wt, accumulate_into=base.clone())
    assert _rel(got.float(), want) < _tol(1e-2, 1e-6)
    slot = E.GradSlot()
    first = slot.deliver(lambda acc: E.dgrad_same_raw(dy, wt, accumulate_into=acc))
    assert first is not None and slot.buf is first
    second = slot.deliver(lambda acc: E.dgrad_same_raw(dy, wt, accumulate_into=acc))
    assert second is None and slot.buf is None
    assert _rel(first.float(), 2 * E.dgrad_same_raw(dy, wt).float()) < _tol(1e-2, 1e-6)


@pytest.mark.parametrize("pad_mode", [0, 1])
@pytest.mark.parametrize("n,h,w,c", [(4, 12, 20, 32), (2, 33, 17, 16), (6, 8, 12, 256)])
def test_bn_act_pad_fwd_bwd(pad_mode, n, h, w, c):
    from mine_b200.ops import conv_engine as E
    ext = E.ext()
    y = _bf(_rand((n, c, h, w), 0) * 2 + 0.3)
    gamma, beta = _rand((c,), 1).abs() + 0.5, _rand((c,), 2) * 0.2
    stats = torch.stack([y.sum(dim=(0, 2, 3)), (y * y).sum(dim=(0, 2, 3))]).contiguous()
    count = float(n * h * w)
    yr = y.clone().requires_grad_(True)
    g_r, b_r = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bn = F.batch_norm(yr, None, None, g_r, b_r, training=True, eps=1e-5)
    ref = F.pad(F.elu(bn), (1, 1, 1, 1), mode="reflect" if pad_mode == 0 else "replicate")
    apad = ext.bn_act_pad_fwd(_act(y), stats, gamma, beta, pad_mode, count, 1e-5)
    assert _rel(_nchw(apad), ref) < _tol(1e-2, 6e-4)        # the stored activation is rounded to TF32 (2^-11)
    dap = _bf(_rand(ref.shape, 3))
    ref.backward(dap)
    g, sums = ext.bn_act_bwd_reduce(_act(dap), _act(y), stats, gamma, beta,
                                    pad_mode, count, 1e-5)
    assert _rel(sums[0], b_r.grad) < _tol(2e-2, 1e-3) and _rel(sums[1], g_r.grad) < _tol(2e-2, 1e-3)
    dy, dsh, dpb = ext.bn_bwd_apply(g, _act(y), stats, gamma, sums, n // 2, True, True, count, 1e-5)
    assert _rel(_nchw(dy), yr.grad) < _tol(2e-2, 1e-3)
    ref_dsh = yr.grad.reshape(2, n // 2, c, h, w).sum(1).permute(0, 2, 3, 1)
    assert _rel(dsh, ref_dsh) < _tol(2e-2, 1e-3)
    assert _rel(dpb, yr.grad.sum(dim=(2, 3))) < _tol(2e-2, 1e-3)
    # default backward form: the activation gradient is never materialised (sums only + apply from the padded gradient)
    sums2 = ext.bn_act_bwd_sums(_act(dap), _act(y), stats, gamma, beta, pad_mode, count, 1e-5)
    assert _rel(sums2, sums) < 1e-5
    dy2, dsh2, dpb2 = ext.bn_bwd_apply_fused(_act(dap), _act(y), stats, gamma, beta, sums, n // 2, True, True, count, 1e-5,
                                             pad_mode)
    assert _rel(_nchw(dy2), yr.grad) < _tol(2e-2, 1e-3)
    assert _rel(dy2, dy) < _tol(1e-2, 1e-5) and _rel(dsh2, dsh) < _tol(1e-2, 1e-5) and _rel(dpb2, dpb) < _tol(1e-2, 1e-4)


def test_fused_layer_and_head_autograd():
    """Upsample-conv + BN + ELU layer followed by the MPI head, all gradients vs fp32 autograd - INCLUDING the sigma
    channel (|x| + 1e-4, the gradient that trains geometry).  |x| is non-smooth at 0, so the upstream sigma gradient
    is masked where the reference pre-activation is within eps of zero (a sign flip there is rounding, not a bug)."""
    from mine_b200.models.norm import BatchNorm
    from mine_b200.ops import conv_engine as E
    n, s, h, w, ci, co = 4, 2, 12, 16, 32, 16
    a = _q(_rand((n, ci, h, w), 0))
    wt = _q(_rand((co, ci, 3, 3), 1, 0.1)).requires_grad_(True)
    wh = _q(_rand((4, co, 3, 3), 2, 0.1)).requires_grad_(True)
    bh = _rand((4,), 3, 0.1).requires_grad_(True)
    gamma, beta = (_rand((co,), 4).abs() + 0.5).requires_grad_(True), (_rand((co,), 5) * 0.1).requires_grad_(True)
    pbias = _rand((n, co), 6).requires_grad_(True)
    smap = _rand((n // s, 2 * h, 2 * w, co), 7).requires_grad_(True)
    xlo = a.clone().requires_grad_(True)
    # reference: fp32 math on operand-format inputs
    up = F.interpolate(xlo, scale_factor=2, mode="nearest")
    y = F.conv2d(F.pad(up, (1, 1, 1, 1), mode="reflect"), wt)
    y = y + pbias[:, :, None, None] + _nchw(smap).repeat_interleave(s, dim=0)
    act = F.elu(F.batch_norm(y, None, None, gamma, beta, training=True, eps=1e-5))
    z = F.conv2d(F.pad(act, (1, 1, 1, 1), mode="reflect"), wh, bh)
    mpi_ref = torch.cat([torch.sigmoid(z[:, :3]), z[:, 3:].abs() + 1e-4], 1)
    gout = _rand(mpi_ref.shape, 8)
    eps = _tol(5e-2, 2e-3)
    gout[:, 3] = gout[:, 3] * (z[:, 3].detach().abs() > eps)     # non-zero sigma gradient away from the kink
    assert (gout[:, 3] != 0).float().mean().item() > 0.8
    (mpi_ref * gout).sum().backward()
    ref_grads = [t.grad.clone() for t in (xlo, wt, wh, bh, gamma, beta, pbias, smap)]
    for t in (xlo, wt, wh, bh, gamma, beta, pbias, smap):
        t.grad = None
    # engine
    xlo2 = xlo.detach().clone().requires_grad_(True)
    xpad = E.pad_nhwc(_act(xlo2), "replicate")
    bn = BatchNorm(co).cuda()
    apad = E.PlaneConvBNAct.apply(xpad, wt, None, pbias, smap, gamma, beta, True, s, 0, bn, None)
    mpi = E.HeadConv.apply(apad, wh, bh, False)
    assert _rel(mpi.permute(0, 3, 1, 2), mpi_ref) < _tol(3e-2, 2e-3)
    (mpi.permute(0, 3, 1, 2) * gout).sum().backward()
    got = [xlo2.grad, wt.grad, wh.grad, bh.grad, gamma.grad, beta.grad, pbias.grad, smap.grad]
    names = ["dx", "dW", "dWhead", "dbhead", "dgamma", "dbeta", "dplane_bias", "dshared"]
    errs = {nme: _rel2(g_, r_) for nme, g_, r_ in zip(names, got, ref_grads)}
    print('fused layer grad errors', errs)
    assert all(v < _tol(8e-2, 5e-3) for v in errs.values()), str(errs)


def test_decoder_engine_matches_fp32_module():
    """Whole decoder on the tcgen05 engine vs the PyTorch module in true fp32 (same fp32 encoder features): forward
    MPIs and every weight gradient, per-parameter bound 3e-2 in tf32 mode (bf16: loose sanity bounds, the operand
    rounding of a 10-layer random-init network dominates)."""
    from mine_b200.models.decoder import DepthDecoder
    from mine_b200.models.encoder import ResnetEncoder
    from mine_b200.ops.conv_engine import ConvEngine
    torch.manual_seed(0)
    enc, dec = ResnetEncoder(pretrained=False).cuda(), DepthDecoder().cuda()
    b, s, h, w = 2, 4, 256, 256
    img = torch.rand(b, 3, h, w, device="cuda")
    disp = torch.rand(b, s, device="cuda") * 0.8 + 0.1
    eng = ConvEngine(enc, dec, {}, torch.device("cuda"), encoder_mode="cudnn")
    outs = eng.predict(img, disp)
    if _is_tf32():
        with torch.no_grad():
            feats = enc(img)
        ref = dec(feats, disp)
    else:                                  # bf16 mode: the library in the same precision is the meaningful reference
        with torch.autocast("cuda", dtype=torch.bfloat16):
            feats = enc(img.contiguous(memory_format=torch.channels_last))
            ref = dec(feats, disp)
    refs = [ref[("disp", k)].float().permute(0, 1, 3, 4, 2) for k in range(4)]
    gouts = [torch.randn_like(o) for o in outs]
    for g, r in zip(gouts, refs):          # sigma gradient away from the |x| kink only (sigma = |x| + 1e-4)
        g[..., 3] = g[..., 3] * (r[..., 3].detach() > 1e-2) if _is_tf32() else 0
    sum((o * g).sum() for o, g in zip(outs, gouts)).backward()
    got = {k: p.grad.clone() for k, p in dec.named_parameters() if p.grad is not None}
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    for k in range(4):
        assert tuple(outs[k].shape) == tuple(refs[k].shape)
        assert _rel2(outs[k], refs[k]) < _tol(3e-2, 1e-2), (k, _rel2(outs[k], refs[k]))
    sum((o * g).sum() for o, g in zip(refs, gouts)).backward()
    bad, worst = [], 0.0
    for kname, p in dec.named_parameters():
        if p.grad is None or kname not in got or kname.endswith("conv.conv.bias"):
            continue        # conv biases in front of BatchNorm have an exactly-zero true gradient (pure noise)
        r = _rel2(got[kname], p.grad)
        worst = max(worst, r)
        if r > _tol(0.15, 3e-2):
            bad.append((kname, round(r, 4)))
    print("decoder engine vs fp32 module: worst per-parameter gradient error", worst)
    assert not bad, bad[:10]


def test_kernels_match_emulator_contract():
    """Same arguments through the sm_100a kernels and through the PyTorch specification (``ops/emu.py``) that the
    CPU tier uses to test the engine's orchestration: fprop with every epilogue term, strided phase dgrad, wgrad."""
    from mine_b200.ops import conv_engine as E
    n, h, w, ci, co = 4, 12, 20, 32, 64
    xlo = _q(_rand((n, h + 2, w + 2, ci), 0)).to(E.ACT_DTYPE)
    wt = _q(_rand((co, ci, 3, 3), 1, 0.1))
    pb, sm = _rand((n, co), 2), _rand((2, 2 * h, 2 * w, co), 3)
    dy = _q(_rand((n, 2 * h, 2 * w, co), 4)).to(E.ACT_DTYPE)

    def run():
        stats = torch.zeros(2, co, device="cuda")
        y = E.conv_up_raw(xlo, wt, plane_bias=pb, shared_map=sm, planes=2, stats=stats)
        return y, stats, E.dgrad_up_raw(dy, wt), E.wgrad_up_raw(dy, xlo)
    real = run()
    E.use_emulator(True, E.ACT_DTYPE, tf32_operands=True)
    try:
        spec = run()
    finally:
        E.use_emulator(False)
    assert _rel(real[0], spec[0]) < _tol(1e-2, 1e-3) and _rel(real[2], spec[2]) < _tol(1e-2, 1e-3)
    assert torch.allclose(real[1], spec[1], rtol=2e-3, atol=1.0)
    assert _rel2(real[3], spec[3]) < _tol(2e-3, 2e-4)
