"""ResNet encoder on the tcgen05 engine vs PyTorch references (wide channel-blocked layers, stride-2 forms,
encoder elementwise kernels, whole trunk), in both operand precisions (``tf32`` default, ``bf16``).  Operands are
pre-rounded to the operand format so single-layer comparisons isolate the kernel (see test_conv_engine_gpu.py)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu]


@pytest.fixture(autouse=True, params=["tf32", "bf16"])
def precision(request):
    from mine_b200.ops import conv_engine as E
    old = E.PRECISION
    E.set_precision(request.param)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield request.param
    torch.backends.cudnn.allow_tf32 = prev
    E.set_precision(old)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _rand(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).cuda()


def _is_tf32():
    from mine_b200.ops import conv_engine as E
    return E.ACT_DTYPE == torch.float32


def _bf(x):
    """Round to the operand format (bf16 rounding or TF32 truncation), returned as fp32."""
    if _is_tf32():
        return (x.float().contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    return x.to(_ACT()).float()


def _op(x):
    from mine_b200.ops import conv_engine as E
    return x.to(E.ACT_DTYPE)


def _ACT():
    from mine_b200.ops import conv_engine as E
    return E.ACT_DTYPE


def _tol(bf16, tf32):
    return tf32 if _is_tf32() else bf16


def _rel2(a, b):
    return (a.float() - b.float()).norm().item() / (b.float().norm().item() + 1e-12)


CASES = [  # k, stride, n, h, w, ci, co
    (1, 1, 2, 16, 24, 64, 256), (1, 1, 2, 8, 12, 2048, 512), (1, 1, 2, 8, 12, 512, 2048), (3, 1, 2, 16, 24, 128, 128),
    (3, 1, 2, 8, 12, 512, 512), (1, 2, 2, 16, 24, 256, 512), (3, 2, 2, 16, 24, 256, 256), (3, 2, 2, 32, 48, 128, 128),
    (1, 2, 2, 32, 48, 256, 1024)]


@pytest.mark.parametrize("k,stride,n,h,w,ci,co", CASES)
def test_encoder_conv_directions(k, stride, n, h, w, ci, co):
    from mine_b200.ops import encoder_engine as EE
    x = _bf(_rand((n, ci, h, w), 0)).requires_grad_()
    wt = _bf(_rand((co, ci, k, k), 1, (ci * k * k) ** -0.5)).requires_grad_()
    ref = F.conv2d(x, wt, None, stride, k // 2)
    stats = torch.zeros(2, co, device="cuda")
    y = EE.conv_fprop(_nhwc(x.detach()).to(_ACT()), wt.detach(), stride, stats)
    assert _rel2(_nchw(y), ref) < _tol(6e-3, 1e-4)
    assert torch.allclose(stats[0], ref.sum(dim=(0, 2, 3)), rtol=1e-2, atol=0.05 * ref.abs().sum(dim=(0, 2, 3)).max().item())
    assert _rel2(stats[1], (ref * ref).sum(dim=(0, 2, 3))) < 5e-3
    dy = _bf(_rand(ref.shape, 2))
    ref.backward(dy)
    dyb = _nhwc(dy).to(_ACT())
    assert _rel2(_nchw(EE.conv_dgrad(dyb, wt.detach(), stride, h, w)), x.grad) < _tol(6e-3, 1e-4)
    assert _rel2(EE.conv_wgrad(dyb, _nhwc(x.detach()).to(_ACT()), k, stride), wt.grad) < _tol(6e-3, 1e-4)


@pytest.mark.parametrize("c,relu,res", [(64, 0.0, False), (256, 0.0, True), (2048, 1.0, False), (1024, 0.1, True)])
def test_encoder_elementwise_match_specification(c, relu, res):
    """``relu`` is the activation slope: 0 = ReLU, 0.1 = LeakyReLU, 1 = identity."""
    from mine_b200.ops import conv_engine as E
    from mine_b200.ops import emu
    n, h, w = 2, 12, 20
    y = _rand((n, h, w, c), 0).to(_ACT())
    r = _rand((n, h, w, c), 1).to(_ACT()) if res else None
    dout = _rand((n, h, w, c), 2).to(_ACT())
    gamma, beta = torch.rand(c, device="cuda") + 0.5, _rand((c,), 3)
    count = float(n * h * w)
    ext = E.ext()
    stats = ext.channel_stats(y)
    want = emu.channel_stats(y)
    assert torch.allclose(stats, want, rtol=1e-3, atol=1e-2)
    out = ext.bn_res_act_fwd(y, stats, gamma, beta, r, relu, count, 1e-5)
    assert _rel2(out, emu.bn_res_act_fwd(y, want, gamma, beta, r, relu, count, 1e-5)) < 5e-3
    g, sums = ext.bn_res_act_bwd_reduce(dout, out, y, stats, gamma, beta, relu, count, 1e-5)
    g2, sums2 = emu.bn_res_act_bwd_reduce(dout, out, y, stats, gamma, beta, relu, count, 1e-5)
    assert _rel2(g, g2) < 1e-3 and torch.allclose(sums, sums2, rtol=2e-3, atol=0.05)
    dy = ext.bn_bwd_apply(g, y, stats, gamma, sums, 1, False, False, count, 1e-5)[0]
    dy2 = emu.bn_bwd_apply(g2, y, stats, gamma, sums2, 1, False, False, count, 1e-5)[0]
    assert _rel2(dy, dy2) < 5e-3


def _skip_bf16_whole_trunk():
    if not _is_tf32():
        pytest.skip("whole-trunk comparisons at random init are rounding-noise dominated with bf16 operands "
                    "(measured in round 1); the kernels are covered layer by layer above and whole-trunk in tf32")


@pytest.mark.parametrize("library_conv", [False, True])
def test_encoder_engine_trunk_matches_specification(library_conv):
    """Whole ResNet-18 trunk: kernels vs the same orchestration run through the PyTorch specification (identical
    rounding points, so only summation order differs), forward and every parameter gradient.  ResNet-50 at random
    init is too badly conditioned in bf16 for a gradient comparison (even in fp32 it needs 5e-2, see
    tests/test_engine_emulated.py); its forward is checked against the library encoder below."""
    from mine_b200.models.encoder import ResnetEncoder
    from mine_b200.ops import conv_engine as E
    from mine_b200.ops.encoder_engine import EncoderEngine
    _skip_bf16_whole_trunk()
    torch.manual_seed(0)
    enc = ResnetEncoder(18, False).cuda()
    img = torch.rand(2, 3, 256, 384, device="cuda")
    state = {k: v.clone() for k, v in enc.state_dict().items()}

    def run():
        enc.load_state_dict(state)
        for p in enc.parameters():
            p.grad = None
        outs = EncoderEngine(enc, library_conv=library_conv)(img)
        torch.manual_seed(1)
        gouts = [torch.randn_like(o.float()) for o in outs]
        sum((o.float() * g).sum() for o, g in zip(outs, gouts)).backward()
        return [o.detach().float() for o in outs], {k: p.grad.clone() for k, p in enc.named_parameters()}
    outs, grads = run()
    E.use_emulator(True, E.ACT_DTYPE, tf32_operands=True)
    try:
        ref_outs, ref_grads = run()
    finally:
        E.use_emulator(False)
    for i, (o, r) in enumerate(zip(outs, ref_outs)):
        assert _rel2(o, r) < 2e-2, (i, _rel2(o, r))
    bad = [(k, round(_rel2(grads[k], ref_grads[k]), 3)) for k in grads if _rel2(grads[k], ref_grads[k]) > 0.1]
    assert len(bad) <= 2, bad[:10]


def test_encoder_engine_resnet50_forward_matches_library_encoder():
    """Whole ResNet-50 trunk (training-mode BatchNorm, random init) against the library encoder in true fp32.  Such a
    network amplifies rounding noise with depth, so the bound is relative to how far the LIBRARY's own TF32 path (the
    reference's numerics: cuDNN TF32 convolutions on fp32 tensors) is from true fp32 on the same input."""
    from mine_b200.models.encoder import ResnetEncoder
    from mine_b200.ops.encoder_engine import EncoderEngine
    _skip_bf16_whole_trunk()
    torch.manual_seed(0)
    enc = ResnetEncoder(50, False).cuda()
    img = torch.rand(2, 3, 256, 384, device="cuda")
    state = {k: v.clone() for k, v in enc.state_dict().items()}
    with torch.no_grad():
        outs = EncoderEngine(enc)(img)
        enc.load_state_dict(state)
        refs = enc(img.contiguous(memory_format=torch.channels_last))                    # true fp32 (fixture)
        enc.load_state_dict(state)
        torch.backends.cudnn.allow_tf32 = True
        lib_tf32 = enc(img.contiguous(memory_format=torch.channels_last))
        torch.backends.cudnn.allow_tf32 = False
    report = []
    for i, (o, r, l) in enumerate(zip(outs, refs, lib_tf32)):
        ours, lib = _rel2(o, r), _rel2(l, r)
        report.append((i, round(ours, 5), round(lib, 5)))
        assert o.shape == r.shape and ours < max(3.0 * lib, 5e-3), report
    print("resnet50 trunk vs fp32: (level, ours tf32, library tf32)", report)


def test_stem_as_single_tap_gemm():
    from mine_b200.ops.encoder_engine import StemConv
    x = _bf(_rand((2, 3, 128, 192), 0)).requires_grad_(False)
    wt = _bf(_rand((64, 3, 7, 7), 1, 0.1)).requires_grad_()
    ref = F.conv2d(x, wt, None, 2, 3)
    y, stats = StemConv.apply(x, wt, True)
    assert _rel2(_nchw(y), ref) < _tol(6e-3, 1e-4)
    assert _rel2(stats[1], (ref * ref).sum(dim=(0, 2, 3))) < 5e-3
    dy = _bf(_rand(ref.shape, 2))
    ref.backward(dy)
    want = wt.grad.clone()
    wt.grad = None
    (y.float() * _nhwc(dy)).sum().backward()
    assert _rel2(wt.grad, want) < 1e-2


def test_library_free_prediction_matches_specification():
    """``ConvEngine(encoder_mode='tcgen05')``: stem, trunk, receptive-field extension, shared skip maps and the
    per-plane decoder all on the engine - kernels vs the same orchestration through the specification."""
    from mine_b200.models.decoder import DepthDecoder
    from mine_b200.models.encoder import ResnetEncoder
    from mine_b200.ops import conv_engine as E
    _skip_bf16_whole_trunk()
    torch.manual_seed(0)
    enc, dec = ResnetEncoder(18, False).cuda(), DepthDecoder(num_ch_enc=[64, 64, 128, 256, 512]).cuda()
    img = torch.rand(2, 3, 256, 256, device="cuda")
    disp = torch.rand(2, 4, device="cuda") * 0.8 + 0.1
    state = ({k: v.clone() for k, v in enc.state_dict().items()}, {k: v.clone() for k, v in dec.state_dict().items()})

    def run():
        enc.load_state_dict(state[0]); dec.load_state_dict(state[1])
        for p in list(enc.parameters()) + list(dec.parameters()):
            p.grad = None
        outs = E.ConvEngine(enc, dec, {}, torch.device("cuda"), encoder_mode="tcgen05").predict(img, disp)
        torch.manual_seed(1)
        gouts = [torch.randn_like(o) for o in outs]
        for g in gouts:
            g[..., 3] = 0
        sum((o * g).sum() for o, g in zip(outs, gouts)).backward()
        grads = {k: p.grad.clone() for k, p in list(dec.named_parameters()) + list(enc.named_parameters())
                 if p.grad is not None}
        return [o.detach() for o in outs], grads
    outs, grads = run()
    E.use_emulator(True, E.ACT_DTYPE, tf32_operands=True)
    try:
        ref_outs, ref_grads = run()
    finally:
        E.use_emulator(False)
    for i, (o, r) in enumerate(zip(outs, ref_outs)):
        assert _rel2(o, r) < 3e-2, (i, _rel2(o, r))
    bad = [(k, round(_rel2(grads[k], ref_grads[k]), 3)) for k in grads
           if not k.endswith("conv.conv.bias") and _rel2(grads[k], ref_grads[k]) > 0.25]
    assert len(bad) <= 3, bad[:10]


@pytest.mark.parametrize("k,stride,n,h,w,ci,co", [(1, 1, 2, 8, 12, 2048, 512), (3, 1, 2, 8, 12, 512, 512),
                                                  (3, 2, 2, 16, 24, 256, 256), (1, 1, 2, 16, 24, 1024, 256)])
def test_split_k_convolution(k, stride, n, h, w, ci, co, monkeypatch):
    """``conv_splitk.cu`` (MINE_B200_SPLITK=1): fprop + BatchNorm sums and the stride-1 data gradient."""
    from mine_b200.ops import encoder_engine as EE
    monkeypatch.setenv("MINE_B200_SPLITK", "1")
    x = _bf(_rand((n, ci, h, w), 0)).requires_grad_()
    wt = _bf(_rand((co, ci, k, k), 1, (ci * k * k) ** -0.5)).requires_grad_()
    ref = F.conv2d(x, wt, None, stride, k // 2)
    stats = torch.zeros(2, co, device="cuda")
    y = EE.conv_fprop(_nhwc(x.detach()).to(_ACT()), wt.detach(), stride, stats)
    assert _rel2(_nchw(y), ref) < _tol(6e-3, 1e-4)
    assert _rel2(stats[1], (ref * ref).sum(dim=(0, 2, 3))) < 5e-3
    if stride == 1:
        dy = _bf(_rand(ref.shape, 2))
        ref.backward(dy)
        got = EE.conv_dgrad(_nhwc(dy).to(_ACT()), wt.detach(), 1, h, w)
        assert _rel2(_nchw(got), x.grad) < _tol(6e-3, 1e-4)


def test_deterministic_reductions_are_bitwise_reproducible_and_match_the_atomic_path():
    """``engine.deterministic``: channel_stats / bn_res_act_bwd_reduce with the two-level fixed-order sums give bitwise
    identical results on repeated calls (the fp32-atomic default does not have to) and agree with the default path."""
    from mine_b200.ops import conv_engine as E
    ext = E.ext()
    if E.ACT_DTYPE != torch.float32:
        pytest.skip("one precision is enough")
    for (n, h, w, c) in [(2, 64, 96, 256), (2, 8, 12, 2048), (2, 128, 192, 64), (3, 5, 7, 512)]:
        gen = torch.Generator(device="cuda").manual_seed(c)
        y = torch.randn((n, h, w, c), device="cuda", generator=gen) * 1.7 + 0.3
        d = torch.randn((n, h, w, c), device="cuda", generator=gen)
        a = torch.relu(y)
        gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        count = float(n * h * w)
        try:
            E.set_deterministic(False)
            ref_stats = ext.channel_stats(y)
            ref_g, ref_sums = ext.bn_res_act_bwd_reduce(d, a, y, ref_stats, gamma, beta, 0.0, count, 1e-5)
            E.set_deterministic(True)
            runs = []
            for _ in range(3):
                st = ext.channel_stats(y)
                g, sums = ext.bn_res_act_bwd_reduce(d, a, y, st, gamma, beta, 0.0, count, 1e-5)
                runs.append((st.clone(), sums.clone(), g.clone()))
        finally:
            E.set_deterministic(False)
        for st, sums, g in runs[1:]:
            assert torch.equal(st, runs[0][0]) and torch.equal(sums, runs[0][1]) and torch.equal(g, runs[0][2])
        want = torch.stack([y.double().sum((0, 1, 2)), (y.double() ** 2).sum((0, 1, 2))]).float()
        assert _rel2(runs[0][0], want) < 1e-5 and _rel2(ref_stats, want) < 1e-5
        assert _rel2(runs[0][1], ref_sums) < 1e-4
