"""Engine orchestration on the CPU tier: ``conv_engine.py`` driven through the PyTorch specification of the kernels
(``mine_b200/ops/emu.py``) in fp32, compared with plain PyTorch modules / functional ops.

What this pins down without a GPU: tap tables, sub-pixel phase packing and its adjoint, strided dgrad gathers, the
padded-activation bookkeeping, BatchNorm forward/backward algebra, shared-skip / plane-bias factorisation, head
activation gradients and the autograd wiring of the whole decoder.  The kernels themselves are held to the same
specification by ``tests/test_conv_engine_gpu.py``."""
import pytest
import torch
import torch.nn.functional as F

from mine_b200.ops import conv_engine as E


@pytest.fixture(autouse=True)
def _emulated_fp32():
    E.use_emulator(True, torch.float32)
    yield
    E.use_emulator(False)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


SHAPES = [(2, 7, 10, 16, 32), (1, 12, 9, 32, 16), (2, 6, 6, 64, 64)]


@pytest.mark.parametrize("n,h,w,ci,co", SHAPES)
def test_same_conv_all_directions(n, h, w, ci, co):
    x = _rand((n, ci, h, w), 0).requires_grad_()
    wt = _rand((co, ci, 3, 3), 1, 0.1).requires_grad_()
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    xp.retain_grad()
    ref = F.conv2d(xp, wt)
    stats = torch.zeros(2, co)
    pb, sm = _rand((n, co), 2), _rand((1, h, w, co), 3)
    y = E.conv_same_raw(_nhwc(xp.detach()), wt.detach(), plane_bias=pb, shared_map=sm, planes=n, stats=stats)
    full = ref + pb[:, :, None, None] + _nchw(sm)
    assert torch.allclose(_nchw(y), full, atol=1e-4)
    assert torch.allclose(stats[0], full.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[1], (full * full).sum(dim=(0, 2, 3)), rtol=1e-4)
    dy = _rand(ref.shape, 4)
    ref.backward(dy)
    assert torch.allclose(_nchw(E.dgrad_same_raw(_nhwc(dy), wt.detach())), xp.grad, atol=1e-4)
    assert torch.allclose(E.wgrad_same_raw(_nhwc(dy), _nhwc(xp.detach())), wt.grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("n,h,w,ci,co", SHAPES)
def test_upsample_conv_all_directions(n, h, w, ci, co):
    """conv3x3(reflect_pad(nearest_up2(x))) == 4 sub-pixel 2x2 convs on the replicate-padded low-res input."""
    x = _rand((n, ci, h, w), 0)
    wt = _rand((co, ci, 3, 3), 1, 0.1).requires_grad_()
    xlo = F.pad(x, (1, 1, 1, 1), mode="replicate").requires_grad_()
    inner = xlo[:, :, 1:-1, 1:-1]
    ref = F.conv2d(F.pad(F.interpolate(inner, scale_factor=2, mode="nearest"), (1, 1, 1, 1), mode="reflect"), wt)
    y = E.conv_up_raw(_nhwc(xlo.detach()), wt.detach())
    assert torch.allclose(_nchw(y), ref, atol=1e-4)
    # gradients w.r.t. the padded low-res tensor: the engine differentiates the phase form, in which the border
    # samples come from the pad ring, so compare after folding the ring back (adjoint of replicate padding)
    dy = _rand(ref.shape, 4)
    ref.backward(dy)

    def fold(g):                     # NCHW gradient of the padded tensor -> gradient of the unpadded tensor
        g = g.clone()
        g[:, :, 1] += g[:, :, 0]; g[:, :, -2] += g[:, :, -1]
        g[:, :, :, 1] += g[:, :, :, 0]; g[:, :, :, -2] += g[:, :, :, -1]
        return g[:, :, 1:-1, 1:-1]
    got = _nchw(E.dgrad_up_raw(_nhwc(dy), wt.detach()))
    assert torch.allclose(fold(got), fold(xlo.grad), atol=1e-4)
    assert torch.allclose(E.wgrad_up_raw(_nhwc(dy), _nhwc(xlo.detach())), wt.grad, rtol=1e-4, atol=1e-3)
    # CUDA packer definition == einsum definition
    assert torch.allclose(E.pack(wt, 1)[:, :co].reshape(4, 4, co, ci), E.pack_up(wt.detach()), atol=1e-6)


@pytest.mark.parametrize("head", ["tcgen05", "direct"])
def test_fused_layer_autograd_matches_module_chain(head, monkeypatch):
    monkeypatch.setenv("MINE_B200_HEAD", head)
    _fused_layer_check()


def _fused_layer_check():
    """PlaneConvBNAct (conv + shared skip + plane bias -> BN -> ELU -> pad) and HeadConv against torch autograd."""
    b, s, h, w, ci, co = 2, 3, 6, 8, 32, 16
    n = b * s
    x = _rand((n, ci, h, w), 0)
    wt = _rand((co, ci, 3, 3), 1, 0.1).requires_grad_()
    sm = _rand((b, h, w, co), 2).requires_grad_()
    pb = _rand((n, co), 3).requires_grad_()
    gamma, beta = (torch.rand(co, generator=torch.Generator().manual_seed(4)) + 0.5).requires_grad_(), _rand((co,), 5).requires_grad_()
    xin = x.clone().requires_grad_()
    xp = F.pad(xin, (1, 1, 1, 1), mode="reflect")
    yref = F.conv2d(xp, wt) + _nchw(sm).repeat_interleave(s, 0) + pb[:, :, None, None]
    aref = F.pad(F.elu(F.batch_norm(yref, None, None, gamma, beta, True, 0.1, E.BN_EPS)), (1, 1, 1, 1), mode="reflect")
    hw = _rand((4, co, 3, 3), 6, 0.2).requires_grad_()
    hb = _rand((4,), 7).requires_grad_()
    z = F.conv2d(aref, hw, hb)
    mref = torch.cat([torch.sigmoid(z[:, :3]), z[:, 3:].abs() + 1e-4], 1)
    gm = _rand(mref.shape, 8)
    (mref * gm).sum().backward()
    want = [t.grad.clone() for t in (xin, wt, sm, pb, gamma, beta, hw, hb)]
    for t in (xin, wt, sm, pb, gamma, beta, hw, hb):
        t.grad = None

    xp2 = F.pad(xin, (1, 1, 1, 1), mode="reflect")
    apad = E.PlaneConvBNAct.apply(_nhwc(xp2), wt, None, pb, sm, gamma, beta, False, s, 0, None, None)
    mpi = E.HeadConv.apply(apad, hw, hb, False)
    assert torch.allclose(_nchw(mpi), mref, atol=1e-4)
    (mpi * _nhwc(gm)).sum().backward()
    got = [t.grad for t in (xin, wt, sm, pb, gamma, beta, hw, hb)]
    for name, g, r in zip("x w shared plane_bias gamma beta head_w head_b".split(), got, want):
        assert torch.allclose(g, r, rtol=2e-3, atol=2e-4 * r.abs().max().item() + 1e-5), name


@pytest.mark.parametrize("encoder_mode", ["cudnn", "tcgen05", "hybrid"])
def test_whole_decoder_on_emulated_engine_matches_modules(encoder_mode):
    from mine_b200.models.decoder import DepthDecoder
    from mine_b200.models.encoder import ResnetEncoder
    torch.manual_seed(0)
    enc, dec = ResnetEncoder(18, False), DepthDecoder(num_ch_enc=[64, 64, 128, 256, 512])
    b, s, h, w = 1, 2, 64, 64
    img = torch.rand(b, 3, h, w)
    disp = torch.rand(b, s) * 0.8 + 0.1
    outs = E.ConvEngine(enc, dec, {}, torch.device("cpu"), encoder_mode=encoder_mode).predict(img, disp)
    gouts = [torch.randn_like(o) for o in outs]
    sum((o * g).sum() for o, g in zip(outs, gouts)).backward()
    got = {k: p.grad.clone() for k, p in list(dec.named_parameters()) + [("enc." + k, p) for k, p in enc.named_parameters()]
           if p.grad is not None}
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    ref = dec(enc(img), disp)
    refs = [ref[("disp", k)].permute(0, 1, 3, 4, 2) for k in range(4)]
    for k in range(4):
        assert torch.allclose(outs[k], refs[k], atol=2e-4), k
    sum((o * g).sum() for o, g in zip(refs, gouts)).backward()
    checked = 0
    for kname, p in list(dec.named_parameters()) + [("enc." + k, p) for k, p in enc.named_parameters()]:
        if p.grad is None or kname.endswith("conv.conv.bias"):
            continue        # conv biases in front of BatchNorm have an exactly-zero true gradient
        assert kname in got, kname
        # absolute floor: at this tiny resolution the receptive-field-extension maps are 1x1, so two of its
        # BatchNorm shifts have an exactly-zero true gradient and both sides return ~1e-4 of rounding noise
        diff = (got[kname] - p.grad).norm().item()
        assert diff < 5e-3 * p.grad.norm().item() + 1e-3, (kname, diff, p.grad.norm().item())
        checked += 1
    assert checked > 80


# ---- encoder on the engine ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,stride,h,w", [(1, 1, 6, 8), (3, 1, 7, 9), (1, 2, 8, 12), (3, 2, 8, 12), (1, 2, 7, 9)])
def test_encoder_conv_directions(k, stride, h, w):
    from mine_b200.ops import encoder_engine as EE
    if k == 3 and stride == 2 and (h % 2 or w % 2):
        pytest.skip("even sizes only")
    n, ci, co = 2, 32, 64
    x = _rand((n, ci, h, w), 0).requires_grad_()
    wt = _rand((co, ci, k, k), 1, 0.1).requires_grad_()
    ref = F.conv2d(x, wt, None, stride, k // 2)
    stats = torch.zeros(2, co)
    y = EE.conv_fprop(_nhwc(x.detach()), wt.detach(), stride, stats)
    assert torch.allclose(_nchw(y), ref, atol=1e-4)
    assert torch.allclose(stats[0], ref.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)
    dy = _rand(ref.shape, 2)
    ref.backward(dy)
    assert torch.allclose(_nchw(EE.conv_dgrad(_nhwc(dy), wt.detach(), stride, h, w)), x.grad, atol=1e-4)
    assert torch.allclose(EE.conv_wgrad(_nhwc(dy), _nhwc(x.detach()), k, stride), wt.grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("layers,train,lib", [(18, True, False), (50, True, False), (50, False, False), (18, True, True),
                                              (50, True, True)])
def test_encoder_engine_matches_module(layers, train, lib):
    from mine_b200.models.encoder import ResnetEncoder
    from mine_b200.ops.encoder_engine import EncoderEngine
    torch.manual_seed(0)
    enc = ResnetEncoder(layers, False)
    enc.train(train)
    # the deep variant is badly conditioned at random init (few samples per BatchNorm in layer4, gradient norms
    # ~1e6): fp32 summation-order noise is amplified, so it gets more pixels and a looser bound than ResNet-18
    img = torch.rand(2, 3, 64, 96) if layers == 18 else torch.rand(2, 3, 128, 128)
    tol = 5e-3 if layers == 18 or not train else 5e-2
    state = {k: v.clone() for k, v in enc.state_dict().items()}
    outs = EncoderEngine(enc, library_conv=lib)(img)
    gouts = [torch.randn_like(o) for o in outs]
    sum((o * g).sum() for o, g in zip(outs, gouts)).backward()
    got = {k: p.grad.clone() for k, p in enc.named_parameters()}
    run_stats = {k: v.clone() for k, v in enc.state_dict().items() if "running" in k or "tracked" in k}
    enc.load_state_dict(state)
    for p in enc.parameters():
        p.grad = None
    refs = enc(img)
    for o, r in zip(outs, refs):
        assert o.shape == r.shape and torch.allclose(o, r, atol=2e-4 * max(1.0, r.abs().max().item()))
    sum((o * g).sum() for o, g in zip(refs, gouts)).backward()
    for k, p in enc.named_parameters():
        diff = (got[k] - p.grad).norm().item()
        assert diff < tol * p.grad.norm().item() + 1e-3, (k, diff, p.grad.norm().item())
    for k, v in enc.state_dict().items():
        if k in run_stats:
            assert torch.allclose(run_stats[k].float(), v.float(), rtol=1e-4, atol=1e-5), k


def test_training_step_through_emulated_engine_matches_module_path(monkeypatch):
    """The whole task step (prediction, rendering, losses, backward, Adam) with the engine orchestration in the loop
    reproduces the plain-module path: same losses, same updated parameters."""
    from mine_b200 import config as C
    from mine_b200.data.synthetic import config_batch
    from mine_b200.task import SynthesisTask
    shape = {"data.img_w": 64, "data.img_h": 64, "mpi.num_bins_coarse": 3, "data.per_gpu_batch_size": 1,
             "data.visible_point_count": 16, "model.imagenet_pretrained": False}

    def run(mode):
        monkeypatch.setenv("MINE_B200_CONV", mode)
        cfg = C.config_for_dataset("llff", shape)
        cfg["device"] = torch.device("cpu")
        torch.manual_seed(0)
        task = SynthesisTask(cfg, None)
        assert task.runner.mode == mode
        torch.manual_seed(1)
        out = task.train_step(config_batch(cfg))
        losses = {k: float(v.detach()) for k, v in out.items() if torch.is_tensor(v) and v.numel() == 1}
        return losses, torch.cat([p.detach().reshape(-1) for p in task.decoder.parameters()])
    la, pa = run("tcgen05")
    lb, pb = run("spec")
    for k in lb:
        assert la[k] == pytest.approx(lb[k], rel=2e-3, abs=1e-4), k
    assert torch.allclose(pa, pb, atol=2.5e-3)          # one Adam step moves every weight by ~lr = 1e-3


def test_sparse_point_node_matches_spec_composition():
    """One autograd node for projection + gather + scale calibration + log-L1 (``ops/sparse.py`` on the kernel
    specification) against the composition of specification ops, including the gradient that reaches the source
    disparity map THROUGH the scale factor used by the other view and the other pyramid levels."""
    from mine_b200.ops import api
    from mine_b200.ops.sparse import sparse_point_loss as fused
    g = torch.Generator().manual_seed(0)
    b, h, w, n = 3, 24, 40, 50
    k = torch.tensor([[30.0, 0, 20], [0, 30.0, 12], [0, 0, 1]]).expand(b, 3, 3).contiguous()
    z = torch.rand(b, 1, n, generator=g) * 4 + 1
    xy = (torch.rand(b, 2, n, generator=g) - 0.5) * torch.tensor([1.6, 1.0]).view(1, 2, 1)      # a few fall outside
    xyz = torch.cat([xy * z, z], dim=1)

    def run(fn):
        ds = (torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(1)) + 0.2).requires_grad_()
        dt = (torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(2)) + 0.2).requires_grad_()
        l_src, scale = fn(ds, k, xyz, None)
        l_tgt, _ = fn(dt, k, xyz * 1.1, scale)
        l_lo, _ = fn(ds[..., ::2, ::2] * 1.0, k * torch.tensor([0.5, 0.5, 1.0]).view(1, 3, 1), xyz, scale)
        total = l_src + 2.0 * l_tgt + 0.5 * l_lo + 0.1 * scale.sum()
        total.backward()
        return [float(v.detach()) for v in (l_src, l_tgt, l_lo)], scale.detach(), ds.grad, dt.grad
    want = run(lambda d, kk, p, s: api.sparse_point_loss(d, kk, p, s))
    got = run(fused)
    assert got[0] == pytest.approx(want[0], rel=1e-5)
    assert torch.allclose(got[1], want[1], rtol=1e-5)
    assert torch.allclose(got[2], want[2], rtol=1e-4, atol=1e-7) and torch.allclose(got[3], want[3], rtol=1e-4, atol=1e-7)
    assert (want[2] != 0).sum() > 10


def test_fused_running_stat_update_matches_framework_ops(monkeypatch):
    from mine_b200.models.norm import BatchNorm
    stats = torch.stack([_rand((16,), 0) * 50, torch.rand(16, generator=torch.Generator().manual_seed(1)) * 500 + 300])
    a, b = BatchNorm(16), BatchNorm(16)
    monkeypatch.setenv("MINE_B200_BN_RUNNING", "aten")
    E.update_running_stats(a, stats, 96.0)
    monkeypatch.setenv("MINE_B200_BN_RUNNING", "fused")
    E.update_running_stats(b, stats, 96.0)
    assert torch.allclose(a.running_mean, b.running_mean) and torch.allclose(a.running_var, b.running_var)
    assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == 1


def test_split_k_path_of_encoder_convs(monkeypatch):
    """MINE_B200_SPLITK=1 routes convolutions with few output tiles through the split-K entry points (fp32 partial
    sums + finalize); values and BatchNorm sums are those of the single-pass form."""
    from mine_b200.ops import encoder_engine as EE
    x = _nhwc(_rand((2, 512, 8, 12), 0))
    wt = _rand((256, 512, 3, 3), 1, 0.02)
    monkeypatch.setenv("MINE_B200_SPLITK", "0")
    s0 = torch.zeros(2, 256)
    y0 = EE.conv_fprop(x, wt, 1, s0)
    monkeypatch.setenv("MINE_B200_SPLITK", "1")
    assert EE.split_factor(2, 72) == 16 and EE.split_factor(400, 72) == 1 and EE.split_factor(8, 8) == 2
    s1 = torch.zeros(2, 256)
    y1 = EE.conv_fprop(x, wt, 1, s1)
    assert torch.allclose(y0, y1, atol=1e-4) and torch.allclose(s0, s1, rtol=1e-4, atol=1e-3)
    y2 = EE.conv_fprop(x, wt, 2, None)
    monkeypatch.setenv("MINE_B200_SPLITK", "0")
    assert torch.allclose(y2, EE.conv_fprop(x, wt, 2, None), atol=1e-4)


def test_split_weight_gradient_equals_plain_slicing():
    """``SplitWeight`` (one gradient buffer in the parameter's layout for the three input-channel groups of a factorised
    decoder conv) == autograd through plain slices, incl. a group that receives no gradient."""
    import torch
    from mine_b200.ops import conv_engine as E
    torch.manual_seed(0)
    w0 = torch.randn(8, 11, 3, 3)
    for use in ((True, True, True), (True, False, True), (False, True, False)):
        grads = []
        for split in (lambda w: E.SplitWeight.apply(w, 4, 5), lambda w: (w[:, :4], w[:, 4:9], w[:, 9:])):
            w = w0.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)
            parts = split(w)
            loss = sum((p * (i + 1.5)).square().sum() for i, (p, u) in enumerate(zip(parts, use)) if u)
            loss.backward()
            grads.append(w.grad.clone())
        assert torch.allclose(grads[0], grads[1], rtol=1e-6, atol=1e-6)
        assert grads[0].stride() == w0.contiguous(memory_format=torch.channels_last).stride()
