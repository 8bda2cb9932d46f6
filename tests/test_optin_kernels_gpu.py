"""Kernel variants behind environment switches, each against its alternative and the PyTorch specification: the
CUDA-core MPI head (``csrc/head_direct.cu``, bf16 operands) vs the tcgen05 head, the two BatchNorm backward
reductions (register-coefficient kernel = default, shared-memory-coefficient kernel = ``MINE_B200_BN_REDUCE=old``),
the fused sparse-point loss and the fused running-statistic update (both default on CUDA)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.fixture(autouse=True)
def _bf16_operands():
    """These variants are compared at bf16 operand precision (the direct head only exists there)."""
    from mine_b200.ops import conv_engine as E
    old = E.PRECISION
    E.set_precision("bf16")
    yield
    E.set_precision(old)


@pytest.mark.parametrize("c,n,h,w,alpha", [(16, 3, 40, 72, False), (32, 2, 33, 50, False), (16, 2, 16, 32, True)])
def test_head_direct_matches_tcgen05_head_and_spec(c, n, h, w, alpha, monkeypatch):
    from mine_b200.ops import conv_engine as E
    from mine_b200.ops import emu
    g = torch.Generator().manual_seed(0)
    apad = torch.randn((n, h + 2, w + 2, c), generator=g).cuda().to(torch.bfloat16)
    wt = (torch.randn((4, c, 3, 3), generator=g) * 0.2).cuda()
    bias = torch.randn(4, generator=g).cuda()
    monkeypatch.setenv("MINE_B200_HEAD", "direct")
    got = E.HeadConv.apply(apad, wt, bias, alpha)
    monkeypatch.setenv("MINE_B200_HEAD", "tcgen05")
    ref = E.HeadConv.apply(apad, wt, bias, alpha)           # bf16-rounded weights on the tensor cores
    spec, sign = emu.head_conv_direct(apad, wt.permute(2, 3, 1, 0).contiguous(), bias, alpha)
    assert got.shape == ref.shape == spec.shape
    assert (got - spec).abs().max().item() < 2e-3            # fp32 weights, fp32 accumulation: only summation order
    assert (got - ref).abs().max().item() < 3e-2
    mpi2, sign2 = E.ext().head_conv_direct(apad, wt.permute(2, 3, 1, 0).contiguous(), bias, alpha)
    far = spec[..., 3] > 1e-2 if not alpha else torch.ones_like(sign, dtype=torch.bool)
    assert torch.equal(sign2[far], sign[far])


_REDUCE_SNIPPET = r"""
import sys, torch
sys.path.insert(0, {repo!r})
from mine_b200.ops import conv_engine as E
E.set_precision("bf16")
g = torch.Generator().manual_seed(1)
out = {{}}
for c, pad in ((16, 0), (64, 1), (256, 0)):
    n, h, w = 6, 20, 36
    dapad = torch.randn((n, h + 2, w + 2, c), generator=g).cuda().to(torch.bfloat16)
    y = torch.randn((n, h, w, c), generator=g).cuda().to(torch.bfloat16)
    yf = y.float()
    stats = torch.stack([yf.sum((0, 1, 2)), (yf * yf).sum((0, 1, 2))])
    gamma, beta = torch.rand(c, generator=g).cuda() + 0.5, torch.randn(c, generator=g).cuda()
    gg, sums = E.ext().bn_act_bwd_reduce(dapad, y, stats, gamma, beta, pad, float(n * h * w), 1e-5)
    out[(c, pad)] = (gg.float().cpu(), sums.cpu())
torch.save(out, sys.argv[1])
"""


def test_bn_backward_reduce_variant2_matches_variant1(tmp_path):
    """``MINE_B200_BN_REDUCE=v2`` (register coefficients, sum(g*y) form) against the default kernel; the switch is
    read once per process, so each variant runs in its own interpreter."""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("v1", {"MINE_B200_BN_REDUCE": "old"}), ("v2", {"MINE_B200_BN_REDUCE": "v2"})):
        path = str(tmp_path / (tag + ".pt"))
        r = subprocess.run([sys.executable, "-c", _REDUCE_SNIPPET.format(repo=repo), path], env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[tag] = torch.load(path)
    for key in res["v1"]:
        g1, s1 = res["v1"][key]
        g2, s2 = res["v2"][key]
        assert torch.allclose(g1, g2, rtol=1e-2, atol=1e-3), key          # bf16 storage, 1-ulp differences in the affine
        assert torch.allclose(s1, s2, rtol=2e-3, atol=2e-2 * s1.abs().max().item()), key


def test_sparse_point_kernels_match_specification():
    """``csrc/sparse.cu`` (MINE_B200_SPARSE=fused) vs its PyTorch specification, forward values and the scattered
    disparity-map gradient, with the scale calibrated in the node and with a given scale."""
    from mine_b200.ops import conv_engine as E
    from mine_b200.ops.sparse import sparse_point_loss as fused
    g = torch.Generator().manual_seed(0)
    b, h, w, n = 2, 256, 384, 256
    k = torch.tensor([[300.0, 0, 192], [0, 300.0, 128], [0, 0, 1]]).expand(b, 3, 3).contiguous().cuda()
    z = torch.rand(b, 1, n, generator=g) * 4 + 1
    xy = (torch.rand(b, 2, n, generator=g) - 0.5) * torch.tensor([1.4, 0.9]).view(1, 2, 1)
    xyz = torch.cat([xy * z, z], dim=1).cuda()

    def run():
        ds = (torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(1)) + 0.2).cuda().requires_grad_()
        dt = (torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(2)) + 0.2).cuda().requires_grad_()
        l_src, scale = fused(ds, k, xyz, None)
        l_tgt, _ = fused(dt, k, xyz * 1.05, scale)
        (l_src + 2.0 * l_tgt + 0.1 * scale.sum()).backward()
        return l_src.detach(), l_tgt.detach(), scale.detach(), ds.grad, dt.grad
    got = run()
    E.use_emulator(True)
    try:
        want = run()
    finally:
        E.use_emulator(False)
    for a, r in zip(got[:3], want[:3]):
        assert torch.allclose(a, r, rtol=1e-4, atol=1e-6)
    for a, r in zip(got[3:], want[3:]):
        assert torch.allclose(a, r, rtol=1e-3, atol=1e-8)


def test_fused_running_stat_update_kernel(monkeypatch):
    from mine_b200.models.norm import BatchNorm
    from mine_b200.ops import conv_engine as E
    g = torch.Generator().manual_seed(0)
    stats = torch.stack([torch.randn(256, generator=g) * 50, torch.rand(256, generator=g) * 500 + 300]).cuda()
    a, b = BatchNorm(256).cuda(), BatchNorm(256).cuda()
    monkeypatch.setenv("MINE_B200_BN_RUNNING", "aten")
    E.update_running_stats(a, stats, 4096.0)
    monkeypatch.setenv("MINE_B200_BN_RUNNING", "fused")
    E.update_running_stats(b, stats, 4096.0)
    assert torch.allclose(a.running_mean, b.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(a.running_var, b.running_var, rtol=1e-5, atol=1e-6)
    assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == 1


def test_deferred_running_statistics_one_launch_for_many_layers():
    """``defer_running_stats``: the updates of a step are queued and applied by ``bn_update_running_multi`` (block = layer,
    more than one launch chunk) - same buffers as the immediate per-layer kernel."""
    from mine_b200.models.norm import BatchNorm
    from mine_b200.ops import conv_engine as E
    widths = [16, 32, 64, 256, 2048] * 11                      # 55 layers: two chunks of the 48-layer launch
    now = [BatchNorm(c).cuda() for c in widths]
    later = [BatchNorm(c).cuda() for c in widths]
    stats = []
    for i, c in enumerate(widths):
        gen = torch.Generator(device="cuda").manual_seed(100 + i)
        x = torch.randn((512, c), device="cuda", generator=gen) * (1 + 0.1 * i) + 0.05 * i
        stats.append(torch.stack([x.sum(0), (x * x).sum(0)]).contiguous())
    for bn, st in zip(now, stats):
        E.update_running_stats(bn, st, 512.0)
    E.defer_running_stats(True)
    for bn, st in zip(later, stats):
        E.update_running_stats(bn, st, 512.0)
    assert all(int(bn.num_batches_tracked) == 0 for bn in later)        # nothing applied yet
    E.defer_running_stats(False)                                         # flush
    for a, b in zip(now, later):
        assert torch.allclose(a.running_mean, b.running_mean, rtol=1e-6, atol=1e-7)
        assert torch.allclose(a.running_var, b.running_var, rtol=1e-6, atol=1e-7)
        assert int(b.num_batches_tracked) == 1


@pytest.mark.parametrize("shape", [(2, 64, 96, 256), (2, 8, 12, 2048), (3, 5, 7, 64), (1, 2, 3, 16), (2, 3, 3, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reflect_pad_nhwc_and_adjoint(shape, dtype):
    """``pad_nhwc.cu`` vs ``F.pad(mode="reflect")`` and its autograd, on NHWC storage (the skip features of the decoder)."""
    from mine_b200.ops import conv_engine as E
    n, h, w, c = shape
    gen = torch.Generator(device="cuda").manual_seed(h * w + c)
    x = torch.randn((n, h, w, c), device="cuda", generator=gen).to(dtype)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = torch.nn.functional.pad(xr, (1, 1, 1, 1), mode="reflect")
    xo = x.clone().requires_grad_(True)
    got = E.ReflectPadNHWC.apply(xo)
    assert got.shape == (n, h + 2, w + 2, c) and got.is_contiguous()
    assert torch.equal(got.float(), ref.detach().permute(0, 2, 3, 1).to(dtype).float())
    g = torch.randn(got.shape, device="cuda", generator=gen).to(dtype)
    got.backward(g)
    ref.backward(g.float().permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1)
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert torch.allclose(xo.grad.float(), want, rtol=tol, atol=tol)
