"""CUDA-core MPI head (``csrc/head_direct.cu``) vs the tcgen05 head and the PyTorch specification.
Opt-in (``MINE_B200_TEST_OPTIN=1``) until the kernel has been run on hardware."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MINE_B200_TEST_OPTIN", "0") != "1",
                                 reason="opt-in kernels (MINE_B200_TEST_OPTIN=1)")]


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
