"""Golden-oracle tests: our PyTorch spec == upstream ops on CPU."""
import math

import numpy as np
import pytest
import torch

from mine_b200 import geometry as geo
from mine_b200.bench.ref_shims import cpu_cuda_sync_noop
from mine_b200.spec import embedder, losses, render, sampling


def _cams(b, h, w, seed=0, rot=0.05, trans=0.1):
    g = torch.Generator().manual_seed(seed)
    f = 0.8 * w
    k = torch.tensor([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1.0]]).repeat(b, 1, 1)
    k[:, 0, 0] += torch.rand(b, generator=g) * 10
    ang = (torch.rand(b, 3, generator=g) - 0.5) * 2 * rot
    rs = []
    for a in ang:
        cx, sx, cy, sy, cz, sz = math.cos(a[0]), math.sin(a[0]), math.cos(a[1]), math.sin(a[1]), math.cos(a[2]), math.sin(a[2])
        rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
        rs.append(rz @ ry @ rx)
    gm = torch.eye(4).repeat(b, 1, 1)
    gm[:, :3, :3] = torch.stack(rs)
    gm[:, :3, 3] = (torch.rand(b, 3, generator=g) - 0.5) * 2 * trans
    return k, gm


def _mpi(b, s, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(b, s, 3, h, w, generator=g)
    sigma = torch.rand(b, s, 1, h, w, generator=g) * 2 + 1e-4
    disp = sampling.stratified_disparity_linspace(b, s, 1.0, 0.01, generator=g)
    return rgb, sigma, disp


def test_inverses():
    k, g = _cams(4, 64, 96)
    assert torch.allclose(geo.inv3x3(k), torch.inverse(k), atol=1e-6)
    assert torch.allclose(geo.inv_rigid(g), torch.inverse(g), atol=1e-6)
    assert torch.allclose(geo.inv_affine4x4(g), torch.inverse(g), atol=1e-6)


def test_embedder_matches_reference(ref):
    ru = ref.load("utils")
    fn, dim = ru.get_embedder(10)
    x = torch.rand(7, 1)
    assert dim == embedder.embedding_dim(10) == 21
    assert torch.allclose(fn(x), embedder.positional_encoding(x, 10), atol=1e-6)


@pytest.mark.parametrize("use_alpha,bg_inf", [(False, False), (False, True), (True, False)])
def test_src_render_matches_reference(ref, use_alpha, bg_inf):
    mr = ref.load("operations.mpi_rendering")
    hs = ref.load("operations.homography_sampler")
    b, s, h, w = 2, 6, 32, 48
    rgb, sigma, disp = _mpi(b, s, h, w)
    if use_alpha:
        sigma = torch.rand_like(sigma)
    k, _ = _cams(b, h, w)
    kinv = geo.inv3x3(k)
    mesh = hs.HomographySample(h, w).meshgrid
    xyz_ref = mr.get_src_xyz_from_plane_disparity(mesh, disp, kinv)
    xyz = render.src_plane_points(kinv, disp, h, w)
    assert torch.allclose(xyz, xyz_ref, rtol=1e-5, atol=1e-5)
    if use_alpha:
        import torch as _t
        real_cuda = _t.Tensor.cuda
        _t.Tensor.cuda = lambda self, *a, **k: self
        try:
            out_ref = mr.render(rgb, sigma, xyz_ref, use_alpha=True, is_bg_depth_inf=bg_inf)
        finally:
            _t.Tensor.cuda = real_cuda
    else:
        out_ref = mr.render(rgb, sigma, xyz_ref, use_alpha=False, is_bg_depth_inf=bg_inf)
    out = render.render(rgb, sigma, xyz, use_alpha, bg_inf)
    for a, r in zip(out, out_ref):
        assert torch.allclose(a, r, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("rot,trans", [(0.0, 0.0), (0.05, 0.1), (0.3, 0.5)])
def test_tgt_render_matches_reference(ref, rot, trans):
    mr = ref.load("operations.mpi_rendering")
    hs = ref.load("operations.homography_sampler")
    b, s, h, w = 2, 5, 32, 48
    rgb, sigma, disp = _mpi(b, s, h, w, seed=3)
    k, g = _cams(b, h, w, seed=5, rot=rot, trans=trans)
    kinv = geo.inv3x3(k)
    sampler = hs.HomographySample(h, w)
    xyz_src = mr.get_src_xyz_from_plane_disparity(sampler.meshgrid, disp, kinv)
    xyz_tgt = mr.get_tgt_xyz_from_plane_disparity(xyz_src, g)
    with cpu_cuda_sync_noop():
        r_rgb, r_depth, r_mask = mr.render_tgt_rgb_depth(sampler, rgb, sigma, disp, xyz_tgt, g, kinv, k)
    o_rgb, o_depth, o_mask = render.render_tgt(rgb, sigma, disp, g, kinv, k)
    assert torch.equal(o_mask, r_mask)
    assert torch.allclose(o_rgb, r_rgb, rtol=1e-4, atol=2e-5)
    assert torch.allclose(o_depth, r_depth, rtol=2e-4, atol=1e-4)
    if rot == 0.0 and trans == 0.0:
        src = render.render_src(rgb, sigma, disp, kinv, blend=False)
        assert (o_rgb - src["rgb"]).abs().max() < 1e-5
        assert torch.all(o_mask == s)


def test_losses_match_reference(ref):
    rs = ref.load("network.ssim")
    rl = ref.load("network.layers")
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(2, 3, 40, 56, generator=g), torch.rand(2, 3, 40, 56, generator=g)
    d = torch.rand(2, 1, 40, 56, generator=g) + 0.1
    assert torch.allclose(losses.ssim(a, b), rs.SSIM()(a, b), atol=1e-6)
    assert torch.allclose(losses.psnr(a, b), rl.psnr(a, b), atol=1e-5)
    assert torch.allclose(losses.edge_aware_loss_v2(a, d), rl.edge_aware_loss_v2(a, d), atol=1e-6)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a_, **k_: self
    try:
        r = rl.edge_aware_loss(a, d, gmin=0.8, grad_ratio=0.2)
    finally:
        torch.Tensor.cuda = real_cuda
    assert torch.allclose(losses.edge_aware_loss(a, d, 0.8, 0.2), r, atol=1e-6)


def test_sampling_matches_reference(ref):
    ru = ref.load("operations.rendering_utils")
    torch.manual_seed(0)
    r = ru.uniformly_sample_disparity_from_linspace_bins(3, 8, 1.0, 0.001, device=torch.device("cpu"))
    torch.manual_seed(0)
    o = sampling.stratified_disparity_linspace(3, 8, 1.0, 0.001)
    assert torch.allclose(r, o)
    assert torch.all(o[:, :-1] > o[:, 1:])
    edges = np.linspace(2.0, 0.1, 9).astype(np.float32)
    torch.manual_seed(1)
    r = ru.uniformly_sample_disparity_from_bins(3, edges, torch.device("cpu"))
    torch.manual_seed(1)
    o = sampling.stratified_disparity(3, torch.from_numpy(edges))
    assert torch.allclose(r, o)
    img = torch.rand(2, 3, 10, 12)
    px = torch.rand(2, 2, 9) * 14 - 1
    assert torch.equal(ru.gather_pixel_by_pxpy(img, px), sampling.gather_nearest(img, px))
    vals = torch.sort(torch.rand(2, 1, 1, 8), dim=-1, descending=True).values
    wts = torch.rand(2, 1, 1, 8)
    torch.manual_seed(2)
    r = ru.sample_pdf(vals, wts, 5)
    torch.manual_seed(2)
    o = sampling.sample_pdf(vals, wts, 5)
    assert torch.allclose(r, o, atol=1e-6)
