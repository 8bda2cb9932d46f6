"""Property tests of the specification layer (hypothesis): invariants the kernels are tested against."""
import math

import pytest

import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from mine_b200 import geometry as geo
from mine_b200.spec import render as R
from mine_b200.spec import sampling as S
from mine_b200.spec.embedder import embedding_dim, positional_encoding


def _pose(rx, ry, rz, tx, ty, tz):
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    r = (torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]]) @ torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
         @ torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    g = torch.eye(4)
    g[:3, :3] = r
    g[:3, 3] = torch.tensor([tx, ty, tz])
    return g[None]


small = st.floats(-0.2, 0.2, allow_nan=False)


@settings(max_examples=25, deadline=None)
@given(small, small, small, small, small, small, st.floats(40.0, 200.0), st.integers(0, 10 ** 6))
def test_inverse_and_homography_consistency(rx, ry, rz, tx, ty, tz, f, seed):
    g = _pose(rx, ry, rz, tx, ty, tz)
    assert torch.allclose(geo.inv_affine4x4(g) @ g, torch.eye(4)[None], atol=1e-5)
    assert torch.allclose(geo.inv_rigid(g), geo.inv_affine4x4(g), atol=1e-5)
    k = torch.tensor([[f, 0, 32.0], [0, f * 1.1, 24.0], [0, 0, 1.0]])[None]
    kinv = geo.inv3x3(k)
    assert torch.allclose(kinv @ k, torch.eye(3)[None], atol=1e-5)
    # a point on the plane z = d in the source frame projects consistently through the plane homography
    d = torch.tensor([[2.5]])
    h = geo.plane_homography(k, kinv, g, d)[0, 0]
    gen = torch.Generator().manual_seed(seed)
    uv = torch.rand(2, generator=gen) * torch.tensor([64.0, 48.0])
    p_src = kinv[0] @ torch.tensor([uv[0], uv[1], 1.0]) * d[0, 0]
    p_tgt = g[0, :3, :3] @ p_src + g[0, :3, 3]
    want = k[0] @ p_tgt
    got = h @ torch.tensor([uv[0], uv[1], 1.0])
    assert torch.allclose(got / got[2], want / want[2], atol=1e-3)


@settings(max_examples=20, deadline=None)
@given(st.integers(2, 9), st.integers(0, 10 ** 6), st.booleans())
def test_compositing_weights_are_a_partition(s, seed, alpha_mode):
    gen = torch.Generator().manual_seed(seed)
    b, h, w = 2, 6, 7
    disp = S.stratified_disparity_linspace(b, s, 1.0, 0.05, generator=gen)
    assert torch.all(disp[:, :-1] > disp[:, 1:])
    kinv = geo.inv3x3(torch.tensor([[50.0, 0, 3.5], [0, 50.0, 3.0], [0, 0, 1.0]])[None].repeat(b, 1, 1))
    xyz = R.src_plane_points(kinv, disp, h, w)
    assert torch.allclose(xyz[:, :, 2], (1.0 / disp)[:, :, None, None].expand(-1, -1, h, w), rtol=1e-5)
    if alpha_mode:
        a = torch.rand(b, s, 1, h, w, generator=gen)
        wts = R.alpha_to_weights(a)
    else:
        sig = 0.05 + torch.rand(b, s, 1, h, w, generator=gen) * 3      # bounded away from 0: the last plane is opaque
        t_acc, wts = R.sigma_to_weights(sig, xyz)
        assert torch.all(t_acc[:, 0] == 1) and torch.all(t_acc[:, 1:] <= t_acc[:, :-1] * (1 + 1e-5) + 1e-5)
        # the last plane (thickness 1e3) is opaque for sigma >= 0.05 (exp(-50)): weights sum to ~1
        assert torch.all((wts.sum(1) - 1).abs() < 1e-3 * s + 1e-4)
    assert torch.all(wts >= 0) and torch.all(wts.sum(1) <= 1 + 1e-4)
    rgb = torch.rand(b, s, 3, h, w, generator=gen)
    out, depth = R.composite(rgb, xyz, wts)
    assert torch.all(out >= -1e-6) and torch.all(out <= 1 + 1e-3)
    if not alpha_mode:
        zmin, zmax = (1.0 / disp).min(1).values, (1.0 / disp).max(1).values
        assert torch.all(depth >= zmin[:, None, None, None] * (1 - 1e-3)) and torch.all(depth <= zmax[:, None, None, None] * (1 + 1e-3))


@settings(max_examples=15, deadline=None)
@given(st.integers(1, 12), st.integers(0, 10 ** 6))
def test_positional_encoding_structure(multires, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(5, 1, generator=gen)
    e = positional_encoding(x, multires)
    assert e.shape[-1] == embedding_dim(multires) == 1 + 2 * multires
    assert torch.equal(e[:, :1], x)
    for k in range(multires):
        assert torch.allclose(e[:, 1 + 2 * k], torch.sin(x[:, 0] * 2.0 ** k), atol=1e-6)
        assert torch.allclose(e[:, 2 + 2 * k], torch.cos(x[:, 0] * 2.0 ** k), atol=1e-6)


@settings(max_examples=15, deadline=None)
@given(st.integers(3, 12), st.integers(1, 6), st.integers(0, 10 ** 6))
def test_importance_sampling_stays_in_range_and_sorted(s, n_fine, seed):
    gen = torch.Generator().manual_seed(seed)
    disp = S.stratified_disparity_linspace(2, s, 1.0, 0.01, generator=gen)
    w = torch.rand(2, s, 1, 4, 4, generator=gen)
    allp = S.refine_disparity(disp, w, n_fine, generator=gen)
    assert allp.shape == (2, s + n_fine)
    assert torch.all(allp[:, :-1] >= allp[:, 1:])
    assert torch.all(allp <= disp.max(1, keepdim=True).values + 1e-6) and torch.all(allp >= disp.min(1, keepdim=True).values - 1e-6)


def _native():
    """The in-tree extension loads without a GPU (host-side helpers only are called here)."""
    try:
        from mine_b200.ops import build as b
        return b.load()
    except Exception as e:                                   # not built in this environment
        pytest.skip("extension not available: %r" % (e,))


@settings(max_examples=300, deadline=None)
@given(st.integers(min_value=0, max_value=2 ** 31 - 1), st.integers(min_value=1, max_value=2 ** 20))
def test_fastdiv_constants_divide_exactly(n, d):
    """``make_fastdiv`` (multiply-high + shift, used by every persistent kernel to decompose work-item indices) equals
    integer division for every 31-bit numerator; divisors up to 2^20 cover tiles, images, planes and channel groups."""
    assert _native().fastdiv_host(n, d) == n // d


def test_fastdiv_edge_divisors():
    mod = _native()
    for d in (1, 2, 3, 7, 16, 17, 384, 386, 49152, 65535, 65536, 2 ** 20 - 1, 2 ** 30, 2 ** 31 - 1):
        for n in (0, 1, d - 1, d, d + 1, 2 * d - 1, 2 ** 31 - 1, 2 ** 31 - 2):
            if 0 <= n < 2 ** 31:
                assert mod.fastdiv_host(n, d) == n // d, (n, d)
