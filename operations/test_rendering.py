"""Runnable counterparts of the reference's manual rendering checks (``operations/test_rendering.py``).

The upstream file is three eyeball scripts around a hard-coded JPEG on the author's disk (alpha compositing of a
3-plane MPI, an Euler-angle convention print-out, and an unfinished homography call).  These versions build their
inputs procedurally, assert the property each script was meant to show, and optionally dump PNGs for inspection::

    python -m operations.test_rendering --out /tmp/render_demo
"""
import argparse
import os

import numpy as np
import torch

from mine_b200 import geometry as geo
from operations import mpi_rendering
from operations.homography_sampler import HomographySample


def _checker(h, w, cell=16):
    v, u = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    board = (((u // cell) + (v // cell)) % 2).float()
    return torch.stack([board, u.float() / w, v.float() / h], dim=0)[None]          # 1x3xHxW


def _save(img_b3hw, path):
    if path is None:
        return
    from PIL import Image
    os.makedirs(os.path.dirname(path), exist_ok=True)
    arr = (img_b3hw[0].clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype(np.uint8)
    Image.fromarray(arr).save(path)


def test_mpi_composition(out_dir=None, h=96, w=128):
    """Front-to-back alpha compositing: a half-transparent white plane, the picture with a hole, a darker copy."""
    img = _checker(h, w)
    a0 = torch.full((1, 1, h, w), 0.5)
    a1 = torch.ones(1, 1, h, w)
    a1[:, :, h // 4:3 * h // 4, w // 4:3 * w // 4] = 0
    a2 = torch.ones(1, 1, h, w)
    rgb = torch.stack([torch.ones_like(img), img, img * 0.5], dim=1)                 # 1x3x3xHxW
    alpha = torch.stack([a0, a1, a2], dim=1)
    out, weights = mpi_rendering.alpha_composition(alpha, rgb)
    assert torch.allclose(weights.sum(1), torch.ones(1, 1, h, w))                    # last plane is opaque
    inside = out[:, :, h // 2, w // 2]
    outside = out[:, :, 2, 2]
    assert torch.allclose(inside, 0.5 + 0.5 * 0.5 * img[:, :, h // 2, w // 2])       # sees the dark copy through the hole
    assert torch.allclose(outside, 0.5 + 0.5 * img[:, :, 2, 2])
    _save(out, None if out_dir is None else os.path.join(out_dir, "composition.png"))
    return out


def rotation_test():
    """Euler convention of the sampler helper: the matrix maps target points into the source frame, so it is the
    transpose of the forward rotation and composes in reverse order."""
    e = HomographySample.euler_to_rotation_matrix
    rx, ry, rz = e(30, 0, 0, degrees=True), e(0, 30, 0, degrees=True), e(0, 0, 30, degrees=True)
    rxyz = e(30, 30, 30, degrees=True)
    assert np.allclose(rxyz @ rxyz.T, np.eye(3), atol=1e-6)
    assert np.allclose(rxyz, rz @ ry @ rx, atol=1e-6)
    return rxyz


test_rotation = rotation_test


def K_from_img_HW(B, H, W, device=None):
    """Pinhole intrinsics with f = max(H, W): 53.13 degrees across the longer side."""
    f = float(max(H, W))
    k = torch.tensor([[f, 0, W * 0.5], [0, f, H * 0.5], [0, 0, 1]], dtype=torch.float32, device=device)
    return k[None].expand(B, 3, 3).contiguous()


def test_homography_sample(out_dir=None, h=96, w=128):
    """Warping a fronto-parallel plane: identity pose reproduces the image; a sideways camera shift of t moves the
    plane at depth d by f*t/d pixels; the validity mask marks target pixels that fall outside the source."""
    img = _checker(h, w)
    k = K_from_img_HW(1, h, w)
    kinv = geo.inv3x3(k)
    sampler = HomographySample(h, w)
    depth = torch.tensor([2.0])
    same, valid = sampler.sample(img, depth, torch.eye(4)[None], kinv, k)
    assert torch.allclose(same, img, atol=1e-5) and bool(valid.all())
    g = torch.eye(4)[None].clone()
    shift_px = 8
    g[0, 0, 3] = shift_px * depth[0] / k[0, 0, 0]                                    # target camera moved to -x
    moved, valid = sampler.sample(img, depth, g, kinv, k)
    assert torch.allclose(moved[..., shift_px:], img[..., :-shift_px], atol=1e-4)
    assert not bool(valid[0, :, :shift_px - 1].any()) and bool(valid[0, :, shift_px:].all())
    _save(moved, None if out_dir is None else os.path.join(out_dir, "homography_shift.png"))
    r = torch.from_numpy(HomographySample.euler_to_rotation_matrix(0, 10, 0, degrees=True))
    g = torch.eye(4)[None].clone()
    g[0, :3, :3] = r
    turned, _ = sampler.sample(img, depth, g, kinv, k)
    _save(turned, None if out_dir is None else os.path.join(out_dir, "homography_yaw.png"))
    return moved


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None, help="directory for PNG dumps")
    a = ap.parse_args()
    test_mpi_composition(a.out)
    rotation_test()
    test_homography_sample(a.out)
    print("rendering checks passed" + ("" if a.out is None else "; images in " + a.out))
