"""Synthetic source/target batches of the named benchmark shapes (no datasets offline).

Produces exactly the collated batch format the task consumes (SURVEY 2.6 "Batch format"):
``src_items = {img B,3,H,W; K; K_inv; xyzs B,3,N; xyzs_ids; depths}``, ``tgt_items = {img B,L,3,H,W;
K; K_inv; xyzs B,L,3,N; G_src_tgt B,L,4,4; ...}``.  Geometry is plausible: a smooth random scene
depth, a small rigid motion between the views, sparse 3-D points that lie on the scene surface,
have positive depth in both cameras and project inside both images.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch


def _rot(ax: float, ay: float, az: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _smooth_image(rng: np.random.Generator, h: int, w: int) -> np.ndarray:
    """Low-frequency colour field + a little texture, in [0,1], CHW float32."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((3, h, w), np.float32)
    for c in range(3):
        for _ in range(4):
            fx, fy = rng.uniform(0.5, 6.0, 2) * 2 * math.pi
            ph = rng.uniform(0, 2 * math.pi)
            img[c] += rng.uniform(0.1, 0.3) * np.sin(fx * xx / w + fy * yy / h + ph)
    img += rng.normal(0, 0.03, img.shape).astype(np.float32)
    img = (img - img.min()) / max(img.max() - img.min(), 1e-6)
    return img.astype(np.float32)


def make_pair(rng: np.random.Generator, h: int, w: int, n_pt: int, depth_range=(1.5, 20.0),
              max_rot: float = 0.03, max_trans: float = 0.15) -> Tuple[Dict, Dict]:
    f = rng.uniform(0.9, 1.3) * w
    k = np.array([[f, 0, w * 0.5], [0, f, h * 0.5], [0, 0, 1]], np.float32)
    k_inv = np.linalg.inv(k).astype(np.float32)
    g_tgt_src = np.eye(4, dtype=np.float32)
    g_tgt_src[:3, :3] = _rot(*rng.uniform(-max_rot, max_rot, 3))
    g_tgt_src[:3, 3] = rng.uniform(-max_trans, max_trans, 3)
    g_src_tgt = np.linalg.inv(g_tgt_src).astype(np.float32)

    def points(cam_from_src: np.ndarray):
        # surface points seen from the source camera, then expressed in the requested camera
        out = np.zeros((3, 0), np.float32)
        while out.shape[1] < n_pt:
            m = 4 * n_pt
            u = rng.uniform(0.05 * w, 0.95 * w, m)
            v = rng.uniform(0.05 * h, 0.95 * h, m)
            z = np.exp(rng.uniform(math.log(depth_range[0]), math.log(depth_range[1]), m))
            xyz_src = (k_inv @ np.stack([u, v, np.ones(m)]).astype(np.float32)) * z
            xyz = cam_from_src[:3, :3] @ xyz_src + cam_from_src[:3, 3:4]
            p = k @ xyz
            ok = (xyz[2] > 0.5) & (p[0] / p[2] > 1) & (p[0] / p[2] < w - 2) & (p[1] / p[2] > 1) & (p[1] / p[2] < h - 2)
            out = np.concatenate([out, xyz[:, ok].astype(np.float32)], 1)
        return out[:, :n_pt]

    xyz_s = points(np.eye(4, dtype=np.float32))
    xyz_t = points(g_tgt_src)
    src = {"img": _smooth_image(rng, h, w), "K": k, "K_inv": k_inv, "xyzs": xyz_s,
           "xyzs_ids": np.arange(n_pt, dtype=np.int64), "depths": xyz_s[2].copy()}
    tgt = {"img": _smooth_image(rng, h, w)[None], "K": k[None], "K_inv": k_inv[None], "xyzs": xyz_t[None],
           "G_src_tgt": g_src_tgt[None], "xyzs_ids": np.arange(n_pt, dtype=np.int64)[None], "depths": xyz_t[2][None].copy()}
    return src, tgt


def collate(samples) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    srcs, tgts = zip(*samples)
    s = {k: torch.from_numpy(np.stack([d[k] for d in srcs])) for k in srcs[0]}
    t = {k: torch.from_numpy(np.stack([d[k] for d in tgts])) for k in tgts[0]}
    return s, t


class SyntheticPairs(torch.utils.data.Dataset):
    """Deterministic synthetic dataset (item i is a pure function of ``seed`` and ``i``)."""

    def __init__(self, length: int, h: int, w: int, n_pt: int = 256, seed: int = 0, **kw):
        self.length, self.h, self.w, self.n_pt, self.seed, self.kw = length, h, w, n_pt, seed, kw
        self.collate_fn = collate

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        return make_pair(np.random.default_rng(self.seed * 1000003 + i), self.h, self.w, self.n_pt, **self.kw)


def synthetic_batch(batch: int, h: int, w: int, n_pt: int = 256, seed: int = 0, pin: bool = False, **kw):
    ds = SyntheticPairs(batch, h, w, n_pt, seed, **kw)
    s, t = collate([ds[i] for i in range(batch)])
    if pin and torch.cuda.is_available():
        s = {k: v.pin_memory() for k, v in s.items()}
        t = {k: v.pin_memory() for k, v in t.items()}
    return s, t


def config_batch(config, seed: int = 0, pin: bool = False):
    return synthetic_batch(int(config["data.per_gpu_batch_size"]), int(config["data.img_h"]), int(config["data.img_w"]),
                           int(config["data.visible_point_count"]), seed=seed, pin=pin)
