"""NeRF-LLFF dataset: COLMAP poses + sparse points, all images resident in RAM.

Capability parity with reference ``input_pipelines/llff/nerf_dataset.py:32-234``:
per scene ``sparse/0/*.bin`` is read, every image under ``images[_<ratio>][_val]`` is resized
(bicubic) to ``img_w x img_h`` and kept as a float tensor; K comes from the SIMPLE_RADIAL focal /
principal point rescaled to the working resolution; ``__getitem__`` returns the source item plus
``supervision_count`` random target views of the same scene (validation: the deterministic
"next" view) and ``visible_points_count`` random sparse points per view.

Differences: validation point sampling is deterministic (the reference's TODO),
per-worker RNG instead of the global ``random`` module, scenes with too few points raise a
readable error instead of a bare assert.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.utils.data as data

from . import colmap


def _collate_fn(batch):
    """``[(src_item, tgt_items)]`` -> ``(src_items dict of B.., tgt_items dict of B,L,..)``."""
    srcs, tgts = zip(*batch)
    src = {k: torch.stack([torch.as_tensor(s[k]) for s in srcs]) for k in srcs[0] if k != "G_cam_world"}
    tgt = {k: torch.stack([torch.stack([torch.as_tensor(v) for v in t[k]]) for t in tgts]) for k in tgts[0]}
    return src, tgt


class NeRFDataset(data.Dataset):
    def __init__(self, config, logger, root: str, is_validation: bool, img_size: Tuple[int, int],
                 supervision_count: int = 5, visible_points_count: int = 8,
                 img_pre_downsample_ratio: Optional[float] = 7.875, seed: int = 0):
        from PIL import Image as PILImage
        self.config, self.logger = config, logger
        self.img_w, self.img_h = int(img_size[0]), int(img_size[1])
        self.is_validation = is_validation
        self.visible_points_count = visible_points_count
        self.supervision_count = supervision_count
        self.collate_fn = _collate_fn
        self._rng = np.random.default_rng(seed)

        ratio = img_pre_downsample_ratio
        folder = "images" if (ratio is None or ratio <= 1) else "images_" + str(ratio)
        if is_validation:
            folder += "_val"
        ratio = 1.0 if (ratio is None or ratio <= 1) else float(ratio)
        self.image_folder = folder

        self.items: List[Dict] = []
        self.scene_of: List[str] = []
        self.scene_to_indices: Dict[str, List[int]] = defaultdict(list)
        for scene in sorted(os.listdir(root)):
            scene_dir = os.path.join(root, scene)
            model_dir = os.path.join(scene_dir, "sparse", "0")
            if not os.path.isdir(model_dir):
                continue
            cameras, images, points3d = colmap.read_model(model_dir, ext=".bin")
            if len(cameras) != 1:
                raise ValueError(f"{scene}: expected a single shared camera, found {len(cameras)}")
            for img_id in sorted(images):
                meta = images[img_id]
                path = os.path.join(scene_dir, folder, meta.name)
                if not os.path.exists(path):
                    continue
                pil = PILImage.open(path).convert("RGB")
                w0, h0 = pil.size
                pil = pil.resize((self.img_w, self.img_h), PILImage.BICUBIC)
                img = torch.from_numpy(np.asarray(pil, dtype=np.float32) / 255.0).permute(2, 0, 1).contiguous()
                tracked = meta.point3D_ids != -1
                item = self._build_item(img, meta, cameras[meta.camera_id].params,
                                        (w0 * ratio / self.img_w, h0 * ratio / self.img_h),
                                        meta.point3D_ids[tracked],
                                        np.stack([points3d[int(p)].xyz for p in meta.point3D_ids[tracked]])
                                        if tracked.any() else np.zeros((0, 3)))
                if item["xyzs"].shape[1] < visible_points_count:
                    raise ValueError(f"{path}: only {item['xyzs'].shape[1]} tracked points, need {visible_points_count}")
                self.scene_to_indices[scene].append(len(self.items))
                self.items.append(item)
                self.scene_of.append(scene)
        if self.logger:
            self.logger.info("Dataset root: {}, is_validation: {}, number of images: {}".format(
                root, self.is_validation, len(self.items)))

    @staticmethod
    def _build_item(img, meta, cam_params, ratio_xy, point_ids, xyz_world) -> Dict:
        rx, ry = ratio_xy
        g = np.eye(4, dtype=np.float32)
        g[:3, :3] = colmap.qvec2rotmat(meta.qvec)
        g[:3, 3] = meta.tvec
        f, px, py = cam_params[0], cam_params[1], cam_params[2]          # SIMPLE_RADIAL: f, cx, cy, k
        k = np.array([[f / rx, 0, px / rx], [0, f / ry, py / ry], [0, 0, 1]], dtype=np.float32)
        xyz_cam = (g[:3, :3] @ xyz_world.T.astype(np.float32) + g[:3, 3:4]).astype(np.float32)   # 3,N
        # depth along the principal axis (sign/norm handling of the projective depth formula;
        # for a pinhole K and rigid G this is simply z)
        p = k @ np.concatenate([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)], 1) @ g
        depths = np.sign(np.linalg.det(p[:, :3])) * (k @ xyz_cam)[2] / np.linalg.norm(p[2, :3])
        return {"img": img, "G_cam_world": g, "K": k, "K_inv": np.linalg.inv(k).astype(np.float32),
                "xyzs": xyz_cam, "xyzs_ids": np.asarray(point_ids, dtype=np.int64), "depths": depths.astype(np.float32)}

    def __len__(self):
        return len(self.items)

    def _pick_points(self, item, index: int):
        n = item["xyzs"].shape[1]
        if self.is_validation:
            sel = np.random.default_rng(index).choice(n, self.visible_points_count, replace=False)
        else:
            sel = self._rng.choice(n, self.visible_points_count, replace=False)
        return sel

    def __getitem__(self, index: int):
        base = self.items[index]
        sel = self._pick_points(base, index)
        src = {"img": base["img"], "G_cam_world": base["G_cam_world"], "K": base["K"], "K_inv": base["K_inv"],
               "xyzs": base["xyzs"][:, sel], "xyzs_ids": base["xyzs_ids"][sel], "depths": base["depths"][sel]}
        others = [i for i in self.scene_to_indices[self.scene_of[index]] if i != index]
        if self.is_validation:
            picks = [others[(index + 1) % len(others) - 1]]
        else:
            picks = list(self._rng.choice(others, self.supervision_count, replace=False))
        tgt = defaultdict(list)
        for j in picks:
            it = self.items[int(j)]
            s2 = self._pick_points(it, int(j) + 7919 * index)
            tgt["img"].append(it["img"])
            tgt["K"].append(it["K"])
            tgt["K_inv"].append(it["K_inv"])
            tgt["G_src_tgt"].append((base["G_cam_world"] @ np.linalg.inv(it["G_cam_world"])).astype(np.float32))
            tgt["xyzs"].append(it["xyzs"][:, s2])
            tgt["xyzs_ids"].append(it["xyzs_ids"][s2])
            tgt["depths"].append(it["depths"][s2])
        return src, dict(tgt)


def resize_llff_images(root: str, ratio: float, val_every: int = 0) -> int:
    """Offline helper: write ``images_<ratio>/`` downsampled copies for every scene
    (reference ``input_pipelines/llff/misc/resize_nerf_llff_images.py``).

    ``val_every = n > 0`` additionally creates the held-out split the dataset class looks for
    (``images_<ratio>_val/``, reference ``nerf_dataset.py:52-53``): every n-th image of a scene (the usual LLFF
    convention is 8) goes there instead of the training folder.  Upstream leaves that split to the user."""
    import cv2
    n = 0
    for scene in sorted(os.listdir(root)):
        src_dir = os.path.join(root, scene, "images")
        if not os.path.isdir(src_dir):
            continue
        train_dir = os.path.join(root, scene, "images_" + str(ratio))
        val_dir = train_dir + "_val"
        os.makedirs(train_dir, exist_ok=True)
        if val_every > 0:
            os.makedirs(val_dir, exist_ok=True)
        for i, name in enumerate(sorted(os.listdir(src_dir))):
            img = cv2.imread(os.path.join(src_dir, name), cv2.IMREAD_COLOR)
            if img is None:
                continue
            h, w = img.shape[:2]
            out = cv2.resize(img, (int(round(w / ratio)), int(round(h / ratio))), interpolation=cv2.INTER_AREA)
            held_out = val_every > 0 and i % val_every == 0
            cv2.imwrite(os.path.join(val_dir if held_out else train_dir, name), out)
            n += 1
    return n
