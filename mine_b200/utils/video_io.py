"""Frame conversion + video writing without moviepy (OpenCV's encoder; falls back to PNG frames)."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np
import torch


def img_tensor_to_np(img: torch.Tensor, colormap: bool = True) -> np.ndarray:
    """``[1,C,H,W]`` in [0,1] -> uint8 HWC RGB; single-channel maps get the HOT colour map
    (reference ``image_to_video.img_tensor_to_np``)."""
    import cv2
    b, c, h, w = img.shape
    if b != 1 or c not in (1, 3):
        raise ValueError("expected 1x{1,3}xHxW")
    arr = img[0].permute(1, 2, 0).detach().float().cpu().numpy()
    arr = np.clip(np.round(arr * 255.0), 0, 255).astype(np.uint8)
    if c == 1 and colormap:
        arr = cv2.cvtColor(cv2.applyColorMap(arr, cv2.COLORMAP_HOT), cv2.COLOR_BGR2RGB)
    return arr


def to_uint8_device(img: torch.Tensor) -> torch.Tensor:
    """Device-side ``round(x*255)`` clamp + NHWC so only 1 byte/px crosses PCIe."""
    return (img.clamp(0, 1) * 255.0 + 0.5).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def write_img_to_disk(img: torch.Tensor, step: int, postfix: str, output_dir: str) -> List[str]:
    import cv2
    os.makedirs(output_dir, exist_ok=True)
    paths = []
    for b in range(img.shape[0]):
        arr = img_tensor_to_np(img[b:b + 1], colormap=False)
        p = os.path.join(output_dir, "%d_%d_%s.png" % (step, b, postfix))
        cv2.imwrite(p, cv2.cvtColor(arr, cv2.COLOR_RGB2BGR) if arr.shape[2] == 3 else arr[:, :, 0])
        paths.append(p)
    return paths


def write_video(path: str, frames: Sequence[np.ndarray], fps: int = 30) -> str:
    """Write RGB uint8 frames to ``path`` (.mp4).  Returns the path actually written (a directory of
    PNGs if no encoder is usable in this OpenCV build)."""
    import cv2
    frames = list(frames)
    h, w = frames[0].shape[:2]
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    for fourcc in ("mp4v", "avc1", "MJPG"):
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), float(fps), (w, h))
        if vw.isOpened():
            for f in frames:
                vw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
            vw.release()
            if os.path.exists(path) and os.path.getsize(path) > 0:
                return path
    out_dir = os.path.splitext(path)[0] + "_frames"
    os.makedirs(out_dir, exist_ok=True)
    for i, f in enumerate(frames):
        cv2.imwrite(os.path.join(out_dir, "%05d.png" % i), cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    return out_dir
