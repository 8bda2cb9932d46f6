"""Import-time shims that let the UNMODIFIED upstream reference run on this image.

Nothing here edits reference code.  The reference imports a few packages that are not in the
offline image (``lpips``, ``kornia``, ``matplotlib``, ``moviepy``) and uses two APIs that newer
numpy/pyyaml removed; we register minimal stand-ins in ``sys.modules`` before importing it:

* ``kornia.filters.spatial_gradient`` - Sobel operator with replicate padding (same maths as
  kornia 0.3: kernel normalised by 8 when ``normalized=True``), output ``[B,C,2,H,W]``.
* ``lpips.LPIPS`` - a module returning zeros (weights are not available offline; the metric is
  only evaluated in validation on rank 0, never inside the timed training step).
* ``matplotlib.pyplot`` / ``moviepy.editor`` - empty modules (imported, never called in training).

``isolated_reference_imports`` temporarily puts the reference root first on ``sys.path`` and
hides this repository's same-named shim modules (``utils``, ``operations`` ...), restoring
everything afterwards, so oracle tests can import both trees in one process.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch
import torch.nn.functional as F

_REF_TOPLEVEL = ("utils", "operations", "network", "input_pipelines", "synthesis_task", "train",
                 "visualizations")


def _sobel_spatial_gradient(x, mode="sobel", order=1, normalized=True):
    b, c, h, w = x.shape
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=x.dtype, device=x.device)
    if normalized:
        kx = kx / kx.abs().sum()
    k = torch.stack([kx, kx.t()])[:, None]
    xp = F.pad(x.reshape(b * c, 1, h, w), (1, 1, 1, 1), mode="replicate")
    return F.conv2d(xp, k).reshape(b, c, 2, h, w)


class _ZeroLPIPS(torch.nn.Module):
    def __init__(self, net="vgg", **kw):
        super().__init__()

    def forward(self, a, b):
        return torch.zeros(a.shape[0], 1, 1, 1, device=a.device, dtype=a.dtype)


def install_import_shims() -> None:
    if "kornia" not in sys.modules:
        try:
            import kornia  # noqa: F401
        except Exception:
            k = types.ModuleType("kornia")
            kf = types.ModuleType("kornia.filters")
            kf.spatial_gradient = _sobel_spatial_gradient
            k.filters = kf
            sys.modules["kornia"], sys.modules["kornia.filters"] = k, kf
    if "lpips" not in sys.modules:
        try:
            import lpips  # noqa: F401
        except Exception:
            m = types.ModuleType("lpips")
            m.LPIPS = _ZeroLPIPS
            sys.modules["lpips"] = m
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            mp = types.ModuleType("matplotlib.pyplot")
            m.pyplot = mp
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, mp
    if "moviepy" not in sys.modules:
        try:
            import moviepy  # noqa: F401
        except Exception:
            m = types.ModuleType("moviepy")
            me = types.ModuleType("moviepy.editor")

            class ImageSequenceClip:  # pragma: no cover - only used by the video script
                def __init__(self, frames, fps=30):
                    self.frames, self.fps = frames, fps

                def write_videofile(self, path, fps=None, **kw):
                    from mine_b200.utils.video_io import write_video
                    write_video(path, self.frames, fps or self.fps)

            me.ImageSequenceClip = ImageSequenceClip
            m.editor = me
            sys.modules["moviepy"], sys.modules["moviepy.editor"] = m, me


@contextlib.contextmanager
def isolated_reference_imports(root: str):
    """Import reference modules from ``root`` without leaking them over this repo's modules."""
    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in _REF_TOPLEVEL}
    for k in saved:
        del sys.modules[k]
    ref_mods = getattr(isolated_reference_imports, "_ref_modules", {}).get(root, {})
    sys.modules.update(ref_mods)
    saved_path = list(sys.path)
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    # this repository ships same-named shim packages (regular packages would shadow the reference's
    # namespace packages), so hide the repo root while importing the oracle
    sys.path[:] = [root] + [p for p in sys.path
                            if os.path.abspath(p or os.getcwd()) not in (repo, os.path.abspath(root))]
    real_sync = torch.cuda.synchronize
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    try:
        yield
    finally:
        if not torch.cuda.is_available():
            torch.cuda.synchronize = real_sync  # restored; reference fns called later re-patch below
        sys.path[:] = saved_path
        now = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_TOPLEVEL}
        store = getattr(isolated_reference_imports, "_ref_modules", {})
        store.setdefault(root, {}).update(now)
        isolated_reference_imports._ref_modules = store
        for k in now:
            del sys.modules[k]
        sys.modules.update(saved)


@contextlib.contextmanager
def cpu_cuda_sync_noop():
    """The reference calls ``torch.cuda.synchronize()`` unconditionally (utils.py:106)."""
    real = torch.cuda.synchronize
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    try:
        yield
    finally:
        torch.cuda.synchronize = real
