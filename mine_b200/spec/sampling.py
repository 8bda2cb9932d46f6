"""Disparity-plane sampling, sparse-point gathers and inverse-CDF refinement (spec).

Semantics: SURVEY 2.7 "Disparity sampling"; reference ``operations/rendering_utils.py`` and
``synthesis_task.py:31-60``.  Planes are ordered near -> far (disparity descending).
"""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np
import torch


def stratified_disparity(batch: int, edges: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """One uniform sample inside each of the S bins delimited by ``edges`` (S+1, descending)."""
    if not bool(edges[0] > edges[-1]):
        raise ValueError("disparity bin edges must be descending (near plane first)")
    lo, width = edges[:-1], edges[1:] - edges[:-1]
    u = torch.rand((batch, lo.numel()), dtype=edges.dtype, device=edges.device, generator=generator)
    return lo[None] + width[None] * u


def stratified_disparity_linspace(batch: int, num_bins: int, start: float, end: float, device=None,
                                  dtype=torch.float32, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    if not start > end:
        raise ValueError("mpi.disparity_start must be larger than mpi.disparity_end")
    edges = torch.linspace(start, end, num_bins + 1, dtype=dtype, device=device)
    # the reference uses the *first* interval width for every bin (identical for a linspace)
    u = torch.rand((batch, num_bins), dtype=dtype, device=device, generator=generator)
    return edges[None, :-1] + (edges[1] - edges[0]) * u


def fixed_disparity(batch: int, num_bins: int, start: float, end: float, device=None, dtype=torch.float32) -> torch.Tensor:
    return torch.linspace(start, end, num_bins, dtype=dtype, device=device)[None].repeat(batch, 1)


def disparity_planes(config: Mapping, batch: int, device=None, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Plane disparities ``[B,S]`` for a forward pass (reference ``_get_disparity_list``).

    * ``mpi.fix_disparity`` + explicit ``mpi.disparity_list`` (S+1 values): use entries 1..S
    * ``mpi.fix_disparity``: ``linspace(start, end, S)``
    * explicit list: stratified sample inside the given bin edges
    * default: stratified sample inside ``linspace(start, end, S+1)`` bins
    """
    s = int(config["mpi.num_bins_coarse"])
    start, end = float(config["mpi.disparity_start"]), float(config["mpi.disparity_end"])
    explicit = config.get("mpi.disparity_list", None)
    has_list = explicit is not None and len(explicit) == s + 1
    if has_list:
        edges = torch.as_tensor(np.asarray(explicit), dtype=torch.float32, device=device)
    if config.get("mpi.fix_disparity", False):
        if has_list:
            return edges[1:][None].repeat(batch, 1)
        return fixed_disparity(batch, s, start, end, device=device)
    if has_list:
        return stratified_disparity(batch, edges, generator)
    return stratified_disparity_linspace(batch, s, start, end, device=device, generator=generator)


def project_points(k: torch.Tensor, xyz: torch.Tensor) -> torch.Tensor:
    """Pinhole projection of camera-frame points ``[B,3,N]`` -> pixel coords ``[B,2,N]``."""
    p = k @ xyz
    return p[:, :2] / p[:, 2:3]


def gather_nearest(img: torch.Tensor, pxpy: torch.Tensor) -> torch.Tensor:
    """Nearest-pixel lookup with clamping: ``img [B,C,H,W]``, ``pxpy [B,2,N]`` -> ``[B,C,N]``
    (reference ``gather_pixel_by_pxpy``; rounding is half-to-even like ``torch.round``)."""
    b, c, h, w = img.shape
    with torch.no_grad():
        ix = torch.round(pxpy[:, 0]).long().clamp(0, w - 1)
        iy = torch.round(pxpy[:, 1]).long().clamp(0, h - 1)
        flat = (iy * w + ix)[:, None].expand(-1, c, -1)
    return torch.gather(img.reshape(b, c, h * w), 2, flat)


def sample_pdf(values: torch.Tensor, weights: torch.Tensor, n_samples: int,
               generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Inverse-CDF sampling of ``n_samples`` new plane disparities from a piecewise-constant pdf
    (``values``/``weights``: ``[B,1,N,S]``) - coarse-to-fine refinement, reference ``sample_pdf``."""
    b, _, n, s = weights.shape
    mid = 0.5 * (values[..., 1:] + values[..., :-1])
    edges = torch.cat([values[..., :1], mid, values[..., -1:]], dim=-1)            # S+1
    pdf = weights / (weights.sum(dim=-1, keepdim=True) + 1e-5)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, dim=-1)], dim=-1)
    u = torch.rand((b, 1, n, n_samples), dtype=weights.dtype, device=weights.device, generator=generator)
    idx = torch.searchsorted(cdf.contiguous(), u.contiguous(), right=True)
    lo = (idx - 1).clamp(min=0)
    hi = idx.clamp(max=s)
    cdf_lo, cdf_hi = torch.gather(cdf, 3, lo), torch.gather(cdf, 3, hi)
    bin_lo, bin_hi = torch.gather(edges, 3, lo), torch.gather(edges, 3, hi)
    span = cdf_hi - cdf_lo
    t = (u - cdf_lo) / span.clamp(min=1e-5)
    t = torch.where(span <= 1e-4, torch.full_like(t, 0.5), t)
    return bin_lo + t * (bin_hi - bin_lo)


def refine_disparity(disparity_coarse: torch.Tensor, plane_weights: torch.Tensor, n_fine: int,
                     generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Coarse-to-fine: ``plane_weights [B,S,1,H,W]`` from a no-grad coarse render -> sorted
    (descending) union of coarse and ``n_fine`` importance-sampled planes
    (reference ``predict_mpi_coarse_to_fine``, ``mpi_rendering.py:244-268``)."""
    w = plane_weights.mean(dim=(2, 3, 4))[:, None, None, :]
    fine = sample_pdf(disparity_coarse[:, None, None, :], w, n_fine, generator)[:, 0, 0]
    merged = torch.cat([disparity_coarse, fine], dim=1)
    return torch.sort(merged, dim=1, descending=True).values
