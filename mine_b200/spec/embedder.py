"""NeRF positional encoding of the plane disparity (spec).

``gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]`` -> 1 + 2L dims
(21 for ``model.pos_encoding_multires = 10``).  Reference ``utils.py:144-193``.
"""
from __future__ import annotations

import torch


def embedding_dim(multires: int, input_dims: int = 1) -> int:
    return input_dims * (1 + 2 * multires)


def positional_encoding(x: torch.Tensor, multires: int) -> torch.Tensor:
    """``x [..., D]`` -> ``[..., D*(1+2*multires)]`` in the reference channel order."""
    freqs = 2.0 ** torch.arange(multires, dtype=x.dtype, device=x.device)          # 2^0 .. 2^(L-1)
    xf = x[..., None, :] * freqs[:, None]                                            # ..., L, D
    sc = torch.stack([torch.sin(xf), torch.cos(xf)], dim=-2)                         # ..., L, 2, D
    return torch.cat([x, sc.reshape(*x.shape[:-1], -1)], dim=-1)


def get_embedder(multires: int, i: int = 0):
    """Reference-style factory: returns ``(embed_fn, out_dim)``."""
    return (lambda x: positional_encoding(x, multires)), embedding_dim(multires)
