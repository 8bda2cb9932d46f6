"""LPIPS-VGG perceptual metric (validation only, rank 0, scale 0).

The reference uses the external ``lpips`` package (``synthesis_task.py:4,91-92,342``).  Its weights
(torchvision VGG16 + the learned 1x1 "lin" layers) cannot be downloaded here, so the network
structure is implemented directly and weights are looked up locally
(``MINE_LPIPS_WEIGHTS`` = a torch file with ``{"vgg": vgg16.features state dict, "lin": [5 tensors
of shape 1xCx1x1]}``).  Without weights the metric is disabled: it reports 0 and says so once -
it is a logging-only quantity and never part of the optimised loss.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

_VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
_TAPS = (3, 8, 15, 22, 29)        # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 in vgg16.features indexing
_CHNS = (64, 128, 256, 512, 512)


class LPIPSVGG(nn.Module):
    def __init__(self):
        super().__init__()
        layers, cin = [], 3
        for v in _VGG_CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=False)]
                cin = v
        self.features = nn.Sequential(*layers)
        self.lins = nn.ParameterList([nn.Parameter(torch.ones(1, c, 1, 1) / c) for c in _CHNS])
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188]).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor([.458, .448, .450]).view(1, 3, 1, 1))
        for p in self.parameters():
            p.requires_grad_(False)

    def _feats(self, x):
        out = []
        for i, layer in enumerate(self.features):
            x = layer(x)
            if i in _TAPS:
                out.append(x)
        return out

    @torch.no_grad()
    def forward(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """Inputs in [0,1] (the reference passes them un-normalised as well)."""
        fa = self._feats((a - self.shift) / self.scale)
        fb = self._feats((b - self.shift) / self.scale)
        total = 0
        for x, y, w in zip(fa, fb, self.lins):
            xn = x / (x.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            yn = y / (y.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            total = total + ((xn - yn).pow(2) * w).sum(1, keepdim=True).mean(dim=(2, 3), keepdim=True)
        return total


_warned = False


def build_lpips(device, logger=None) -> Optional[LPIPSVGG]:
    global _warned
    path = os.environ.get("MINE_LPIPS_WEIGHTS")
    if not path or not os.path.exists(path):
        if logger is not None and not _warned:
            logger.info("LPIPS weights not found (MINE_LPIPS_WEIGHTS); lpips_tgt will be reported as 0")
            _warned = True
        return None
    sd = torch.load(path, map_location="cpu")
    net = LPIPSVGG()
    net.features.load_state_dict(sd["vgg"])
    for p, w in zip(net.lins, sd["lin"]):
        p.data.copy_(w.reshape(p.shape))
    return net.to(device).eval()
