"""ResNet-50 image encoder (five-level feature pyramid).

Capability parity with reference ``network/monodepth2/resnet_encoder.py:63-108``: ImageNet
mean/std normalisation inside ``forward`` and the outputs ``conv1 (64,H/2)``, ``layer1 (256,H/4)``,
``layer2 (512,H/8)``, ``layer3 (1024,H/16)``, ``layer4 (2048,H/32)``.  Differences by design:

* written from scratch (no torchvision model object), state-dict names under ``encoder.`` equal
  torchvision's so reference checkpoints load; the classifier head ``fc`` - never used upstream yet
  all-reduced as zeros every step and the reason for ``find_unused_parameters=True``
  - does not exist here (the checkpoint adapter drops/re-adds it);
* mean/std are buffers that follow ``.to(device)`` instead of tensors pinned to ``cuda:0``;
* BatchNorm is :class:`mine_b200.models.norm.BatchNorm` (cross-replica stats via our reducer).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .norm import BatchNorm

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: bool = False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = BatchNorm(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)   # v1.5: stride on 3x3
        self.bn2 = BatchNorm(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = BatchNorm(planes * 4)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                            BatchNorm(planes * 4))
        self.stride = stride

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return F.relu(out + idt)


class ResNet50Trunk(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm(64)
        inplanes = 64
        for li, (planes, blocks, stride) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)], 1):
            layers = [Bottleneck(inplanes, planes, stride, downsample=True)]
            inplanes = planes * 4
            layers += [Bottleneck(inplanes, planes) for _ in range(blocks - 1)]
            setattr(self, f"layer{li}", nn.Sequential(*layers))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class ResnetEncoder(nn.Module):
    """``ResnetEncoder(num_layers=50, pretrained=False)`` - reference constructor signature."""

    def __init__(self, num_layers: int = 50, pretrained: bool = False, num_input_images: int = 1,
                 pretrained_path: str | None = None):
        super().__init__()
        if num_layers != 50 or num_input_images != 1:
            raise ValueError("MINE uses a single-image ResNet-50 encoder")
        self.num_ch_enc = [64, 256, 512, 1024, 2048]
        self.encoder = ResNet50Trunk()
        self.register_buffer("img_mean", torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("img_std", torch.tensor(IMAGENET_STD).view(1, 3, 1, 1), persistent=False)
        if pretrained:
            self._load_imagenet(pretrained_path)

    def _load_imagenet(self, path):
        """ImageNet initialisation from a local torchvision ``resnet50`` state dict (no network
        access here; the reference downloads it, ``resnet_encoder.py:71-83``)."""
        import os
        cands = [path, os.environ.get("MINE_RESNET50_WEIGHTS"),
                 os.path.expanduser("~/.cache/torch/hub/checkpoints/resnet50-0676ba61.pth"),
                 os.path.expanduser("~/.cache/torch/hub/checkpoints/resnet50-19c8e357.pth")]
        for c in cands:
            if c and os.path.exists(c):
                sd = torch.load(c, map_location="cpu")
                sd = {k: v for k, v in sd.items() if not k.startswith("fc.")}
                self.encoder.load_state_dict(sd, strict=True)
                return
        import warnings
        warnings.warn("model.imagenet_pretrained=true but no local resnet50 weights were found "
                      "(set MINE_RESNET50_WEIGHTS); continuing with random initialisation")

    def forward(self, img: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        e = self.encoder
        x = (img - self.img_mean.to(img.dtype)) / self.img_std.to(img.dtype)
        c1 = F.relu(e.bn1(e.conv1(x)))
        b1 = e.layer1(F.max_pool2d(c1, 3, 2, 1))
        b2 = e.layer2(b1)
        b3 = e.layer3(b2)
        b4 = e.layer4(b3)
        return c1, b1, b2, b3, b4
