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

from typing import Tuple

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


class BasicBlock(nn.Module):
    """Two 3x3 convolutions (ResNet-18/34)."""
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: bool = False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = BatchNorm(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = BatchNorm(planes)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False), BatchNorm(planes))
        self.stride = stride

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + idt)


RESNET_LAYOUTS = {18: (BasicBlock, (2, 2, 2, 2)), 34: (BasicBlock, (3, 4, 6, 3)), 50: (Bottleneck, (3, 4, 6, 3)),
                  101: (Bottleneck, (3, 4, 23, 3)), 152: (Bottleneck, (3, 8, 36, 3))}


class ResNetTrunk(nn.Module):
    """ResNet stem + four stages with torchvision's parameter names, no classifier head.

    ``num_input_images > 1`` widens the stem to ``3 * num_input_images`` input channels - the multi-frame
    variant the reference vendors from monodepth2 (``resnet_encoder.py:18-60``, ``ResNetMultiImageInput``).
    """

    def __init__(self, num_layers: int = 50, num_input_images: int = 1):
        super().__init__()
        if num_layers not in RESNET_LAYOUTS:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        block, counts = RESNET_LAYOUTS[num_layers]
        self.conv1 = nn.Conv2d(3 * num_input_images, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm(64)
        inplanes = 64
        for li, (planes, blocks, stride) in enumerate(zip((64, 128, 256, 512), counts, (1, 2, 2, 2)), 1):
            need_proj = stride != 1 or inplanes != planes * block.expansion
            layers = [block(inplanes, planes, stride, downsample=need_proj)]
            inplanes = planes * block.expansion
            layers += [block(inplanes, planes) for _ in range(blocks - 1)]
            setattr(self, f"layer{li}", nn.Sequential(*layers))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class ResNet50Trunk(ResNetTrunk):
    def __init__(self):
        super().__init__(50, 1)


class ResNetMultiImageInput(ResNetTrunk):
    """Name kept from the reference (``resnet_encoder.py:18``): a trunk fed ``num_input_images`` stacked frames."""

    def __init__(self, num_layers: int = 18, num_input_images: int = 1):
        super().__init__(num_layers, num_input_images)


def adapt_stem_to_multi_image(state: dict, num_input_images: int) -> dict:
    """ImageNet stem -> multi-frame stem: tile ``conv1.weight`` over the frames and divide by their number so the
    response to identical frames is unchanged (reference ``resnet_encoder.py:56-58``)."""
    out = dict(state)
    out["conv1.weight"] = torch.cat([state["conv1.weight"]] * num_input_images, dim=1) / num_input_images
    return out


def resnet_multiimage_input(num_layers: int, pretrained: bool = False, num_input_images: int = 1,
                            pretrained_path: str | None = None) -> ResNetMultiImageInput:
    """Reference factory (``resnet_encoder.py:41-60``): 18- or 50-layer multi-frame trunk, optionally initialised
    from local ImageNet weights (there is no download here; see :meth:`ResnetEncoder._find_imagenet`)."""
    assert num_layers in (18, 50), "Can only run with 18 or 50 layer resnet"
    model = ResNetMultiImageInput(num_layers, num_input_images)
    if pretrained:
        path = ResnetEncoder._find_imagenet(pretrained_path, num_layers)
        if path is not None:
            sd = {k: v for k, v in torch.load(path, map_location="cpu").items() if not k.startswith("fc.")}
            model.load_state_dict(adapt_stem_to_multi_image(sd, num_input_images), strict=True)
    return model


class ResnetEncoder(nn.Module):
    """``ResnetEncoder(num_layers, pretrained, num_input_images=1)`` - reference constructor signature
    (``resnet_encoder.py:63-95``); MINE itself always builds the single-image 50-layer variant."""

    def __init__(self, num_layers: int = 50, pretrained: bool = False, num_input_images: int = 1,
                 pretrained_path: str | None = None):
        super().__init__()
        if num_layers not in RESNET_LAYOUTS:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        self.num_layers, self.num_input_images = num_layers, num_input_images
        exp = RESNET_LAYOUTS[num_layers][0].expansion
        self.num_ch_enc = [64] + [c * exp for c in (64, 128, 256, 512)]       # x4 beyond ResNet-34 (reference :84-85)
        if num_input_images > 1:
            self.encoder = resnet_multiimage_input(num_layers, pretrained, num_input_images, pretrained_path)
        else:
            self.encoder = ResNetTrunk(num_layers, 1)
        mean = torch.tensor(IMAGENET_MEAN * num_input_images).view(1, -1, 1, 1)
        std = torch.tensor(IMAGENET_STD * num_input_images).view(1, -1, 1, 1)
        self.register_buffer("img_mean", mean, persistent=False)
        self.register_buffer("img_std", std, persistent=False)
        self.pretrained_requested = bool(pretrained)
        self.pretrained_source = None
        if pretrained and num_input_images == 1:
            self._load_imagenet(pretrained_path)

    @staticmethod
    def _find_imagenet(path, num_layers: int = 50):
        """Local torchvision ``resnet<N>`` state dict (no network access here; the reference downloads it,
        ``resnet_encoder.py:71-83``): explicit path, ``MINE_RESNET<N>_WEIGHTS``, then the torch hub cache."""
        import glob
        import os
        cands = [path, os.environ.get("MINE_RESNET%d_WEIGHTS" % num_layers)]
        cands += sorted(glob.glob(os.path.expanduser("~/.cache/torch/hub/checkpoints/resnet%d-*.pth" % num_layers)))
        for c in cands:
            if c and os.path.exists(c):
                return c
        import warnings
        warnings.warn("model.imagenet_pretrained=true but no local resnet%d weights were found "
                      "(set MINE_RESNET%d_WEIGHTS); continuing with random initialisation" % (num_layers, num_layers))
        return None

    def _load_imagenet(self, path):
        found = self._find_imagenet(path, self.num_layers)
        self.pretrained_source = found                 # None: requested but not found -> random initialisation
        if found is not None:
            sd = torch.load(found, map_location="cpu")
            self.encoder.load_state_dict({k: v for k, v in sd.items() if not k.startswith("fc.")}, strict=True)

    def forward(self, img: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        e = self.encoder
        x = (img - self.img_mean.to(img.dtype)) / self.img_std.to(img.dtype)
        c1 = F.relu(e.bn1(e.conv1(x)))
        b1 = e.layer1(F.max_pool2d(c1, 3, 2, 1))
        b2 = e.layer2(b1)
        b3 = e.layer3(b2)
        b4 = e.layer4(b3)
        return c1, b1, b2, b3, b4
