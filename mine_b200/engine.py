"""Model execution engine: how encoder + decoder run on the current device.

* CPU                      -> the PyTorch modules in fp32 (plumbing path / oracle)
* CUDA, ``engine.conv=tcgen05`` (default once built) -> the sm_100a conv engine
  (``mine_b200/ops/conv_engine.py``): NHWC bf16 implicit-GEMM convolutions on tcgen05/TMEM fed by TMA,
  BN statistics from the conv epilogue, fused BN-apply/ELU/upsample/reflect-pad producers.
* CUDA, ``engine.conv=cudnn`` -> library baseline for A/B runs (bf16 autocast, channels_last).

All executors return the four MPIs packed as ``[B,S,H,W,4]`` fp32.
"""
from __future__ import annotations

import os
from typing import List

import torch

from .ops import api as ops


class ModelRunner:
    def __init__(self, backbone, decoder, config, device):
        self.backbone, self.decoder, self.config = backbone, decoder, config
        self.device = torch.device(device)
        mode = os.environ.get("MINE_B200_CONV", config.get("engine.conv", "auto"))
        if self.device.type != "cuda":
            # off-GPU the modules run as plain PyTorch - unless the kernel specification has been switched on
            # (tests / analysis tools), in which case the engine orchestration itself runs through ops/emu.py
            from .ops import conv_engine
            mode = "tcgen05" if (mode == "tcgen05" and conv_engine._emulated) else "spec"
        elif mode == "auto":
            from .ops import conv_engine
            mode = "tcgen05" if conv_engine.AVAILABLE else "cudnn"
        self.mode = mode
        self._engine = None
        if mode == "tcgen05":
            from .ops.conv_engine import ConvEngine
            self._engine = ConvEngine(backbone, decoder, config, self.device)

    def predict(self, src_imgs: torch.Tensor, disparity: torch.Tensor) -> List[torch.Tensor]:
        if self.mode == "tcgen05":
            return self._engine.predict(src_imgs, disparity)
        if self.mode == "cudnn_fp32":          # library convs in fp32 (numerics tests)
            feats = self.backbone(src_imgs)
            out = self.decoder(feats, disparity)
            return [ops.pack_mpi(out[("disp", s)]).contiguous() for s in range(4)]
        if self.mode == "cudnn":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                feats = self.backbone(src_imgs.contiguous(memory_format=torch.channels_last))
                out = self.decoder(feats, disparity)
            return [ops.pack_mpi(out[("disp", s)].float()).contiguous() for s in range(4)]
        feats = self.backbone(src_imgs)
        out = self.decoder(feats, disparity)
        return [ops.pack_mpi(out[("disp", s)]) for s in range(4)]
