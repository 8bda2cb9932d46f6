"""Model execution engine: how encoder + decoder run on the current device.

* CPU                      -> the PyTorch modules in fp32 (plumbing path / oracle)
* CUDA, ``engine.conv=tcgen05`` (default once built) -> the sm_100a conv engine
  (``mine_b200/ops/conv_engine.py``): NHWC implicit-GEMM convolutions on tcgen05/TMEM fed by TMA,
  BN statistics from the conv epilogue, fused BN-apply/ELU/upsample/reflect-pad producers.
* CUDA, ``engine.conv=cudnn`` -> library baseline for A/B runs (channels_last).

``engine.precision`` (``MINE_B200_PRECISION`` when the key is absent): ``tf32`` (default) = fp32 tensors with TF32 tensor-core
convolutions and fp32 accumulation, the numerics class of the reference (fp32 + cuDNN TF32, never autocast);
``bf16`` = bf16 activations / operands with fp32 accumulation and fp32 master weights (the fast mode).

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
        self.precision = str(config.get("engine.precision") or os.environ.get("MINE_B200_PRECISION", "tf32"))
        if self.precision not in ("tf32", "bf16"):
            raise ValueError("engine.precision must be 'tf32' or 'bf16', got %r" % (self.precision,))
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
            from .ops import conv_engine
            if not conv_engine._emulated:
                conv_engine.set_precision(self.precision)
                conv_engine.set_deterministic(bool(config.get("engine.deterministic", False)) or
                                              os.environ.get("MINE_B200_DETERMINISTIC", "0") == "1")
            self._engine = conv_engine.ConvEngine(backbone, decoder, config, self.device)

    def predict(self, src_imgs: torch.Tensor, disparity: torch.Tensor) -> List[torch.Tensor]:
        if self.mode == "tcgen05":
            return self._engine.predict(src_imgs, disparity)
        if self.mode == "cudnn_fp32":          # library convs in fp32 (numerics tests)
            feats = self.backbone(src_imgs)
            out = self.decoder(feats, disparity)
            return [ops.pack_mpi(out[("disp", s)]).contiguous() for s in range(4)]
        if self.mode == "cudnn":
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.precision == "bf16"):
                feats = self.backbone(src_imgs.contiguous(memory_format=torch.channels_last))
                out = self.decoder(feats, disparity)
            return [ops.pack_mpi(out[("disp", s)].float()).contiguous() for s in range(4)]
        feats = self.backbone(src_imgs)
        out = self.decoder(feats, disparity)
        return [ops.pack_mpi(out[("disp", s)]) for s in range(4)]
