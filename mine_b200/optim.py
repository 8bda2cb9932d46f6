"""Adam over a flat parameter arena + MultiStepLR.

Reference: ``torch.optim.Adam(params=[backbone group, decoder group], weight_decay=4e-5)`` with
``MultiStepLR`` stepped once per epoch (``synthesis_task.py:83-87,116-118,666``).  In torch 1.8
that is a per-tensor Python loop (221 tensors x ~8 launches).  Here all parameters, gradients and
both moment buffers are contiguous fp32 arenas, so one step is ONE fused multi-tensor kernel
(``mine_b200/ops/csrc/adam.cu``) per learning-rate group - or a handful of flat torch ops on
CPU.  ``state_dict()`` / ``load_state_dict()`` speak ``torch.optim.Adam``'s format (per-parameter
``step / exp_avg / exp_avg_sq``) so optimizer state is interchangeable with reference checkpoints.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Sequence

import torch

from .parallel.grad_sync import FlatArena


class ArenaAdam:
    def __init__(self, arena: FlatArena, group_sizes: Sequence[int], lrs: Sequence[float],
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        assert sum(group_sizes) == len(arena.params)
        self.arena = arena
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(arena.data)
        self.exp_avg_sq = torch.zeros_like(arena.data)
        self.step_count = 0
        self._hyper = None                       # per group device tensor [lr, step] (CUDA path)
        self.param_groups: List[Dict] = []
        first = 0
        for n, lr in zip(group_sizes, lrs):
            lo, hi = arena.slice_of(first, first + n - 1)
            if first + n < len(arena.params):
                hi = arena.offsets[first + n]              # include alignment padding
            self.param_groups.append({"lr": float(lr), "initial_lr": float(lr), "betas": betas, "eps": eps,
                                      "weight_decay": weight_decay, "amsgrad": False,
                                      "params": list(range(first, first + n)), "_range": (lo, hi)})
            first += n

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.arena.zero_grad()

    def sync_hyper(self) -> None:
        """Push (lr, step) of every group to the device.  Called outside captured regions (after a
        checkpoint restore or a scheduler step); inside a step only ``step += 1`` happens, on the device."""
        if not self.arena.data.is_cuda:
            return
        if self._hyper is None:
            self._hyper = [torch.zeros(2, dtype=torch.float32, device=self.arena.data.device) for _ in self.param_groups]
        for h, g in zip(self._hyper, self.param_groups):
            h.copy_(torch.tensor([g["lr"], float(self.step_count)], dtype=torch.float32), non_blocking=True)
            g["_lr_on_device"] = g["lr"]

    @torch.no_grad()
    def step(self) -> None:
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1 ** self.step_count
        bc2 = 1.0 - b2 ** self.step_count
        use_kernel = self.arena.data.is_cuda and os.environ.get("MINE_B200_FORCE_SPEC", "0") != "1"
        if use_kernel and (self._hyper is None or any(g.get("_lr_on_device") != g["lr"] for g in self.param_groups)):
            self.step_count -= 1
            self.sync_hyper()
            self.step_count += 1
        for gi, g in enumerate(self.param_groups):
            lo, hi = g["_range"]
            p, gr = self.arena.data[lo:hi], self.arena.grad[lo:hi]
            m, v = self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi]
            if use_kernel:
                from .ops import cuda as C
                h = self._hyper[gi]
                h[1:2].add_(1.0)                 # device-side step counter (graph-replay safe)
                C.fused_adam_(p, gr, m, v, h, b1, b2, self.eps, self.weight_decay)
                continue
            grad = gr.add(p, alpha=self.weight_decay) if self.weight_decay != 0 else gr
            m.mul_(b1).add_(grad, alpha=1 - b1)
            v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-g["lr"] / bc1)

    # ---- torch.optim.Adam-compatible (de)serialisation --------------------------------------
    def state_dict(self) -> Dict:
        state = {}
        if self.step_count > 0:
            for i in range(len(self.arena.params)):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.arena.view_of(self.exp_avg, i).detach().cpu().contiguous().clone(),
                            "exp_avg_sq": self.arena.view_of(self.exp_avg_sq, i).detach().cpu().contiguous().clone()}
        groups = [{k: v for k, v in g.items() if not k.startswith("_")} for g in self.param_groups]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: Dict) -> None:
        steps = []
        for i, st in sd.get("state", {}).items():
            i = int(i)
            shape = self.arena.params[i].shape
            self.arena.view_of(self.exp_avg, i).copy_(st["exp_avg"].reshape(shape))
            self.arena.view_of(self.exp_avg_sq, i).copy_(st["exp_avg_sq"].reshape(shape))
            steps.append(int(float(st["step"])))
        self.step_count = max(steps) if steps else 0
        for g, saved in zip(self.param_groups, sd.get("param_groups", [])):
            g["lr"] = float(saved.get("lr", g["lr"]))
            g["initial_lr"] = float(saved.get("initial_lr", g["initial_lr"]))
        self.sync_hyper()


class MultiStepLR:
    """lr = initial_lr * gamma^(#milestones <= epoch); ``step()`` once per epoch."""

    def __init__(self, optimizer, milestones: Sequence[int], gamma: float = 0.1, last_epoch: int = 0):
        self.optimizer, self.milestones, self.gamma = optimizer, sorted(int(m) for m in milestones), gamma
        self.last_epoch = last_epoch
        self._apply()

    def _apply(self):
        k = sum(1 for m in self.milestones if m <= self.last_epoch)
        for g in self.optimizer.param_groups:
            g["lr"] = g.get("initial_lr", g["lr"]) * (self.gamma ** k)
        if hasattr(self.optimizer, "sync_hyper"):
            self.optimizer.sync_hyper()

    def step(self):
        self.last_epoch += 1
        self._apply()

    def state_dict(self):
        return {"last_epoch": self.last_epoch}

    def load_state_dict(self, sd):
        self.last_epoch = int(sd["last_epoch"])
        self._apply()
