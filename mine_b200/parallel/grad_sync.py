"""Data-parallel gradient synchronisation over a flat gradient arena.

Reference behaviour being replaced: two ``DistributedDataParallel`` wrappers with
``find_unused_parameters=True`` and ``broadcast_buffers=True`` (``synthesis_task.py:106-113``):
~6 bucketed NCCL all-reduces of 145 MiB fp32 per step, a used-parameter bitmap all-reduce, an
autograd graph walk, and a BN-buffer broadcast before every forward.

Here: parameters and gradients live in two flat fp32 arenas (:class:`FlatArena`); the graph is
static (no unused parameters: the ``fc`` head does not exist) so there is no bitmap exchange; BN
running statistics are bit-identical on all ranks by construction (they are computed from the
all-reduced batch statistics) so there is no buffer broadcast.  Buckets are contiguous arena
slices in *reverse registration order* (the order backward produces gradients); a
post-accumulate-grad hook counts arrivals and, when a bucket is complete, launches the mean
all-reduce for that slice on a side stream, overlapping the rest of backward.  ``finish()`` joins
the side stream before the optimizer runs.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .comm import Communicator


class FlatArena:
    """Re-homes the given parameters (and their ``.grad``) into contiguous fp32 buffers."""

    def __init__(self, params: Sequence[torch.nn.Parameter], align: int = 64, grad_alloc=None,
                 channels_last: bool = True):
        self.params = [p for p in params]
        device = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + align - 1) // align * align
        self.numel = off
        self.data = torch.zeros(off, dtype=torch.float32, device=device)
        # ``grad_alloc(numel)`` lets the communicator place the gradient arena on its symmetric heap so
        # the all-reduce runs in place over NVLink peer mappings
        self.grad = grad_alloc(off) if grad_alloc is not None else torch.zeros(off, dtype=torch.float32, device=device)
        # 4-D (convolution) weights are stored channels-last inside the arena: cuDNN then consumes them
        # without per-step NCHW<->NHWC conversion kernels, and the tcgen05 packers read contiguous Ci.
        self.channels_last = [bool(channels_last and p.dim() == 4 and p.is_cuda) for p in self.params]
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            view = self.view_of(self.data, i)
            view.copy_(p.data.float())
            p.data = view
            p.grad = self.view_of(self.grad, i)

    def view_of(self, flat: torch.Tensor, i: int) -> torch.Tensor:
        """Logical-shape view of parameter ``i`` inside a flat buffer laid out like this arena."""
        p, o = self.params[i], self.offsets[i]
        chunk = flat[o:o + p.numel()]
        if self.channels_last[i]:
            co, ci, kh, kw = p.shape
            return chunk.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return chunk.view(p.shape)

    # ---- gradients -----------------------------------------------------------------------------------------------
    # Two modes.  "accumulate" (CPU, and anything that calls backward() more than once per step): ``.grad`` of every
    # parameter is a view of the gradient arena, autograd adds into it (one elementwise kernel per parameter).
    # "gather" (CUDA default, ``MINE_B200_GRAD_GATHER=0`` disables): ``.grad`` is None during backward, so autograd's
    # AccumulateGrad just keeps the incoming tensor (no kernel); :meth:`gather_grads` then moves whole parameter ranges
    # into the arena with a few multi-tensor launches (``adam.cu::multi_copy_kernel``) and re-attaches the views.
    def gather_mode(self) -> bool:
        import os
        return self.data.is_cuda and os.environ.get("MINE_B200_GRAD_GATHER", "1") == "1"

    def zero_grad(self) -> None:
        if self.gather_mode():
            for p in self.params:
                p.grad = None
            self._gathered = [False] * len(self.params)
            return
        self.grad.zero_()
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):   # re-attach if something set grads to None
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.view_of(self.grad, i)

    def gather_grads(self, first: int = 0, last: int = -1) -> None:
        """Move the autograd-owned gradients of parameters ``first..last`` into their arena slices (zero where a
        parameter received none) and point ``.grad`` at the arena again.  No-op outside gather mode."""
        if not self.gather_mode():
            return
        if last < 0:
            last = len(self.params) - 1
        done = getattr(self, "_gathered", None)
        if done is None:
            done = self._gathered = [False] * len(self.params)
        srcs, offs, nums, slow = [], [], [], []
        base = self.grad.data_ptr()
        for i in range(first, last + 1):
            if done[i]:
                continue
            done[i] = True
            p, o = self.params[i], self.offsets[i]
            g = p.grad
            if g is not None and g.data_ptr() == base + 4 * o:
                continue                                       # already lives in the arena
            if g is not None and (g.dtype != torch.float32 or g.stride() != p.stride() or not g.is_cuda):
                slow.append((i, g))                            # unexpected layout: plain copy
                continue
            srcs.append(g)
            offs.append(o)
            nums.append(p.numel())
        if srcs:
            from ..ops import cuda as C
            C._ext.gather_into_arena(self.grad, srcs, offs, nums)
            C.LAUNCHES["count"] += max(1, len(srcs) // 40)
        for i, g in slow:
            self.view_of(self.grad, i).copy_(g)
        for i in range(first, last + 1):
            self.params[i].grad = self.view_of(self.grad, i)

    def slice_of(self, first: int, last: int):
        """Arena range covering params ``first..last`` inclusive."""
        lo = self.offsets[first]
        hi = self.offsets[last] + self.params[last].numel()
        return lo, hi


class GradSync:
    def __init__(self, arena: FlatArena, comm: Communicator, bucket_bytes: int = 32 << 20, overlap: bool = True):
        self.arena, self.comm = arena, comm
        self.enabled = comm.world_size > 1
        self.overlap = overlap and arena.data.is_cuda
        self.stream = torch.cuda.Stream(arena.data.device) if (self.enabled and arena.data.is_cuda) else None
        # buckets: walk parameters in reverse registration order, close a bucket at ~bucket_bytes
        self.buckets: List[dict] = []
        hi_idx, acc = len(arena.params) - 1, 0
        for i in range(len(arena.params) - 1, -1, -1):
            acc += arena.params[i].numel() * 4
            if acc >= bucket_bytes or i == 0:
                lo, hi = arena.slice_of(i, hi_idx)
                self.buckets.append({"first": i, "last": hi_idx, "lo": lo, "hi": hi, "pending": 0,
                                     "count": hi_idx - i + 1})
                hi_idx, acc = i - 1, 0
        self._bucket_of = {}
        for b_i, b in enumerate(self.buckets):
            for i in range(b["first"], b["last"] + 1):
                self._bucket_of[i] = b_i
        self._handles = []
        if self.enabled and self.overlap:
            for i, p in enumerate(arena.params):
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
        self.launched = 0

    def _make_hook(self, i: int):
        def hook(param):
            b = self.buckets[self._bucket_of[i]]
            b["pending"] += 1
            if b["pending"] == b["count"]:
                self._launch(b)
        return hook

    def _launch(self, b: dict) -> None:
        b["pending"] = 0
        b["done"] = True
        self.arena.gather_grads(b["first"], b["last"])          # gather mode: this bucket's gradients into the arena
        view = self.arena.grad[b["lo"]:b["hi"]]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(view.device))
            self.comm.allreduce_mean_(view, stream=self.stream)
        else:
            self.comm.allreduce_mean_(view)
        self.launched += 1

    def begin_step(self) -> None:
        for b in self.buckets:
            b["pending"], b["done"] = 0, False

    def finish(self) -> None:
        """Reduce whatever was not launched from hooks, then make the compute stream wait."""
        if not self.enabled:
            self.arena.gather_grads()                           # single replica: everything at once
            return
        for b in self.buckets:
            if not b.get("done", False):
                self._launch(b)
        if self.stream is not None:
            torch.cuda.current_stream(self.arena.grad.device).wait_stream(self.stream)

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
