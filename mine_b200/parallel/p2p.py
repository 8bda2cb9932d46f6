"""NVLink peer-memory communicator: our all-reduce kernels on a symmetric heap.

``torch.distributed._symmetric_memory`` is used for what it is good at - allocating buffers that every
rank can map and exchanging the handles (rendezvous) - and nothing else: the reductions are the
hand-written kernels in ``ops/csrc/comm.cu`` (one-shot flag-synchronised SUM for BatchNorm statistics,
two-shot in-place mean for gradient buckets, optional NVSwitch multicast ``multimem`` path).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .comm import Communicator

SMALL_CAP = 64 * 1024            # floats per one-shot slot
LL_CAP = 8 * 1024                # floats per source in the low-latency (data+flag words) receive buffer
FLAG_CHANNELS = 72               # channel 0: one-shot; 1..: CTAs of the two-shot kernel
TWO_SHOT_BLOCKS = 64             # CTAs of the gradient kernel (leaves SMs for the overlapped backward)


class P2PComm(Communicator):
    name = "p2p-nvlink (own kernels)"

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from ..ops import cuda as C
        self._ext = C._ext
        self._symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device)
        # NVSwitch in-fabric reduction (multimem.*) pays off from 4 ranks up; below that plain P2P loads win
        # (measured at 2 ranks, 152 MB: P2P 0.26 ms, multimem 0.41 ms, NCCL 0.38 ms)
        mm = os.environ.get("MINE_B200_MULTIMEM", "auto")
        self.use_multimem = (dist.get_world_size(self.group) >= 4) if mm == "auto" else (mm == "1")
        self._small = self._alloc(2 * SMALL_CAP, torch.float32)
        self.use_ll = os.environ.get("MINE_B200_LL", "1") == "1"
        self._ll = self._alloc(2 * self.world_size * LL_CAP * 2, torch.int32)      # uint2 words
        self._ll["tensor"].zero_()
        self._epoch_ll = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=self.device)       # fused exchanges (ll_exchange.cuh)
        self._flags = self._alloc(FLAG_CHANNELS * 16, torch.int32)
        self._flags["tensor"].zero_()
        self._small["tensor"].zero_()
        # epochs are device-resident and advanced by the kernels themselves -> CUDA-graph replay safe
        self._epoch_small = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._epoch_big = torch.zeros(TWO_SHOT_BLOCKS, dtype=torch.int32, device=self.device)
        self.graph_safe = True
        self._arena = None
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    # ---- symmetric allocation -------------------------------------------------------------------
    def _alloc(self, numel: int, dtype) -> dict:
        symm = self._symm
        t = symm.empty(numel, dtype=dtype, device=self.device)
        hdl = symm.rendezvous(t, self.group.group_name if hasattr(self.group, "group_name") else self.group)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        return {"tensor": t, "handle": hdl, "ptrs": ptrs, "mc": mc}

    def alloc_symmetric(self, numel: int) -> torch.Tensor:
        """fp32 buffer mapped by every rank (used for the flat gradient arena -> in-place all-reduce)."""
        self._arena = self._alloc(numel, torch.float32)
        self._arena["tensor"].zero_()
        return self._arena["tensor"]

    # ---- collectives ----------------------------------------------------------------------------
    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if t.numel() > SMALL_CAP or t.dtype != torch.float32:
            dist.all_reduce(t, group=self.group)                  # cold path (never hit by BN statistics)
            return t
        if self.use_ll and t.numel() <= LL_CAP:
            self._ext.allreduce_small_ll(t, self._ll["ptrs"], self.rank, LL_CAP, self._epoch_ll)
        else:
            self._ext.allreduce_small(t, self._small["ptrs"], self._flags["ptrs"], self.rank, SMALL_CAP, self._epoch_small)
        from ..ops import cuda as C
        C.LAUNCHES["count"] += 1
        return t

    def fused_handle(self):
        """Arguments of the ``*_x`` kernels that run the low-latency statistic exchange in their own prologue instead of a
        separate all-reduce launch (``csrc/ll_exchange.cuh``): peer receive buffers, rank, slot capacity, the epoch
        counter shared with :meth:`allreduce_sum_`, and the CTA ticket counter.  ``None``: not available."""
        if not self.use_ll or os.environ.get("MINE_B200_FUSED_BN", "1") != "1":
            return None
        return (self._ll["ptrs"], self.rank, LL_CAP, self._epoch_ll, self._ticket)

    def allreduce_mean_(self, t: torch.Tensor, stream=None) -> torch.Tensor:
        a = self._arena
        if a is None or t.dtype != torch.float32:
            raise RuntimeError("allreduce_mean_ expects a slice of the symmetric gradient arena")
        off = (t.data_ptr() - a["tensor"].data_ptr()) // 4
        if off < 0 or off + t.numel() > a["tensor"].numel():
            raise RuntimeError("tensor is not inside the symmetric gradient arena")
        lo, hi = off, off + t.numel()
        if lo % 4 or hi % 4:
            raise RuntimeError("bucket bounds must be 16-byte aligned")
        mc = a["mc"] if self.use_multimem else 0
        ctx = torch.cuda.stream(stream) if stream is not None else _Null()
        with ctx:
            self._ext.allreduce_mean(a["ptrs"], self._flags["ptrs"], mc, lo, hi, self.rank, self._epoch_big,
                                     TWO_SHOT_BLOCKS)
        from ..ops import cuda as C
        C.LAUNCHES["count"] += 1
        return t

    def barrier(self) -> None:
        dist.barrier(self.group)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
