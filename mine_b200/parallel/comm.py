"""Communicators: SUM all-reduce for small statistic vectors and for flat gradient buckets.

Two implementations behind one interface:

* :class:`TorchDistComm` - ``torch.distributed`` collectives (NCCL on GPU, gloo on CPU).  This is
  the *baseline* path (what the reference does through DDP/SyncBN) and the CPU test path.
* :class:`mine_b200.parallel.p2p.P2PComm` - our own sm_100a kernels over NVLink peer memory
  (one-shot all-reduce for <=64 KiB statistic vectors, two-shot reduce-scatter/all-gather fused
  with the 1/world scaling for gradient buckets).  Selected with ``engine.comm: p2p`` on CUDA.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class Communicator:
    world_size: int = 1
    rank: int = 0
    name = "local"

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place SUM over ranks on the current stream; returns ``t``."""
        return t

    def allreduce_mean_(self, t: torch.Tensor, stream=None) -> torch.Tensor:
        """In-place mean over ranks (gradient buckets): SUM fused with the 1/world scaling."""
        return t

    def barrier(self) -> None:
        pass

    def close(self) -> None:
        pass


class TorchDistComm(Communicator):
    name = "torch.distributed"

    def __init__(self, group=None):
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def allreduce_sum_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_mean_(self, t, stream=None):
        if stream is not None and t.is_cuda:
            with torch.cuda.stream(stream):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.mul_(1.0 / self.world_size)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.mul_(1.0 / self.world_size)
        return t

    def barrier(self):
        dist.barrier(group=self.group)


def single_node(world_size: int, env=None) -> bool:
    """True when every rank of the job lives on this node (``LOCAL_WORLD_SIZE`` / ``GROUP_WORLD_SIZE`` exported by
    ``torch.distributed.run``).  The peer-memory kernels map buffers over NVLink/NVSwitch, i.e. inside one node."""
    import os
    env = os.environ if env is None else env
    if int(env.get("GROUP_WORLD_SIZE", "1")) > 1:
        return False
    local = env.get("LOCAL_WORLD_SIZE")
    return local is None or int(local) >= world_size


def make_communicator(kind: str = "auto", device: Optional[torch.device] = None) -> Communicator:
    """``kind``: ``p2p`` (own NVLink kernels; CUDA only), ``nccl``/``torch`` (library baseline),
    ``auto`` (p2p on CUDA, torch.distributed otherwise).  Jobs that span several nodes use the library
    collectives for now (the peer-memory kernels are node-local; a hierarchical NVLink + network scheme is not
    implemented)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return Communicator()
    on_cuda = device is not None and torch.device(device).type == "cuda"
    if kind in ("p2p", "auto") and on_cuda and not single_node(dist.get_world_size()):
        import warnings
        warnings.warn("multi-node job: peer-memory (NVLink) collectives are node-local, using NCCL")
        return TorchDistComm()
    if kind in ("p2p", "auto") and on_cuda:
        try:
            from .p2p import P2PComm
            return P2PComm(device)
        except Exception as e:          # no symmetric memory on this box: say so loudly, use the library baseline
            import warnings
            warnings.warn(f"P2P communicator unavailable ({type(e).__name__}: {e}); falling back to NCCL")
    return TorchDistComm()
