"""Process bootstrap: one process per GPU, ``torch.distributed`` for rendezvous only.

Reference: ``train.py:59-63,152`` (``CUDA_VISIBLE_DEVICES`` = one GPU, ``init_process_group("nccl")``,
one ``dist.barrier()``).  Here the device is selected with ``torch.cuda.set_device(LOCAL_RANK)``
(all GPUs stay visible, which peer-to-peer mapping over NVLink needs), NCCL is used on CUDA and
gloo on CPU, and a world of 1 needs no process group at all.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


@dataclass
class DistContext:
    rank: int = 0
    world_size: int = 1
    local_rank: int = 0
    device: torch.device = torch.device("cpu")
    backend: Optional[str] = None

    @property
    def is_main(self) -> bool:
        return self.rank == 0

    @property
    def distributed(self) -> bool:
        return self.world_size > 1


def init_distributed(local_rank: Optional[int] = None, backend: Optional[str] = None, device: Optional[str] = None,
                     timeout_s: int = 600) -> DistContext:
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and device != "cpu"
    if use_cuda:
        torch.cuda.set_device(local_rank % max(torch.cuda.device_count(), 1))
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
    if backend is None:
        backend = "nccl" if use_cuda else "gloo"
    if world > 1 or "MASTER_ADDR" in os.environ and os.environ.get("MINE_FORCE_PG", "0") == "1":
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = dev
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
            _make_patient_group()
    return DistContext(rank, world, local_rank, dev, backend if dist.is_initialized() else None)


_PATIENT = {"group": None}


def _make_patient_group() -> None:
    """A host-side (gloo) group with a 24 h timeout for waits that are not bounded by a training step: the ranks that do
    not evaluate wait for rank 0's validation pass there (round-1 advice: a RealEstate10K validation longer than the
    10-minute collective timeout of the main group aborted the job)."""
    try:
        _PATIENT["group"] = dist.new_group(backend="gloo", timeout=datetime.timedelta(hours=24))
    except Exception:                                    # gloo unavailable: fall back to the main group
        _PATIENT["group"] = None


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def patient_barrier() -> None:
    """Barrier for long host-side waits (evaluation on rank 0): gloo group with a 24 h timeout when available."""
    if dist.is_available() and dist.is_initialized():
        g = _PATIENT["group"]
        if g is not None:
            dist.barrier(group=g)
        else:
            dist.barrier()


def allreduce_host_sum(values):
    """SUM of a small list of Python floats over all ranks through the host-side group (evaluation meters)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64)
    g = _PATIENT["group"]
    if g is not None:
        dist.all_reduce(t, group=g)
    else:                                              # no gloo group: go through the default backend's device
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.all_reduce(t)
        t = t.cpu()
    return t.tolist()


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        _PATIENT["group"] = None
        dist.destroy_process_group()


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def broadcast_module_state(*modules: torch.nn.Module, src: int = 0) -> None:
    """One-off parameter + buffer broadcast from ``src`` (what the DDP constructor does, SURVEY N4:
    this is how the rank-0-only checkpoint restore reaches the other ranks).  Cold path -> NCCL/gloo."""
    if world_size() == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src)


def any_rank(flag: bool, device: Optional[torch.device] = None) -> bool:
    """True on every rank if ``flag`` is true on at least one (MAX all-reduce of one integer; cold path)."""
    if world_size() == 1:
        return bool(flag)
    dev = device if (device is not None and dist.get_backend() == "nccl") else torch.device("cpu")
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item())


def broadcast_object(obj, src: int = 0):
    if world_size() == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]
