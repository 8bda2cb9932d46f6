"""Rank-sharded sampling and a pinned, double-buffered host->device input pipeline.

The reference copies pageable host tensors into persistent GPU buffers synchronously every step
(``synthesis_task.py:187-196``) and then calls ``torch.cuda.synchronize()``.  ``DevicePrefetcher``
stages batch i+1 into pinned memory and issues its H2D copies on a dedicated copy stream while
step i computes; the consumer only waits on an event.
"""
from __future__ import annotations

import math
from typing import Iterable

import torch
import torch.utils.data as data


class ShardedSampler(data.Sampler):
    """Same contract as ``DistributedSampler`` (rank-strided shard of a seeded permutation,
    padded to equal length, ``set_epoch``), reference ``train.py:83``."""

    def __init__(self, dataset, world_size: int = 1, rank: int = 0, shuffle: bool = True, seed: int = 0,
                 drop_last: bool = False):
        self.n, self.world_size, self.rank, self.shuffle, self.seed, self.drop_last = len(dataset), world_size, rank, shuffle, seed, drop_last
        self.epoch = 0
        self.start = 0                      # one-shot offset into this rank's shard (mid-epoch resume)
        self.num_samples = self.n // world_size if drop_last else math.ceil(self.n / world_size)
        self.total = self.num_samples * world_size

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def set_start(self, consumed: int):
        """Skip the first ``consumed`` samples of this rank's shard in the NEXT iteration only."""
        self.start = max(0, int(consumed))

    def __len__(self):
        return self.num_samples

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = torch.randperm(self.n, generator=g).tolist()
        else:
            idx = list(range(self.n))
        if self.drop_last:
            idx = idx[:self.total]
        else:
            idx = (idx * math.ceil(self.total / max(len(idx), 1)))[:self.total]
        shard = idx[self.rank:self.total:self.world_size]
        start, self.start = self.start, 0
        return iter(shard[start:])


def _pin(tree):
    if torch.is_tensor(tree):
        return tree.pin_memory() if not tree.is_pinned() else tree
    if isinstance(tree, dict):
        return {k: _pin(v) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_pin(v) for v in tree)
    return tree


def _to_device(tree, device, non_blocking=True):
    if torch.is_tensor(tree):
        return tree.to(device, non_blocking=non_blocking)
    if isinstance(tree, dict):
        return {k: _to_device(v, device, non_blocking) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_to_device(v, device, non_blocking) for v in tree)
    return tree


def tree_bytes(tree) -> int:
    if torch.is_tensor(tree):
        return tree.numel() * tree.element_size()
    if isinstance(tree, dict):
        return sum(tree_bytes(v) for v in tree.values())
    if isinstance(tree, (tuple, list)):
        return sum(tree_bytes(v) for v in tree)
    return 0


class DevicePrefetcher:
    """Wraps a batch iterable; yields device-resident batches, copy of batch i+1 overlapped."""

    def __init__(self, loader: Iterable, device: torch.device):
        self.loader, self.device = loader, torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.sampler = getattr(loader, "sampler", None)

    def __len__(self):
        return len(self.loader)

    def _stage(self, it):
        try:
            batch = next(it)
        except StopIteration:
            return None
        if not self.cuda:
            return batch, None
        batch = _pin(batch)
        with torch.cuda.stream(self.stream):
            dev = _to_device(batch, self.device)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev, batch          # keep the pinned source alive until consumed

    def __iter__(self):
        it = iter(self.loader)
        nxt = self._stage(it)
        while nxt is not None:
            cur = nxt
            nxt = self._stage(it)
            if self.cuda:
                torch.cuda.current_stream(self.device).wait_event(cur[1])
            yield cur[0]
