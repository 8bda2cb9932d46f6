"""Small helpers: shell commands, disparity visualisation, logging setup, seeding."""
from __future__ import annotations

import logging
import os
import random
import subprocess
import sys
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


def run_shell_cmd(args_list: Sequence[str], logger=None, check: bool = False) -> Tuple[int, bytes, bytes]:
    """Run a command, return ``(returncode, stdout, stderr)`` (reference ``utils.run_shell_cmd``;
    unlike the reference a non-zero exit is logged, and raised when ``check``)."""
    if logger:
        logger.info("Running system command: {0}".format(" ".join(args_list)))
    try:
        proc = subprocess.run(list(args_list), capture_output=True)
    except FileNotFoundError as e:
        if logger:
            logger.warning("command not found: %s", e)
        if check:
            raise
        return 127, b"", str(e).encode()
    if proc.returncode != 0:
        if logger:
            logger.warning("command failed (%d): %s", proc.returncode, proc.stderr[-500:])
        if check:
            raise subprocess.CalledProcessError(proc.returncode, args_list, proc.stdout, proc.stderr)
    return proc.returncode, proc.stdout, proc.stderr


def run_shell_cmd_shell(cmd: str, logger=None):
    if logger:
        logger.info("Running system command: {0}".format(cmd))
    proc = subprocess.run(cmd, shell=True, capture_output=True)
    return proc.returncode, proc.stdout, proc.stderr


def disparity_normalization_vis(disparity: torch.Tensor) -> torch.Tensor:
    """Per-image min/max normalisation to [0,1] of a ``[B,1,H,W]`` disparity map."""
    if disparity.dim() != 4 or disparity.shape[1] != 1:
        raise ValueError("expected Bx1xHxW")
    lo = torch.amin(disparity, (1, 2, 3), keepdim=True)
    hi = torch.amax(disparity, (1, 2, 3), keepdim=True)
    return ((disparity - lo) / (hi - lo)).clamp(0.0, 1.0)


def linspace_batch(start: torch.Tensor, end: torch.Tensor, steps: int, dtype=None, device=None) -> torch.Tensor:
    """Row-wise linspace ``[B,steps]`` between per-row ``start`` and ``end``."""
    t = torch.linspace(0.0, 1.0, steps, dtype=dtype, device=device)
    start = start.to(dtype=dtype, device=device)
    end = end.to(dtype=dtype, device=device)
    return start[:, None] + (end - start)[:, None] * t[None]


def make_logger(name: str = "mine", log_file: Optional[str] = None, stdout: bool = True) -> logging.Logger:
    logger = logging.getLogger(name)
    fmt = logging.Formatter("[%(asctime)s %(filename)s] %(message)s")
    handlers = []
    if log_file:
        os.makedirs(os.path.dirname(os.path.abspath(log_file)) or ".", exist_ok=True)
        fh = logging.FileHandler(log_file)
        fh.setFormatter(fmt)
        handlers.append(fh)
    if stdout:
        sh = logging.StreamHandler(sys.stdout)
        sh.setFormatter(fmt)
        handlers.append(sh)
    logger.handlers = handlers
    logger.setLevel(logging.INFO)
    logger.propagate = False
    return logger


class NullLogger:
    """Stands in for ``logger=None`` on non-zero ranks so call sites need no guards."""

    def info(self, *a, **k):
        pass

    warning = error = debug = info


def seed_everything(seed: int, rank: int = 0) -> None:
    """The reference sets no seeds at all; we make runs reproducible per rank."""
    s = int(seed) * 1000003 + int(rank)
    random.seed(s)
    np.random.seed(s % (2 ** 32))
    torch.manual_seed(s)


def rng_state() -> dict:
    st = {"python": random.getstate(), "numpy": np.random.get_state(), "torch": torch.get_rng_state()}
    if torch.cuda.is_available():
        st["cuda"] = torch.cuda.get_rng_state()
    return st


def set_rng_state(st: dict) -> None:
    random.setstate(st["python"])
    np.random.set_state(st["numpy"])
    torch.set_rng_state(st["torch"])
    if "cuda" in st and torch.cuda.is_available():
        torch.cuda.set_rng_state(st["cuda"])


class StopRequest:
    """Cooperative shutdown on SIGTERM / SIGUSR1 (preemption, ``torchrun`` tearing the job down): the handler only
    sets a flag; the training loop polls it at a step boundary where all ranks agree, writes the resumable
    ``checkpoint_latest.pth`` and returns.  The reference has no failure handling at all."""

    def __init__(self, install: bool = True):
        self._flag = False
        self._previous = {}
        if install:
            self.install()

    def install(self) -> None:
        import signal
        import threading
        if threading.current_thread() is not threading.main_thread():
            return                                   # signal handlers can only be set from the main thread
        for sig in (signal.SIGTERM, signal.SIGUSR1):
            try:
                self._previous[sig] = signal.signal(sig, self._handle)
            except (ValueError, OSError):
                pass

    def uninstall(self) -> None:
        import signal
        for sig, prev in self._previous.items():
            try:
                signal.signal(sig, prev)
            except (ValueError, OSError):
                pass
        self._previous = {}

    def _handle(self, signum, frame) -> None:
        self._flag = True

    def set(self) -> None:
        self._flag = True

    def is_set(self) -> bool:
        return self._flag
