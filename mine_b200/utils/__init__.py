from .meters import AverageMeter, DeviceTimer, PhaseProfiler, nvtx_range
from .misc import (NullLogger, disparity_normalization_vis, linspace_batch, make_logger, run_shell_cmd,
                   run_shell_cmd_shell, seed_everything)

__all__ = ["AverageMeter", "DeviceTimer", "PhaseProfiler", "nvtx_range", "NullLogger",
           "disparity_normalization_vis", "linspace_batch", "make_logger", "run_shell_cmd",
           "run_shell_cmd_shell", "seed_everything"]
