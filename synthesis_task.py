"""Public entry point with the upstream module name: ``from synthesis_task import SynthesisTask``."""
from mine_b200.task import SynthesisTask, _get_disparity_list  # noqa: F401

__all__ = ["SynthesisTask"]
