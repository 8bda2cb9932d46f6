"""Upstream path of the LLFF dataset (reference ``input_pipelines/llff/nerf_dataset.py:15-234``): ``mine_b200/data/llff.py``."""
from mine_b200.data.llff import NeRFDataset, _collate_fn  # noqa: F401
