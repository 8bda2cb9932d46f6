"""Upstream path of the COLMAP model reader/writer (reference ``input_pipelines/colmap_utils.py``): implemented in
``mine_b200/data/colmap.py``."""
from mine_b200.data.colmap import *  # noqa: F401,F403
from mine_b200.data.colmap import qvec2rotmat, read_model, rotmat2qvec, write_model  # noqa: F401
