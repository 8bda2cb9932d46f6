"""Upstream path of the COLMAP sqlite wrapper (reference ``input_pipelines/database.py``): ``mine_b200/data/colmap.py``."""
from mine_b200.data.colmap import COLMAPDatabase, image_ids_to_pair_id, pair_id_to_image_ids  # noqa: F401
