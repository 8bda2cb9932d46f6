"""Upstream path of the (unused) view-dependent radiance predictor (reference
``network/monodepth2/view_dependent_radiance_predictor.py:16-45``): ``mine_b200/models/geometry_layers.py``."""
from mine_b200.models.geometry_layers import VDRPredictor  # noqa: F401
