"""Upstream path of the disparity-conditioned decoder (reference ``network/monodepth2/depth_decoder.py:35-148``):
``mine_b200/models/decoder.py`` (engine: ``mine_b200/ops/conv_engine.py``)."""
from mine_b200.models.decoder import DepthDecoder  # noqa: F401
