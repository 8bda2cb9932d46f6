"""Upstream path of the ResNet encoders (reference ``network/monodepth2/resnet_encoder.py:18-108``):
``mine_b200/models/encoder.py``."""
from mine_b200.models.encoder import (ResnetEncoder, ResNetMultiImageInput,  # noqa: F401
                                      resnet_multiimage_input)
