from mine_b200.models.encoder import (ResnetEncoder, ResNetMultiImageInput,  # noqa: F401
                                      resnet_multiimage_input)
