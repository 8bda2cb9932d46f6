from mine_b200.models.decoder import DepthDecoder  # noqa: F401
