from mine_b200.models.geometry_layers import VDRPredictor  # noqa: F401
