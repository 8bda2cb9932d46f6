from mine_b200.models.encoder import ResnetEncoder  # noqa: F401
