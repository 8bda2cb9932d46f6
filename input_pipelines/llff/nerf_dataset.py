from mine_b200.data.llff import NeRFDataset, _collate_fn  # noqa: F401
