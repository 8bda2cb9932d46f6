from mine_b200.data.colmap import COLMAPDatabase, image_ids_to_pair_id, pair_id_to_image_ids  # noqa: F401
