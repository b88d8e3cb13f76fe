from ...data.dataset import IndexedTokenDataset, JsonDataset, SyntheticDataset  # noqa: F401
