"""(ref: python/hetu/utils/data/{dataloader,dataset}.py)"""
from ...data.dataloader import DataLoader, build_data_loader  # noqa: F401
from ...data.dataset import IndexedTokenDataset, JsonDataset, SyntheticDataset  # noqa: F401
