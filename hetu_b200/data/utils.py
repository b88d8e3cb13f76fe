"""batch helpers (ref: python/hetu/data/utils.py)"""
from .bucket import generate_cp_pack_data, get_sorted_batch_and_len, pack_sequences, pad_sequences  # noqa: F401
from .dataloader import build_data_loader, parallel_data_provider  # noqa: F401
