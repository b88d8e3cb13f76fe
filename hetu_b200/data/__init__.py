from .dataset import JsonDataset, SyntheticDataset, IndexedTokenDataset  # noqa: F401
from .bucket import Bucket, pack_sequences, pad_sequences, generate_cp_pack_data, get_sorted_batch_and_len  # noqa: F401
from .dataloader import DataLoader, build_data_loader, parallel_data_provider  # noqa: F401
from .tokenizers import build_tokenizer, ByteTokenizer  # noqa: F401
from .messages import ChatTemplate, PromptTemplate, build_chat_sample  # noqa: F401
from .indexed import (IndexedDatasetBuilder, open_indexed, preprocess_jsonl, GPTSampleDataset, BlendedDataset, blending_indices,  # noqa: F401,E402
                      split_documents)
