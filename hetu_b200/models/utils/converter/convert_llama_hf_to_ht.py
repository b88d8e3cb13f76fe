"""HuggingFace Llama state dict -> this framework's parameter names and fused layouts
(ref: python/hetu/models/utils/converter/convert_llama_hf_to_ht.py)"""
from ....utils.checkpoint.legacy import convert_llama_hf_to_ht  # noqa: F401
