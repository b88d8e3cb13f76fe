from .llama_model import LlamaConfig  # noqa: F401
