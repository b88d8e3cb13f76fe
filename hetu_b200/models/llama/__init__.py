from .llama_model import LlamaConfig, LlamaLMHeadModel, LlamaModel  # noqa: F401
