from .gpt_model import GPTConfig, GPTLMHeadModel, GPTModel  # noqa: F401
