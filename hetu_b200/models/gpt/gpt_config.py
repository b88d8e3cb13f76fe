from .gpt_model import GPTConfig  # noqa: F401
