"""prompt templates (ref: python/hetu/data/messages/prompt_template.py)"""
from . import PromptTemplate  # noqa: F401
