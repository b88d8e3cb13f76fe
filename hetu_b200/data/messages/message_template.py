"""chat message templates (ref: python/hetu/data/messages/message_template.py)"""
from . import ChatTemplate, build_chat_sample  # noqa: F401
