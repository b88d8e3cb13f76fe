"""(ref: python/hetu/rpc/elastic_arg_parser.py)"""
from .elastic_server import elastic_arg_parser  # noqa: F401
