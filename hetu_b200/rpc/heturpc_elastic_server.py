"""python -m hetu.rpc.heturpc_elastic_server ...   (ref: python/hetu/rpc/heturpc_elastic_server.py)"""
from .elastic_server import ElasticServer, ElasticStrategy, elastic_arg_parser, main  # noqa: F401

if __name__ == "__main__":
    main()
