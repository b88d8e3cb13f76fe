"""python -m hetu.rpc.pssh_start_elastic ...: elastic controller + worker launch (ref: python/hetu/rpc/pssh_start_elastic.py)"""
from .elastic_server import main  # noqa: F401

if __name__ == "__main__":
    main()
