"""(ref: python/hetu/rpc/pssh_workers.py)"""
from .launcher import local_start, pssh_start, read_hosts_yaml  # noqa: F401
