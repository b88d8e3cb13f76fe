"""python -m hetu.rpc.pssh_start --hosts hosts.yaml --port 29500 -- python3 train.py ...   (ref: python/hetu/rpc/pssh_start.py)"""
import runpy

if __name__ == "__main__":
    runpy.run_module("hetu_b200.rpc.launcher", run_name="__main__")
