"""python -m hetu.rpc.local_start --ngpus 8 -- python3 train.py ...   (ref: python/hetu/rpc/local_start.py)"""
import runpy

if __name__ == "__main__":
    runpy.run_module("hetu_b200.rpc.launcher", run_name="__main__")
