"""python -m hetu.rpc.heturpc_polling_server --world-size N --port P   (ref: python/hetu/rpc/heturpc_polling_server.py)"""
import argparse

from .server import DeviceControllerServer, serve  # noqa: F401


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--world-size", "-n", type=int, required=True)
    ap.add_argument("--port", type=int, default=23457)
    ap.add_argument("--host", default="0.0.0.0")
    a = ap.parse_args(argv)
    serve(a.world_size, a.host, a.port)


if __name__ == "__main__":
    main()
