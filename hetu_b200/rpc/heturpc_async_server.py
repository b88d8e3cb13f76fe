"""(ref: python/hetu/rpc/heturpc_async_server.py) -- the controller handles every connection on its own thread, so the
polling and async variants of the reference are one implementation here"""
from .heturpc_polling_server import DeviceControllerServer, main, serve  # noqa: F401

if __name__ == "__main__":
    main()
