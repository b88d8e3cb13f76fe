"""Standalone KV store + producer/consumer queue used between a planner process and the trainers (Hydraulis)
(ref: python/hetu/rpc/kv_store/{server,client,producer_consumer,const}.py)."""
from __future__ import annotations

import time
from typing import Any, Optional

from ..client import DeviceClient
from ..server import DeviceControllerServer

TIMEOUT = 60.0           # seconds a blocking get waits for a key (ref: python/hetu/rpc/kv_store/const.py)
DEFAULT_PORT = 23460


class KeyValueStoreServer(DeviceControllerServer):
    def __init__(self, host: str = "127.0.0.1", port: int = 23458):
        super().__init__(world_size=0, host=host, port=port)


class KeyValueStoreClient:
    def __init__(self, address: str = "127.0.0.1:23458"):
        self.c = DeviceClient(address)

    def put(self, key: str, value: Any):
        self.c.put_json(key, value)

    def get(self, key: str, timeout: Optional[float] = None) -> Any:
        return self.c.call("Get", key=key, kind="json", **({"timeout": timeout} if timeout else {}))

    def remove(self, key: str) -> bool:
        return self.c.remove(key)

    def register_dict(self, name: str) -> "RemoteDict":
        return RemoteDict(self, name)


class RemoteDict:
    def __init__(self, client: KeyValueStoreClient, name: str):
        self.client, self.name = client, name

    def __setitem__(self, k, v):
        self.client.put(f"{self.name}/{k}", v)

    def __getitem__(self, k):
        return self.client.get(f"{self.name}/{k}")


class ProducerConsumer:
    """ordered single-producer / multi-consumer stream of plans: item i is published under `<name>/<i>`; consumers block
    until their index appears (the trainer asks for the strategy of step i while the planner works ahead)."""

    def __init__(self, client: KeyValueStoreClient, name: str, max_ahead: int = 8):
        self.client, self.name, self.max_ahead = client, name, max_ahead
        self._produced = 0

    def produce(self, item: Any):
        while self._produced - self.consumed_upto() >= self.max_ahead:
            time.sleep(0.01)
        self.client.put(f"{self.name}/{self._produced}", item)
        self._produced += 1

    def consume(self, index: int, timeout: float = 300.0) -> Any:
        v = self.client.get(f"{self.name}/{index}", timeout=timeout)
        self.client.put(f"{self.name}/__consumed", max(index + 1, self.consumed_upto()))
        return v

    def consumed_upto(self) -> int:
        try:
            return int(self.client.get(f"{self.name}/__consumed", timeout=0.001))
        except Exception:   # noqa: BLE001
            return 0
