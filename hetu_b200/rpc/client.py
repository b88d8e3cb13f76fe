"""Worker-side client of the DeviceController (ref: hetu/impl/communication/rpc_client.{h,cc}: Connect, GetRank,
CommitHostName/GetHostName, CommitDeviceInfo/GetDeviceInfo, CommitNcclId/GetNcclId, Put/Get{Double,Int,String,Bytes,Json},
Barrier, Consistent, HeartBeat thread, WorkerStop, Exit)."""
from __future__ import annotations

import atexit
import socket
import threading
import time
import uuid
from typing import Any, Optional, Sequence

from .protocol import b64, recv_msg, send_msg, unb64


class DeviceClient:
    def __init__(self, address: str = "127.0.0.1:23457", hostname: Optional[str] = None, heartbeat_interval: float = 2.0,
                 connect_timeout: float = 60.0):
        host, port = address.rsplit(":", 1)
        self.addr = (host, int(port))
        self.hostname = hostname or socket.gethostname()
        self.client_id = uuid.uuid4().hex
        self._lock = threading.Lock()
        self._sock = self._connect(connect_timeout)
        self.rank = self.local_device = self.world_size = None
        self._hb_stop = threading.Event()
        self._hb_thread: Optional[threading.Thread] = None
        self.heartbeat_interval = heartbeat_interval
        self.stop_requested = False

    def _connect(self, timeout):
        end = time.time() + timeout
        while True:
            try:
                s = socket.create_connection(self.addr, timeout=5)
                s.settimeout(None)
                return s
            except OSError:
                if time.time() > end:
                    raise
                time.sleep(0.2)

    def call(self, method: str, **args) -> Any:
        with self._lock:
            send_msg(self._sock, {"method": method, "args": args})
            rep = recv_msg(self._sock)
        if not rep["ok"]:
            raise RuntimeError(f"rpc {method} failed: {rep['error']}")
        return rep["value"]

    # ------------------------------------------------------------------ bootstrap
    def connect(self, start_heartbeat: bool = True):
        self.call("Connect", client_id=self.client_id, hostname=self.hostname)
        info = self.call("GetRank", client_id=self.client_id)
        self.rank, self.local_device, self.world_size = info["rank"], info["local_device"], info["world_size"]
        if start_heartbeat:
            self.launch_heartbeat()
        atexit.register(self._exit_quiet)
        return self.rank, self.local_device, self.world_size

    def launch_heartbeat(self):
        # a second connection so a blocking RPC (barrier) never delays the liveness signal
        hb_sock = self._connect(10)

        def loop():
            while not self._hb_stop.wait(self.heartbeat_interval):
                try:
                    send_msg(hb_sock, {"method": "HeartBeat", "args": {"rank": self.rank}})
                    rep = recv_msg(hb_sock)
                    if rep.get("ok") and rep.get("value"):
                        self.stop_requested = True
                except (ConnectionError, OSError):
                    return
        self._hb_thread = threading.Thread(target=loop, daemon=True)
        self._hb_thread.start()

    def all_gather_hostnames(self):
        self.call("CommitHostName", rank=self.rank, hostname=self.hostname)
        return [self.call("GetHostName", rank=r) for r in range(self.world_size)]

    def exchange_device_info(self, info):
        self.call("CommitDeviceInfo", rank=self.rank, info=info)
        return [self.call("GetDeviceInfo", rank=r) for r in range(self.world_size)]

    def commit_nccl_id(self, ranks: Sequence[int], stream: int, nccl_id: bytes):
        self.call("CommitNcclId", key=f"{sorted(ranks)}:{stream}", nccl_id=b64(nccl_id))

    def get_nccl_id(self, ranks: Sequence[int], stream: int) -> bytes:
        return unb64(self.call("GetNcclId", key=f"{sorted(ranks)}:{stream}"))

    # ------------------------------------------------------------------ typed KV
    def put_double(self, k, v): self.call("Put", key=k, value=float(v), kind="double")       # noqa: E704
    def get_double(self, k): return float(self.call("Get", key=k, kind="double"))             # noqa: E704
    def put_int(self, k, v): self.call("Put", key=k, value=int(v), kind="int")                 # noqa: E704
    def get_int(self, k): return int(self.call("Get", key=k, kind="int"))                      # noqa: E704
    def put_string(self, k, v): self.call("Put", key=k, value=str(v), kind="string")           # noqa: E704
    def get_string(self, k): return self.call("Get", key=k, kind="string")                     # noqa: E704
    def put_bytes(self, k, v: bytes): self.call("Put", key=k, value=b64(v), kind="bytes")      # noqa: E704
    def get_bytes(self, k) -> bytes: return unb64(self.call("Get", key=k, kind="bytes"))       # noqa: E704
    def put_json(self, k, v): self.call("Put", key=k, value=v, kind="json")                    # noqa: E704
    def get_json(self, k): return self.call("Get", key=k, kind="json")                         # noqa: E704
    def remove(self, k, kind="json"): return self.call("Remove", key=k, kind=kind)             # noqa: E704

    # ------------------------------------------------------------------ sync
    def barrier(self, ranks: Optional[Sequence[int]] = None, tag: str = ""):
        return self.call("Barrier", rank=self.rank, world_ranks=list(ranks) if ranks else None, tag=tag)

    def consistent(self, value, ranks: Optional[Sequence[int]] = None, tag: str = "") -> bool:
        return self.call("Consistent", rank=self.rank, value=value, world_ranks=list(ranks) if ranks else None, tag=tag)

    def worker_stop(self):
        self.call("WorkerStop")

    def already_stop(self) -> bool:
        return bool(self.call("AlreadyStop")) or self.stop_requested

    def exit(self):
        self._hb_stop.set()
        self.call("Exit", rank=self.rank)

    def _exit_quiet(self):
        try:
            if self.rank is not None:
                self.exit()
        except Exception:   # noqa: BLE001
            pass


class NativeDeviceClient:
    """Same surface as `DeviceClient`, transported by the C++ client (`csrc/runtime/rpc_client.cc`: sockets, framing, JSON
    scanning, base64 and the heart-beat thread are native; blocking calls release the GIL).  This is the client the runtime
    bootstrap uses (ref: hetu/impl/communication/rpc_client.cc DeviceClientImpl)."""

    def __init__(self, address: str = "127.0.0.1:23457", hostname: Optional[str] = None, heartbeat_interval: float = 2.0,
                 connect_timeout: float = 60.0):
        import json

        from .. import _C
        host, port = address.rsplit(":", 1)
        self._json = json
        self._c = _C.RpcClient(host, int(port), hostname or socket.gethostname(), float(heartbeat_interval), float(connect_timeout))
        self.hostname = hostname or socket.gethostname()
        self.rank = self.local_device = self.world_size = None

    @property
    def client_id(self):
        return self._c.client_id

    @property
    def heartbeats_sent(self):
        return self._c.heartbeats_sent

    def call(self, method: str, **args) -> Any:
        return self._json.loads(self._c.call(method, self._json.dumps(args)))

    def connect(self, start_heartbeat: bool = True):
        self._c.connect(start_heartbeat)
        self.rank, self.local_device, self.world_size = self._c.rank, self._c.local_device, self._c.world_size
        return self.rank, self.local_device, self.world_size

    def all_gather_hostnames(self):
        self._c.commit_hostname()
        return [self._c.get_hostname(r) for r in range(self.world_size)]

    def exchange_device_info(self, info):
        self.call("CommitDeviceInfo", rank=self.rank, info=info)
        return [self.call("GetDeviceInfo", rank=r) for r in range(self.world_size)]

    def commit_nccl_id(self, ranks, stream, nccl_id: bytes): self._c.commit_nccl_id(list(ranks), int(stream), bytes(nccl_id))   # noqa: E704
    def get_nccl_id(self, ranks, stream) -> bytes: return self._c.get_nccl_id(list(ranks), int(stream))                         # noqa: E704
    def put_double(self, k, v): self._c.put_double(k, float(v))                 # noqa: E704
    def get_double(self, k): return self._c.get_double(k)                        # noqa: E704
    def put_int(self, k, v): self._c.put_int(k, int(v))                          # noqa: E704
    def get_int(self, k): return self._c.get_int(k)                              # noqa: E704
    def put_string(self, k, v): self._c.put_string(k, str(v))                    # noqa: E704
    def get_string(self, k): return self._c.get_string(k)                        # noqa: E704
    def put_bytes(self, k, v: bytes): self._c.put_bytes(k, bytes(v))             # noqa: E704
    def get_bytes(self, k) -> bytes: return self._c.get_bytes(k)                 # noqa: E704
    def put_json(self, k, v): self._c.put_json(k, self._json.dumps(v))           # noqa: E704
    def get_json(self, k): return self._json.loads(self._c.get_json(k))          # noqa: E704
    def remove(self, k, kind="json"): return self._c.remove(k, kind)             # noqa: E704

    def barrier(self, ranks: Optional[Sequence[int]] = None, tag: str = ""):
        self._c.barrier(list(ranks) if ranks else [], tag)
        return True

    def consistent(self, value, ranks: Optional[Sequence[int]] = None, tag: str = "") -> bool:
        return self._c.consistent(self._json.dumps(value), list(ranks) if ranks else [], tag)

    def worker_stop(self): self._c.worker_stop()                                 # noqa: E704
    def already_stop(self) -> bool: return self._c.already_stop()                # noqa: E704
    def exit(self): self._c.exit()                                               # noqa: E704
