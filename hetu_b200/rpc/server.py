"""DeviceController server: rank assignment, hostname / device-info / communicator-id exchange, typed KV, barriers,
consistency checks, heartbeats and stop flags (ref: python/hetu/rpc/heturpc_polling_server.py, heturpc_async_server.py;
the 30 RPCs of protos/heturpc.proto)."""
from __future__ import annotations

import socketserver
import threading
import time
from typing import Any, Dict, Optional

from .protocol import recv_msg, send_msg


class _State:
    def __init__(self, world_size: int):
        self.world_size = world_size
        self.lock = threading.Condition()
        self.ranks: Dict[str, int] = {}          # connect order -> rank (keyed by client uuid)
        self.local_counts: Dict[str, int] = {}   # hostname -> next local device index
        self.hostnames: Dict[int, str] = {}
        self.device_infos: Dict[int, Any] = {}
        self.nccl_ids: Dict[str, str] = {}
        self.kv: Dict[str, Any] = {}
        self.barriers: Dict[str, Dict[str, int]] = {}
        self.consistent: Dict[str, Dict[int, Any]] = {}
        self.last_beat: Dict[int, float] = {}
        self.stop_flag = False
        self.exited: set = set()
        # elastic re-mapping (ref: heturpc_elastic_server.py host_name_to_idx / host_name_to_local_idx): when set, the k-th worker
        # that connects from host h becomes rank host_to_ranks[h][k] on local device host_to_local[h][k] -- the planner, not the
        # registration order, decides which physical GPU plays which rank
        self.host_to_ranks: Dict[str, list] = {}
        self.host_to_local: Dict[str, list] = {}
        self.local_device_of: Dict[int, int] = {}


class DeviceControllerServer:
    def __init__(self, world_size: int, host: str = "127.0.0.1", port: int = 23457, heartbeat_timeout: float = 10.0,
                 host_to_ranks: Optional[Dict[str, list]] = None, host_to_local: Optional[Dict[str, list]] = None):
        self.state = _State(world_size)
        if host_to_ranks:
            self.state.host_to_ranks = {h: list(v) for h, v in host_to_ranks.items()}
            self.state.host_to_local = {h: list(v) for h, v in (host_to_local or {}).items()}
        self.heartbeat_timeout = heartbeat_timeout
        outer = self

        class Handler(socketserver.BaseRequestHandler):
            def handle(self):
                try:
                    while True:
                        req = recv_msg(self.request)
                        send_msg(self.request, outer.dispatch(req))
                except (ConnectionError, OSError):
                    return

        class Server(socketserver.ThreadingTCPServer):
            allow_reuse_address = True
            daemon_threads = True

        self.server = Server((host, port), Handler)
        self.address = self.server.server_address
        self.thread: Optional[threading.Thread] = None

    # ------------------------------------------------------------------ lifecycle
    def start(self):
        self.thread = threading.Thread(target=self.server.serve_forever, daemon=True)
        self.thread.start()
        return self

    def shutdown(self):
        self.server.shutdown()
        self.server.server_close()

    # ------------------------------------------------------------------ RPCs
    def dispatch(self, req: dict) -> dict:
        fn = getattr(self, "rpc_" + req.get("method", ""), None)
        if fn is None:
            return {"ok": False, "error": f"unknown method {req.get('method')}"}
        try:
            return {"ok": True, "value": fn(**req.get("args", {}))}
        except Exception as e:   # noqa: BLE001 -- reported to the caller
            return {"ok": False, "error": f"{type(e).__name__}: {e}"}

    def _wait(self, pred, timeout=None):
        st = self.state
        end = None if timeout is None else time.time() + timeout
        while not pred():
            remaining = None if end is None else end - time.time()
            if remaining is not None and remaining <= 0:
                raise TimeoutError("rendezvous timed out")
            st.lock.wait(remaining if remaining is not None else 1.0)

    def rpc_Connect(self, client_id: str, hostname: str):
        st = self.state
        with st.lock:
            if client_id not in st.ranks:
                k = st.local_counts.get(hostname, 0)
                if st.host_to_ranks:
                    planned = st.host_to_ranks.get(hostname)
                    if planned is None or k >= len(planned):
                        raise RuntimeError(f"host {hostname} has no rank left in the elastic plan (worker #{k})")
                    st.ranks[client_id] = int(planned[k])
                    loc = st.host_to_local.get(hostname)
                    if loc is not None and k < len(loc):
                        st.local_device_of[int(planned[k])] = int(loc[k])
                else:
                    st.ranks[client_id] = len(st.ranks)
                r = st.ranks[client_id]
                st.hostnames[r] = hostname
                st.local_counts[hostname] = k + 1
                st.last_beat[r] = time.time()
                st.lock.notify_all()
            return st.ranks[client_id]

    def rpc_GetRank(self, client_id: str, timeout: float = 300.0):
        """blocks until the whole world has connected -> (rank, local device index, world size)"""
        st = self.state
        with st.lock:
            self._wait(lambda: len(st.ranks) >= st.world_size, timeout)
            r = st.ranks[client_id]
            host = st.hostnames[r]
            local = st.local_device_of.get(r)
            if local is None:
                local = sorted(k for k, h in st.hostnames.items() if h == host).index(r)
            return {"rank": r, "local_device": local, "world_size": st.world_size}

    def rpc_CommitHostName(self, rank: int, hostname: str):
        with self.state.lock:
            self.state.hostnames[rank] = hostname
            self.state.lock.notify_all()

    def rpc_GetHostName(self, rank: int, timeout: float = 300.0):
        st = self.state
        with st.lock:
            self._wait(lambda: rank in st.hostnames, timeout)
            return st.hostnames[rank]

    def rpc_CommitDeviceInfo(self, rank: int, info):
        with self.state.lock:
            self.state.device_infos[rank] = info
            self.state.lock.notify_all()

    def rpc_GetDeviceInfo(self, rank: int, timeout: float = 300.0):
        st = self.state
        with st.lock:
            self._wait(lambda: rank in st.device_infos, timeout)
            return st.device_infos[rank]

    def rpc_CommitNcclId(self, key: str, nccl_id: str):
        with self.state.lock:
            self.state.nccl_ids[key] = nccl_id
            self.state.lock.notify_all()

    def rpc_GetNcclId(self, key: str, timeout: float = 300.0):
        st = self.state
        with st.lock:
            self._wait(lambda: key in st.nccl_ids, timeout)
            return st.nccl_ids[key]

    def rpc_Put(self, key: str, value, kind: str = "json"):
        with self.state.lock:
            self.state.kv[f"{kind}:{key}"] = value
            self.state.lock.notify_all()

    def rpc_Get(self, key: str, kind: str = "json", timeout: float = 300.0):
        st = self.state
        with st.lock:
            self._wait(lambda: f"{kind}:{key}" in st.kv, timeout)
            return st.kv[f"{kind}:{key}"]

    def rpc_Remove(self, key: str, kind: str = "json"):
        with self.state.lock:
            return self.state.kv.pop(f"{kind}:{key}", None) is not None

    def rpc_Barrier(self, rank: int, world_ranks=None, tag: str = "", timeout: float = 600.0):
        st = self.state
        ranks = tuple(sorted(world_ranks)) if world_ranks else tuple(range(st.world_size))
        name = f"{tag}:{ranks}"
        with st.lock:
            b = st.barriers.setdefault(name, {"count": 0, "gen": 0})
            gen = b["gen"]
            b["count"] += 1
            if b["count"] == len(ranks):
                b["count"] = 0
                b["gen"] += 1
                st.lock.notify_all()
            else:
                self._wait(lambda: b["gen"] != gen, timeout)
            return True

    def rpc_Consistent(self, rank: int, value, world_ranks=None, tag: str = "", timeout: float = 600.0):
        """all ranks must present the same value (used to verify that every process built the same graph/strategy)"""
        st = self.state
        ranks = tuple(sorted(world_ranks)) if world_ranks else tuple(range(st.world_size))
        name = f"{tag}:{ranks}"
        with st.lock:
            d = st.consistent.setdefault(name, {})
            d[rank] = value
            st.lock.notify_all()
            self._wait(lambda: len(d) >= len(ranks), timeout)
            vals = list(d.values())
            return all(v == vals[0] for v in vals)

    def rpc_HeartBeat(self, rank: int):
        with self.state.lock:
            self.state.last_beat[rank] = time.time()
            return self.state.stop_flag

    def rpc_WorkerStop(self):
        with self.state.lock:
            self.state.stop_flag = True
            self.state.lock.notify_all()

    def rpc_AlreadyStop(self):
        return self.state.stop_flag

    def rpc_Exit(self, rank: int):
        with self.state.lock:
            self.state.exited.add(rank)
            self.state.lock.notify_all()
            return len(self.state.exited)

    # ------------------------------------------------------------------ monitoring
    def dead_ranks(self):
        now = time.time()
        with self.state.lock:
            return sorted(r for r, t in self.state.last_beat.items() if now - t > self.heartbeat_timeout and r not in self.state.exited)

    def all_exited(self):
        with self.state.lock:
            return len(self.state.exited) >= self.state.world_size


def serve(world_size: int, host: str = "0.0.0.0", port: int = 23457):
    """blocking entry point: `python -m hetu_b200.rpc.server --world N --port P`"""
    srv = DeviceControllerServer(world_size, host, port).start()
    try:
        while not srv.all_exited():
            time.sleep(0.5)
    finally:
        srv.shutdown()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=23457)
    a = ap.parse_args()
    serve(a.world, a.host, a.port)
