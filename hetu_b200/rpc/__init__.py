"""Rendezvous / KV / heartbeat / elastic-restart services and launchers
(ref: python/hetu/rpc/** and hetu/impl/communication/rpc_client.{h,cc}, protos/heturpc.proto service DeviceController).
The reference speaks gRPC; here the same RPC surface is a length-prefixed JSON protocol over TCP (stdlib only), since the
data plane bootstraps through torch.distributed and this control plane only carries small metadata."""
from .server import DeviceControllerServer, serve  # noqa: F401
from .client import DeviceClient, NativeDeviceClient  # noqa: F401
from .elastic_server import ElasticServer, ElasticStrategy  # noqa: F401
from .kv_store import KeyValueStoreClient, KeyValueStoreServer, ProducerConsumer  # noqa: F401
from .launcher import local_start, pssh_start, read_hosts_yaml  # noqa: F401
