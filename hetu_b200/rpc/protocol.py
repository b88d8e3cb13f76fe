from __future__ import annotations

import base64
import json
import socket
import struct


def send_msg(sock: socket.socket, obj) -> None:
    data = json.dumps(obj).encode()
    sock.sendall(struct.pack("!I", len(data)) + data)


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed")
        buf += chunk
    return buf


def recv_msg(sock: socket.socket):
    (n,) = struct.unpack("!I", _recv_exact(sock, 4))
    return json.loads(_recv_exact(sock, n).decode())


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode()


def unb64(s: str) -> bytes:
    return base64.b64decode(s.encode())
