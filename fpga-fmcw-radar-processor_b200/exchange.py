"""Rendezvous for the ranks of the sharded large-CPI mode: all-gather of small Python objects (the 64-byte CUDA
IPC handles) and a barrier, over plain TCP sockets -- no torch, no MPI.  Rank 0 listens on MASTER_ADDR :
(MASTER_PORT + port_offset), the other ranks connect; the radar data itself never goes through here (it crosses
NVLink as K1's peer stores)."""
from __future__ import annotations

import os
import time
from multiprocessing.connection import Client, Listener

_AUTH = b"rdp-sharded-cpi"


class SocketExchange:
    def __init__(self, rank=None, world=None, addr=None, port=None, port_offset=1, timeout=120.0):
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else int(world)
        addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(port if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + port_offset)
        self._listener, self._conns, self._conn = None, [], None
        if self.world == 1:
            return
        if self.rank == 0:
            self._listener = Listener((addr, port), authkey=_AUTH)
            peers = {}
            while len(peers) < self.world - 1:
                c = self._listener.accept()
                peers[int(c.recv())] = c
            self._conns = [peers[r] for r in range(1, self.world)]
        else:
            deadline = time.time() + timeout
            while True:
                try:
                    self._conn = Client((addr, port), authkey=_AUTH)
                    break
                except (ConnectionRefusedError, OSError):
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            self._conn.send(self.rank)

    def all_gather(self, obj):
        """-> [obj of rank 0, ..., obj of rank world-1] on every rank."""
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            out = [obj] + [c.recv() for c in self._conns]
            for c in self._conns:
                c.send(out)
            return out
        self._conn.send(obj)
        return self._conn.recv()

    def barrier(self):
        self.all_gather(None)

    def close(self):
        for c in self._conns:
            c.close()
        if self._conn is not None:
            self._conn.close()
        if self._listener is not None:
            self._listener.close()
        self._conns, self._conn, self._listener = [], None, None
