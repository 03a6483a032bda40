"""One CPI sharded over the GPUs of a box (BASELINE config 4: the distributed corner turn).

The chirps of a CPI are dealt to `world` ranks (one process per GPU).  Each rank windows and range-FFTs its
own chirps, and K1's transposed store writes every range line straight into the memory of the rank that
owns it -- peer buffers mapped through CUDA IPC, stores over NVLink -- so the all-to-all corner turn costs
no separate collective and no repack (`rdp_range_stage_scatter`).  After a barrier each rank runs the
Doppler stage on its n_samples / world range lines, copies the CFAR halo lines from its two neighbours'
maps, and runs the CFAR stage on (own + halo) lines.  Detections come back with global range indices.

Ranks only need a way to exchange 64-byte IPC handles and to barrier: `exchange.SocketExchange` does that over
plain TCP sockets (no torch, no MPI); the data itself never goes through a collective.  For the hot path of
the benchmark none of this is needed -- CPIs shard across GPUs with no exchange at all (sharding.py); this
mode exists for a CPI too large for one GPU's memory, and is exercised at small sizes by the tests.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .rdp import DET_DTYPE, RadarCore, RdpError, RdpScatter


class ShardedCpi:
    def __init__(self, n_chirps, n_samples, world, rank, n_channels=1, n_cpi=1, input_format="int16", device=0,
                 library=None, _test_allow_emulator=False, max_dets_per_cpi=1 << 16, **params):
        if n_chirps % world or n_samples % world:
            raise RdpError(-1, "n_chirps and n_samples must be multiples of the number of ranks")
        if params.get("cfar_edges", "zero") != "zero":
            raise RdpError(-1, "the sharded mode supports cfar_edges='zero' only")
        self.world, self.rank, self.n_cpi, self.n_channels = world, rank, n_cpi, n_channels
        self.n_chirps, self.n_samples = n_chirps, n_samples
        self.chirps_local, self.rows_local = n_chirps // world, n_samples // world
        kw = dict(library=library, _test_allow_emulator=_test_allow_emulator, device=device)
        fft_only = {k: v for k, v in params.items() if k in ("window_gain",)}
        # hA: range stage on the local chirps; hB: Doppler + CFAR stages on the local range lines
        self.core_a = RadarCore(n_range=n_samples, n_doppler=self.chirps_local, n_channels=n_channels,
                                input_format=input_format, max_cpis=n_cpi, **fft_only, **kw)
        self.core_b = RadarCore(n_range=self.rows_local, n_doppler=n_chirps, n_channels=n_channels, max_cpis=n_cpi,
                                max_dets_per_cpi=max_dets_per_cpi, **params, **kw)
        as_built = params.get("cfar_axes", "as_built") == "as_built"
        ref_r, ref_d = params.get("cfar_ref_r", 4), params.get("cfar_ref_d", 4)
        g_r, g_d = params.get("cfar_guard_r", 2), params.get("cfar_guard_d", 1)
        self.halo = (ref_d + g_d) if as_built else (ref_r + g_r)          # CFAR half extent along range
        if self.halo > self.rows_local:
            raise RdpError(-1, "CFAR halo exceeds the range lines per rank")
        maps = n_cpi * n_channels
        self.maps = maps
        self.rows_pad = self.rows_local + 2 * self.halo
        self.d_xrc = self.core_b.device_buffer(maps * self.rows_local * n_chirps * 8)
        self.d_map = self.core_b.device_buffer(maps * self.rows_pad * n_chirps * 4)
        self.d_dets = self.core_b.device_buffer(n_cpi * max_dets_per_cpi * DET_DTYPE.itemsize)
        self.d_cnt, self.d_st = self.core_b.device_buffer(max(n_cpi * 4, 16)), self.core_b.device_buffer(max(n_cpi * 4, 16))
        self.core_b._check(self.core_b.lib.rdp_device_memset(self.core_b.h, self.d_map.ptr, 0, self.d_map.nbytes))
        self.core_b.sync()                                           # halo lines past the global edges stay 0
        self.peer_xrc = [None] * world
        self.peer_map = [None] * world
        self._opened = []

    # -- phase A / B: exchange of IPC handles -------------------------------------------------------
    def export_handles(self):
        lib, h = self.core_b.lib, self.core_b.h
        out = []
        for buf in (self.d_xrc, self.d_map):
            hb = C.create_string_buffer(64)
            self.core_b._check(lib.rdp_ipc_export(h, buf.ptr, hb))
            out.append(hb.raw)
        # an IPC handle cannot be opened by the process that exported it: ranks simulated inside one process
        # (tests) use the raw pointers instead
        return (os.getpid(), self.d_xrc.ptr.value, self.d_map.ptr.value) + tuple(out)

    def connect(self, handles_of_all_ranks):
        lib, h = self.core_b.lib, self.core_b.h
        for r, (pid, px, pm, hx, hm) in enumerate(handles_of_all_ranks):
            if r == self.rank:
                self.peer_xrc[r], self.peer_map[r] = self.d_xrc.ptr.value, self.d_map.ptr.value
                continue
            if pid == os.getpid():
                self.peer_xrc[r], self.peer_map[r] = px, pm
                continue
            ptrs = []
            for raw in (hx, hm):
                p = C.c_void_p()
                self.core_b._check(lib.rdp_ipc_open(h, raw, C.byref(p)))
                self._opened.append(p)
                ptrs.append(p.value)
            self.peer_xrc[r], self.peer_map[r] = ptrs

    # -- phase C: range stage with the corner turn fused into peer stores -------------------------------
    def upload_local(self, cube_local: np.ndarray):
        """This rank's chirps [cpi, channel, n_chirps / world, sample(, 2)] into a device buffer (caller frees)."""
        cube = self.core_a._as_cube(cube_local)
        if cube.shape[0] != self.n_cpi:
            raise RdpError(-1, f"expected {self.n_cpi} CPIs, got {cube.shape[0]}")
        return self.core_a.device_buffer(cube.nbytes).upload(cube)

    def range_scatter_device(self, d_in):
        sc = RdpScatter()
        for r in range(self.world):
            sc.peer[r] = self.peer_xrc[r]
        sc.world, sc.rank = self.world, self.rank
        self.core_a._check(self.core_a.lib.rdp_range_stage_scatter(self.core_a.h, d_in.ptr, self.n_cpi, C.byref(sc)))
        self.core_a.sync()                                           # the peer stores have landed before the barrier

    def range_scatter(self, cube_local: np.ndarray):
        d_in = self.upload_local(cube_local)
        try:
            self.range_scatter_device(d_in)
        finally:
            d_in.free()

    # -- phase D: Doppler stage on the local range lines, into the halo-padded map ----------------------
    def doppler(self):
        cb = self.core_b                                             # one launch: every map gets its halo lines left free
        cb._check(cb.lib.rdp_doppler_stage_rows(cb.h, self.d_xrc.ptr, self.maps * self.rows_local, self.rows_local,
                                                self.halo, self.d_map.ptr))
        cb.sync()

    # -- phase E: CFAR halo lines from the neighbouring ranks' maps (device-to-device, over NVLink) ------
    def halo_exchange(self):
        lib, h, nc, hl = self.core_b.lib, self.core_b.h, self.n_chirps, self.halo
        if hl == 0:
            return
        pitch, rowb = self.rows_pad * nc * 4, hl * nc * 4            # one call per neighbour: `maps` pieces of hl lines
        if self.rank > 0:                                            # lines above mine = the last `hl` own lines of rank-1
            src = self.peer_map[self.rank - 1] + self.rows_local * nc * 4
            self.core_b._check(lib.rdp_copy_d2d_rows(h, self.d_map.ptr, pitch, C.c_void_p(src), pitch, rowb, self.maps))
        if self.rank < self.world - 1:                               # lines below mine = the first `hl` own lines of rank+1
            src = self.peer_map[self.rank + 1] + hl * nc * 4
            dst = self.d_map.ptr.value + (hl + self.rows_local) * nc * 4
            self.core_b._check(lib.rdp_copy_d2d_rows(h, C.c_void_p(dst), pitch, C.c_void_p(src), pitch, rowb, self.maps))

    # -- phase F: CFAR on own + halo lines; detections with global range indices ----------------------------
    def cfar(self):
        cb = self.core_b
        cb._check(cb.lib.rdp_cfar_stage_rows(cb.h, self.d_map.ptr, self.n_cpi, self.rows_pad, None, self.d_dets.ptr,
                                             self.d_cnt.ptr, self.d_st.ptr))
        cb.sync()
        dets, counts = cb.unpack_device_dets(self.d_dets, self.d_cnt, self.n_cpi)
        status = self.d_st.download((self.n_cpi,), np.uint32)
        if np.any(counts > cb.max_dets_per_cpi):
            raise RdpError(-5, "detection list overflow in the sharded CFAR; raise max_dets_per_cpi")
        own = (dets["range"] >= self.halo) & (dets["range"] < self.halo + self.rows_local)
        dets = dets[own].copy()
        dets["range"] += self.rank * self.rows_local - self.halo
        return dets, status

    def local_map(self) -> np.ndarray:
        """This rank's range lines of the RD map: float32 [cpi, channel, rows_local, n_chirps]."""
        full = self.d_map.download((self.n_cpi, self.n_channels, self.rows_pad, self.n_chirps), np.float32)
        return full[:, :, self.halo:self.halo + self.rows_local].copy()

    # -- the whole thing, one process per GPU -----------------------------------------------------------
    def run(self, cube_local: np.ndarray, exchange):
        """cube_local: this rank's chirps [cpi, channel, n_chirps / world, sample(, 2)].  Returns (this rank's
        detections with global range indices, its lines of the RD map)."""
        if self.peer_xrc[self.rank] is None:
            self.connect(exchange.all_gather(self.export_handles()) if self.world > 1 else [self.export_handles()])
        exchange.barrier()
        self.range_scatter(cube_local)
        exchange.barrier()                                           # every rank's lines are complete
        self.doppler()
        exchange.barrier()                                           # every rank's map is complete
        self.halo_exchange()
        dets, _ = self.cfar()
        exchange.barrier()                                           # nobody reuses a buffer a neighbour still reads
        return dets, self.local_map()

    def run_device(self, d_in, exchange):
        """The same pipeline on a cube that is already on the device (`upload_local`); returns this rank's
        detections only (the map stays on the device)."""
        if self.peer_xrc[self.rank] is None:
            self.connect(exchange.all_gather(self.export_handles()) if self.world > 1 else [self.export_handles()])
            exchange.barrier()
        self.range_scatter_device(d_in)
        exchange.barrier()
        self.doppler()
        exchange.barrier()
        self.halo_exchange()
        dets, _ = self.cfar()
        exchange.barrier()
        return dets

    def close(self):
        for p in self._opened:
            self.core_b.lib.rdp_ipc_close(self.core_b.h, p)
        self._opened = []
        for b in (self.d_xrc, self.d_map, self.d_dets, self.d_cnt, self.d_st):
            b.free()
        self.core_a.close()
        self.core_b.close()
