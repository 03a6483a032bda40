"""Python host side of the B200 range-Doppler + OS-CFAR path (ctypes over include/rdp.h).

The reference (Aurellia-Beam/fpga-fmcw-radar-processor) has no software API; its interface
for this path is the `radar_core` entity -- generics, ADC stream in, detection stream out,
two control pins (rtl/src/radar_core.vhd:11-57) -- and the `r d mag` text records its
testbenches write (rtl/src/tb_radar_core.vhd:149-160).  `RadarCore` mirrors that entity:
constructor arguments are the generics / pins under their RTL names, `process()` is one or
more CPIs in, range-Doppler map + detection list out.

No PyTorch, no cuFFT, no CPU fallback: everything numeric happens in librdp.so's sm_100a
kernels; if the library or a GPU is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(_HERE, "librdp.so")

RDP_OK = 0
ERRORS = {-1: "RDP_E_INVALID", -2: "RDP_E_NO_DEVICE", -3: "RDP_E_CUDA", -4: "RDP_E_NOMEM", -5: "RDP_E_CAPACITY"}
STATUS_DET_OVERFLOW = 1
STATUS_WINDOW_SAT = 2          # arith="fixed": a window_multiplier output clipped (window_multiplier.vhd:151-158)
STATUS_ADC_FULL_SCALE = 4      # int16 input: a sample of the CPI sits on a rail
STATUS_MAP_RANGE = 8           # cfar_stage_fixed: a magnitude word >= 2^24 (outside the supported range)

INPUT_FORMATS = {"int16": 0, "complex64": 1}
MAG_MODES = {"amb": 0, "abs": 1, "pow": 2, "logpow": 3}
CFAR_AXES = {"as_built": 0, "as_named": 1}
CFAR_EDGES = {"zero": 0, "raster": 1}
ARITH = {"float32": 0, "fixed": 1}
COMBINE = {"none": 0, "nci": 1, "beams": 2}

# one reported track; layout == struct rdp_track
TRACK_DTYPE = np.dtype([(n, "<i4") for n in ("cpi", "channel", "id", "range_q2", "doppler_q2", "vel_r", "vel_d",
                                              "quality", "status")])
TRACK_FIRM, TRACK_COAST = 2, 3

# one `r d mag` record (+ CPI, channel); layout == struct rdp_det
DET_DTYPE = np.dtype([("cpi", "<i4"), ("channel", "<i4"), ("range", "<i4"), ("doppler", "<i4"), ("mag", "<f4")])


class RdpError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


class RdpConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "struct_size", "n_chirps", "n_samples", "n_channels", "cfar_ref_r", "cfar_ref_d", "cfar_guard_r",
        "cfar_guard_d", "rank_pct", "scale_min", "scale_nom", "scale_max", "scale_override", "mti_mode",
        "input_format", "mag_mode", "cfar_axes", "cfar_edges")] + [("window_gain", C.c_float)] + [
        (n, C.c_int32) for n in ("max_dets_per_cpi", "max_cpis", "chunk_cpis", "arith", "combine", "peak_group")]


class RdpTrackerConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max_tracks", "init_hits", "coast_max", "assoc_gate_r", "assoc_gate_d",
                                         "alpha_gain", "beta_gain")]


class RdpScatter(C.Structure):
    _fields_ = [("peer", C.c_void_p * 16), ("world", C.c_int32), ("rank", C.c_int32)]


class RdpInfo(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("sm_count", C.c_int32),
                ("chunk_cpis", C.c_int32), ("cfar_tma", C.c_int32), ("reserved", C.c_int32), ("scratch_bytes", C.c_int64), ("kernel_launches", C.c_int64),
                ("device_name", C.c_char * 128)]


# every symbol include/rdp.h declares: name -> (restype, argtypes)
_VP, _I32, _SZ = C.c_void_p, C.c_int32, C.c_size_t
SYMBOLS = {
    "rdp_default_config": (None, [C.POINTER(RdpConfig)]),
    "rdp_create": (C.c_int, [C.POINTER(RdpConfig), _I32, C.POINTER(_VP)]),
    "rdp_destroy": (C.c_int, [_VP]),
    "rdp_last_error": (C.c_char_p, [_VP]),
    "rdp_get_info": (C.c_int, [_VP, C.POINTER(RdpInfo)]),
    "rdp_process_device": (C.c_int, [_VP, _VP, _I32, _VP, _VP, _VP, _VP]),
    "rdp_process_host": (C.c_int, [_VP, _VP, _I32, _VP, _VP, _VP, _VP]),
    "rdp_sync": (C.c_int, [_VP]),
    "rdp_process_fixed_device": (C.c_int, [_VP, _VP, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rdp_range_stage": (C.c_int, [_VP, _VP, _I32, _VP]),
    "rdp_doppler_stage": (C.c_int, [_VP, _VP, _I32, _VP]),
    "rdp_cfar_stage": (C.c_int, [_VP, _VP, _I32, _VP, _VP, _VP, _VP]),
    "rdp_cfar_stage_u32": (C.c_int, [_VP, _VP, _I32, _VP, _VP, _VP, _VP]),
    "rdp_fixed_window": (C.c_int, [_VP, _VP, C.c_int64, _I32, _VP, C.POINTER(C.c_uint32)]),
    "rdp_fixed_mti": (C.c_int, [_VP, _VP, C.c_int64, _I32, _I32, _VP]),
    "rdp_fixed_magnitude": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "rdp_fixed_corner_turn": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _VP]),
    "rdp_fixed_bfp_fft": (C.c_int, [_VP, _VP, C.c_int64, _I32, _VP, _VP]),
    "rdp_range_stage_scatter": (C.c_int, [_VP, _VP, _I32, C.POINTER(RdpScatter)]),
    "rdp_doppler_stage_rows": (C.c_int, [_VP, _VP, C.c_int64, _I32, _I32, _VP]),
    "rdp_cfar_stage_rows": (C.c_int, [_VP, _VP, _I32, _I32, _VP, _VP, _VP, _VP]),
    "rdp_ipc_export": (C.c_int, [_VP, _VP, C.c_char_p]),
    "rdp_ipc_open": (C.c_int, [_VP, C.c_char_p, C.POINTER(_VP)]),
    "rdp_ipc_close": (C.c_int, [_VP, _VP]),
    "rdp_copy_d2d": (C.c_int, [_VP, _VP, _VP, _SZ]),
    "rdp_copy_d2d_rows": (C.c_int, [_VP, _VP, _SZ, _VP, _SZ, _SZ, _SZ]),
    "rdp_device_memset": (C.c_int, [_VP, _VP, C.c_int, _SZ]),
    "rdp_gather_dets": (C.c_int, [_VP, _VP, _VP, _I32, _VP, _VP, C.POINTER(C.c_int64)]),
    "rdp_tracker_reset": (C.c_int, [_VP, C.POINTER(RdpTrackerConfig)]),
    "rdp_track_device": (C.c_int, [_VP, _VP, _VP, _I32, _VP, _VP, _VP]),
    "rdp_stream_open": (C.c_int, [_VP, _I32, _I32]),
    "rdp_stream_submit": (C.c_int, [_VP, _I32, _VP, _I32]),
    "rdp_stream_wait": (C.c_int, [_VP, _I32, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_I32)]),
    "rdp_stream_close": (C.c_int, [_VP]),
    "rdp_device_alloc": (C.c_int, [_VP, _SZ, C.POINTER(_VP)]),
    "rdp_device_free": (C.c_int, [_VP, _VP]),
    "rdp_host_alloc": (C.c_int, [_VP, _SZ, C.POINTER(_VP)]),
    "rdp_host_free": (C.c_int, [_VP, _VP]),
    "rdp_copy_h2d": (C.c_int, [_VP, _VP, _VP, _SZ]),
    "rdp_copy_d2h": (C.c_int, [_VP, _VP, _VP, _SZ]),
    "rdp_timer_start": (C.c_int, [_VP]),
    "rdp_timer_stop": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "rdp_set_stage_timing": (C.c_int, [_VP, _I32]),
    "rdp_get_stage_timing": (C.c_int, [_VP, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "rdp_flush_l2": (C.c_int, [_VP]),
}

_LIB_CACHE = {}


def load_library(path: str | None = None, _test_allow_emulator: bool = False):
    """dlopen librdp.so and bind every symbol of include/rdp.h.  Fails loudly when the
    library has not been built (python __graft_entry__.py / make -C <package>)."""
    path = os.path.abspath(path or DEFAULT_LIBRARY)
    key = (path, _test_allow_emulator)
    if key in _LIB_CACHE:
        return _LIB_CACHE[key]
    if not os.path.exists(path):
        raise RdpError(-3, f"{path} not found: build the CUDA library first "
                           f"(`python __graft_entry__.py` or `make -C {_HERE}`); there is no CPU fallback")
    lib = C.CDLL(path)
    if hasattr(lib, "rdp_emu_marker") and not _test_allow_emulator:
        raise RdpError(-3, f"{path} is the host-emulation test harness, not the CUDA library; refusing to use it")
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = library / header drift
        fn.restype, fn.argtypes = res, args
    _LIB_CACHE[key] = lib
    return lib


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class DeviceBuffer:
    """Caller-owned device memory (rdp_device_alloc)."""

    def __init__(self, core: "RadarCore", nbytes: int):
        self.core, self.nbytes = core, int(nbytes)
        p = C.c_void_p()
        core._check(core.lib.rdp_device_alloc(core.h, self.nbytes, C.byref(p)))
        self.ptr = p

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        self.core._check(self.core.lib.rdp_copy_h2d(self.core.h, self.ptr, _ptr(a), a.nbytes))
        return self

    def download(self, shape, dtype, offset: int = 0) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        if out.nbytes:
            self.core._check(self.core.lib.rdp_copy_d2h(self.core.h, _ptr(out), C.c_void_p(self.ptr.value + offset), out.nbytes))
        return out

    def free(self):
        if self.ptr is not None and self.core.h:
            self.core.lib.rdp_device_free(self.core.h, self.ptr)
        self.ptr = None


class _BufferView:
    """A DeviceBuffer seen from a byte offset (no ownership)."""

    def __init__(self, buf, offset):
        self.ptr, self.nbytes, self.core = C.c_void_p(buf.ptr.value + offset), buf.nbytes - offset, buf.core

    download = DeviceBuffer.download


class RadarCore:
    """The `radar_core` entity (rtl/src/radar_core.vhd:11-57) on one B200.

    Generics keep their RTL names (n_range = N_RANGE = samples per chirp, n_doppler =
    N_DOPPLER = chirps per CPI, cfar_ref_r ... cfar_guard_d), `mti_mode=0` is the
    `mti_bypass` pin, `scale_override` the `cfar_scale_ovr` pin.
    """

    def __init__(self, n_range=1024, n_doppler=128, n_channels=1, cfar_ref_r=4, cfar_ref_d=4,
                 cfar_guard_r=2, cfar_guard_d=1, rank_pct=75, scale_min=2, scale_nom=4, scale_max=6,
                 scale_override=0, mti_mode=0, input_format="int16", mag_mode="amb",
                 cfar_axes="as_built", cfar_edges="zero", window_gain=1.0, max_dets_per_cpi=4096,
                 max_cpis=1, chunk_cpis=0, arith="float32", combine="none", peak_group=False, device=0, library=None, _test_allow_emulator=False):
        self.lib = load_library(library, _test_allow_emulator)
        cfg = RdpConfig()
        self.lib.rdp_default_config(C.byref(cfg))
        cfg.n_samples, cfg.n_chirps, cfg.n_channels = n_range, n_doppler, n_channels
        cfg.cfar_ref_r, cfg.cfar_ref_d, cfg.cfar_guard_r, cfg.cfar_guard_d = cfar_ref_r, cfar_ref_d, cfar_guard_r, cfar_guard_d
        cfg.rank_pct, cfg.scale_min, cfg.scale_nom, cfg.scale_max = rank_pct, scale_min, scale_nom, scale_max
        cfg.scale_override, cfg.mti_mode = scale_override, mti_mode
        try:
            cfg.input_format = INPUT_FORMATS[input_format]
            cfg.mag_mode = MAG_MODES[mag_mode]
            cfg.cfar_axes = CFAR_AXES[cfar_axes]
            cfg.cfar_edges = CFAR_EDGES[cfar_edges]
            cfg.arith = ARITH[arith]
            cfg.combine = COMBINE[combine]
        except KeyError as e:
            raise RdpError(-1, f"unknown option {e}") from None
        cfg.window_gain = window_gain
        cfg.max_dets_per_cpi, cfg.max_cpis, cfg.chunk_cpis = max_dets_per_cpi, max_cpis, chunk_cpis
        cfg.peak_group = 1 if peak_group else 0
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.lib.rdp_create(C.byref(cfg), device, C.byref(self.h))
        if rc != RDP_OK:
            self.h = C.c_void_p()
            raise RdpError(rc, (self.lib.rdp_last_error(None) or b"").decode())
        self.n_range, self.n_doppler, self.n_channels = n_range, n_doppler, n_channels
        self.input_format = input_format
        self.arith = arith
        n_beams = 1
        while n_beams < n_channels:
            n_beams <<= 1
        # maps per CPI the CFAR stage sees: one per channel, one combined map (nci), or one per beam (beams)
        self.map_channels = {"nci": 1, "beams": n_beams}.get(combine, n_channels)
        self.map_dtype = np.uint32 if arith == "fixed" else np.float32     # words of the range-Doppler map
        self.max_dets_per_cpi, self.max_cpis = max_dets_per_cpi, max_cpis
        self._pinned = []
        self._out = {}           # page-locked result buffers of process(), reused between calls

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc != RDP_OK:
            raise RdpError(rc, (self.lib.rdp_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self._out = {}
            for p in self._pinned:
                self.lib.rdp_host_free(self.h, p)
            self._pinned = []
            self.lib.rdp_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        i = RdpInfo()
        self._check(self.lib.rdp_get_info(self.h, C.byref(i)))
        return dict(abi_version=i.abi_version, device=i.device, sm_count=i.sm_count, chunk_cpis=i.chunk_cpis, cfar_tma=bool(i.cfar_tma),
                    scratch_bytes=i.scratch_bytes, kernel_launches=i.kernel_launches,
                    device_name=i.device_name.decode())

    def sync(self):
        self._check(self.lib.rdp_sync(self.h))

    def device_buffer(self, nbytes) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array backed by page-locked host memory (rdp_host_alloc); freed by close()."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self.lib.rdp_host_alloc(self.h, max(n, 16), C.byref(p)))
        self._pinned.append(p)
        buf = (C.c_ubyte * max(n, 16)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # -- shapes -----------------------------------------------------------------------------
    @property
    def cells_per_cpi(self) -> int:
        return self.n_channels * self.n_range * self.n_doppler

    @property
    def in_bytes_per_cpi(self) -> int:
        return self.cells_per_cpi * (4 if self.input_format == "int16" else 8)

    def _as_cube(self, cube: np.ndarray) -> np.ndarray:
        """-> C-contiguous [cpi, channel, chirp, sample(, 2)] of int16 / complex64."""
        cube = np.asarray(cube)
        if self.input_format == "int16":
            if cube.dtype != np.int16 or cube.shape[-1] != 2:
                raise RdpError(-1, "int16 input must be int16 [..., chirp, sample, 2] = (I, Q)")
            core = cube.shape[:-1]
        else:
            if cube.dtype != np.complex64:
                raise RdpError(-1, "complex64 input must be complex64 [..., chirp, sample]")
            core = cube.shape
        if core[-2:] != (self.n_doppler, self.n_range):
            raise RdpError(-1, f"frame must be [{self.n_doppler} chirps, {self.n_range} samples], got {core[-2:]}")
        lead = core[:-2]
        if len(lead) == 0:
            lead = (1, 1)
        elif len(lead) == 1:
            lead = (1, lead[0])
        elif len(lead) != 2:
            raise RdpError(-1, "input must be [chirp, sample], [channel, ...] or [cpi, channel, ...]")
        if lead[1] != self.n_channels:
            raise RdpError(-1, f"expected {self.n_channels} channels, got {lead[1]}")
        tail = (self.n_doppler, self.n_range) + ((2,) if self.input_format == "int16" else ())
        return np.ascontiguousarray(cube).reshape(lead + tail)

    # -- the path ---------------------------------------------------------------------------
    def _out_buffer(self, name, shape, dtype):
        """Page-locked result buffer, (re)allocated only when it has to grow."""
        n = int(np.prod(shape))
        cur = self._out.get(name)
        if cur is None or cur.size < n or cur.dtype != np.dtype(dtype):
            cur = self.pinned_empty((max(n, 1),), dtype)
            self._out[name] = cur
        return cur[:n].reshape(shape)

    def process(self, cube: np.ndarray, want_map: bool = True):
        """HOST buffers in, HOST results out (rdp_process_host).
        Returns (rd_map [cpi, channel, range, doppler] or None -- float32, or with arith="fixed" the uint32 17-bit
                 magnitude words of magnitude_calc,
                 detections DET_DTYPE in canonical (cpi, channel, range, doppler) order,
                 n_dets int32 [cpi] (as found; > max_dets_per_cpi means truncated), status uint32 [cpi]).
        The arrays are views of page-locked buffers owned by this object: valid until the next
        process() call or close() -- `.copy()` whatever must outlive that (touching a view after close()
        reads freed memory).  process_cpi / process_batch below return copies."""
        cube = self._as_cube(cube)
        n_cpi = cube.shape[0]
        rd_map = self._out_buffer("map", (n_cpi, self.map_channels, self.n_range, self.n_doppler), self.map_dtype) if want_map else None
        dets = self._out_buffer("dets", (n_cpi * self.max_dets_per_cpi,), DET_DTYPE)
        n_dets = self._out_buffer("n_dets", (n_cpi,), np.int32)
        status = self._out_buffer("status", (n_cpi,), np.uint32)
        self._check(self.lib.rdp_process_host(self.h, _ptr(cube), n_cpi, _ptr(rd_map) if want_map else None,
                                              _ptr(dets), _ptr(n_dets), _ptr(status)))
        total = int(np.minimum(n_dets, self.max_dets_per_cpi).sum())
        return rd_map, dets[:total], n_dets, status

    def process_device(self, d_cube: DeviceBuffer, n_cpi: int, d_map: DeviceBuffer | None,
                       d_dets: DeviceBuffer, d_counts: DeviceBuffer, d_status: DeviceBuffer):
        """DEVICE buffers in and out, asynchronous (rdp_process_device)."""
        self._check(self.lib.rdp_process_device(self.h, d_cube.ptr, n_cpi, d_map.ptr if d_map else None,
                                                d_dets.ptr, d_counts.ptr, d_status.ptr))

    def process_fixed(self, cube: np.ndarray):
        """arith="fixed" with the intermediate words exposed (rdp_process_fixed_device; host arrays in and out).
        -> (mag uint32 [cpi, channel, range, doppler], exp_range int32 [cpi, channel, chirp],
            exp_doppler int32 [cpi, channel, range], detections, n_dets, status)."""
        cube = self._as_cube(cube)
        n_cpi, n_ch = cube.shape[0], self.n_channels
        d_in = self.device_buffer(cube.nbytes).upload(cube)
        d_map = self.device_buffer(n_cpi * self.cells_per_cpi * 4)
        d_er = self.device_buffer(max(n_cpi * n_ch * self.n_doppler * 4, 16))
        d_ed = self.device_buffer(max(n_cpi * n_ch * self.n_range * 4, 16))
        d_dets = self.device_buffer(n_cpi * self.max_dets_per_cpi * DET_DTYPE.itemsize)
        d_counts, d_status = self.device_buffer(max(n_cpi * 4, 16)), self.device_buffer(max(n_cpi * 4, 16))
        try:
            self._check(self.lib.rdp_process_fixed_device(self.h, d_in.ptr, n_cpi, d_map.ptr, d_er.ptr, d_ed.ptr, d_dets.ptr,
                                                          d_counts.ptr, d_status.ptr))
            self.sync()
            dets, counts = self.unpack_device_dets(d_dets, d_counts, n_cpi)
            return (d_map.download((n_cpi, n_ch, self.n_range, self.n_doppler), np.uint32),
                    d_er.download((n_cpi, n_ch, self.n_doppler), np.int32), d_ed.download((n_cpi, n_ch, self.n_range), np.int32),
                    dets, counts, d_status.download((n_cpi,), np.uint32))
        finally:
            for b in (d_in, d_map, d_er, d_ed, d_dets, d_counts, d_status):
                b.free()

    def unpack_device_dets(self, d_dets: DeviceBuffer, d_counts: DeviceBuffer, n_cpi: int):
        if n_cpi > self.max_cpis:                                    # the pack scratch holds max_cpis CPIs
            parts = [self.unpack_device_dets(_BufferView(d_dets, i * self.max_dets_per_cpi * DET_DTYPE.itemsize),
                                             _BufferView(d_counts, i * 4), min(self.max_cpis, n_cpi - i))
                     for i in range(0, n_cpi, self.max_cpis)]
            return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        counts = np.zeros(n_cpi, np.int32)
        if getattr(self, "_gather_buf", None) is None or len(self._gather_buf) < 1024:
            self._gather_buf = np.empty(1024, DET_DTYPE)
        total = C.c_int64()
        peek = d_counts.download((n_cpi,), np.int32) if n_cpi else counts      # size the (reused) host buffer first
        need = int(np.minimum(peek, self.max_dets_per_cpi).sum())
        if need > len(self._gather_buf):
            self._gather_buf = np.empty(max(need, 2 * len(self._gather_buf)), DET_DTYPE)
        self._check(self.lib.rdp_gather_dets(self.h, d_dets.ptr, d_counts.ptr, n_cpi, _ptr(self._gather_buf), _ptr(counts),
                                             C.byref(total)))
        return self._gather_buf[:total.value].copy(), counts

    # -- stage taps (device work, host arrays in/out; for parity tests) -----------------------
    def range_stage(self, cube: np.ndarray) -> np.ndarray:
        cube = self._as_cube(cube)
        n_cpi = cube.shape[0]
        d_in = self.device_buffer(cube.nbytes).upload(cube)
        d_out = self.device_buffer(n_cpi * self.cells_per_cpi * 8)
        try:
            self._check(self.lib.rdp_range_stage(self.h, d_in.ptr, n_cpi, d_out.ptr))
            self.sync()
            return d_out.download((n_cpi, self.n_channels, self.n_range, self.n_doppler), np.complex64)
        finally:
            d_in.free(); d_out.free()

    def doppler_stage(self, x_rc: np.ndarray) -> np.ndarray:
        x_rc = np.ascontiguousarray(x_rc, dtype=np.complex64).reshape(-1, self.n_channels, self.n_range, self.n_doppler)
        n_cpi = x_rc.shape[0]
        d_in = self.device_buffer(x_rc.nbytes).upload(x_rc)
        d_out = self.device_buffer(n_cpi * self.map_channels * self.n_range * self.n_doppler * 4)
        try:
            self._check(self.lib.rdp_doppler_stage(self.h, d_in.ptr, n_cpi, d_out.ptr))
            self.sync()
            return d_out.download((n_cpi, self.map_channels, self.n_range, self.n_doppler), np.float32)
        finally:
            d_in.free(); d_out.free()

    def cfar_stage(self, rd_map: np.ndarray, want_dense: bool = True):
        """-> (detections, n_dets, status, cfar_out or None); cfar_out is the RTL's dense output
        word per cell (CUT magnitude or 0)."""
        rd_map = np.ascontiguousarray(rd_map, dtype=np.float32).reshape(-1, self.map_channels, self.n_range, self.n_doppler)
        n_cpi = rd_map.shape[0]
        d_map = self.device_buffer(rd_map.nbytes).upload(rd_map)
        d_dense = self.device_buffer(rd_map.nbytes) if want_dense else None
        d_dets = self.device_buffer(n_cpi * self.max_dets_per_cpi * DET_DTYPE.itemsize)
        d_counts = self.device_buffer(n_cpi * 4)
        d_status = self.device_buffer(n_cpi * 4)
        try:
            self._check(self.lib.rdp_cfar_stage(self.h, d_map.ptr, n_cpi, d_dense.ptr if want_dense else None,
                                                d_dets.ptr, d_counts.ptr, d_status.ptr))
            self.sync()
            dets, counts = self.unpack_device_dets(d_dets, d_counts, n_cpi)
            status = d_status.download((n_cpi,), np.uint32)
            dense = d_dense.download(rd_map.shape, np.float32) if want_dense else None
            return dets, counts, status, dense
        finally:
            for b in (d_map, d_dense, d_dets, d_counts, d_status):
                if b is not None:
                    b.free()

    # -- track-while-scan processor (tws_tracker) on the detection list -------------------------
    def tracker_reset(self, max_tracks=32, init_hits=2, coast_max=5, assoc_gate_r=10, assoc_gate_d=5,
                      alpha_gain=128, beta_gain=64):
        """aresetn of every tracker instance (one per channel); generics as radar_core.vhd:425-429."""
        cfg = RdpTrackerConfig(max_tracks, init_hits, coast_max, assoc_gate_r, assoc_gate_d, alpha_gain, beta_gain)
        self._check(self.lib.rdp_tracker_reset(self.h, C.byref(cfg)))

    def track(self, dets: np.ndarray, n_dets: np.ndarray):
        """One scan per CPI, one tracker per channel: `dets` = the canonical list process() returns (CPI indices
        0..n_cpi-1), `n_dets` its per-CPI counts.  -> (tracks TRACK_DTYPE in (cpi, channel, id) order,
        n_tracks int32 [cpi, channel], active int32 [cpi, channel]).  Track files persist between calls."""
        n_dets = np.ascontiguousarray(n_dets, dtype=np.int32)
        n_cpi = n_dets.size
        kept = np.minimum(n_dets, self.max_dets_per_cpi)
        rows = np.zeros((n_cpi, self.max_dets_per_cpi), dtype=DET_DTYPE)
        off = 0
        for i in range(n_cpi):
            rows[i, :kept[i]] = dets[off:off + kept[i]]
            off += int(kept[i])
        n_ch = self.n_channels
        d_dets = self.device_buffer(rows.nbytes).upload(rows)
        d_cnt = self.device_buffer(max(n_cpi * 4, 16)).upload(n_dets)
        d_trk = self.device_buffer(max(n_cpi * n_ch * 32 * TRACK_DTYPE.itemsize, 16))
        d_nt, d_act = self.device_buffer(max(n_cpi * n_ch * 4, 16)), self.device_buffer(max(n_cpi * n_ch * 4, 16))
        try:
            self._check(self.lib.rdp_track_device(self.h, d_dets.ptr, d_cnt.ptr, n_cpi, d_trk.ptr, d_nt.ptr, d_act.ptr))
            self.sync()
            trk = d_trk.download((n_cpi, n_ch, 32), TRACK_DTYPE)
            nt = d_nt.download((n_cpi, n_ch), np.int32)
            act = d_act.download((n_cpi, n_ch), np.int32)
            out = [trk[i, c, :nt[i, c]] for i in range(n_cpi) for c in range(n_ch)]
            return (np.concatenate(out) if out else np.zeros(0, TRACK_DTYPE)), nt, act
        finally:
            for b in (d_dets, d_cnt, d_trk, d_nt, d_act):
                b.free()

    # -- 16-bit fixed-point stage taps (bit-exact RTL stages; host arrays in/out) ----------------
    def _fixed_io(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.int16)
        if x.shape[-1] != 2:
            raise RdpError(-1, "fixed-point input must be int16 [..., 2] = (I, Q)")
        return x, self.device_buffer(x.nbytes).upload(x), self.device_buffer(max(x.nbytes, 16))

    def fixed_window(self, x: np.ndarray):
        """window_multiplier along the second-to-last axis' rows: x int16 [..., n, 2] -> (y int16, saturated bool)."""
        x, d_in, d_out = self._fixed_io(x)
        n = x.shape[-2]
        sat = C.c_uint32(0)
        try:
            self._check(self.lib.rdp_fixed_window(self.h, d_in.ptr, x.size // (2 * n), n, d_out.ptr, C.byref(sat)))
            self.sync()
            return d_out.download(x.shape, np.int16), bool(sat.value)
        finally:
            d_in.free(); d_out.free()

    def fixed_mti(self, x: np.ndarray, mode: int = 2) -> np.ndarray:
        """doppler_notch along lines: x int16 [..., n, 2] -> y int16."""
        x, d_in, d_out = self._fixed_io(x)
        n = x.shape[-2]
        try:
            self._check(self.lib.rdp_fixed_mti(self.h, d_in.ptr, x.size // (2 * n), n, mode, d_out.ptr))
            self.sync()
            return d_out.download(x.shape, np.int16)
        finally:
            d_in.free(); d_out.free()

    def fixed_magnitude(self, x: np.ndarray) -> np.ndarray:
        """magnitude_calc: x int16 [..., 2] -> uint32 [...]."""
        x, d_in, d_out = self._fixed_io(x)
        try:
            self._check(self.lib.rdp_fixed_magnitude(self.h, d_in.ptr, x.size // 2, d_out.ptr))
            self.sync()
            return d_out.download(x.shape[:-1], np.uint32)
        finally:
            d_in.free(); d_out.free()

    def fixed_corner_turn(self, x: np.ndarray) -> np.ndarray:
        """corner_turner: x int16 [map, chirp, sample, 2] -> [map, range, chirp, 2]."""
        x, d_in, d_out = self._fixed_io(x)
        n_maps, n_chirps, n_samples = x.shape[0], x.shape[1], x.shape[2]
        try:
            self._check(self.lib.rdp_fixed_corner_turn(self.h, d_in.ptr, n_maps, n_chirps, n_samples, d_out.ptr))
            self.sync()
            return d_out.download((n_maps, n_samples, n_chirps, 2), np.int16)
        finally:
            d_in.free(); d_out.free()

    def fixed_bfp_fft(self, x: np.ndarray):
        """The block-floating-point transform of arith="fixed" along the second-to-last axis: x int16 [..., n, 2] ->
        (X int16 [..., n, 2], exponent int32 [...]), DFT(x) ~= X * 2^exponent (oracle.rtl_fixed.bfp_fft)."""
        x, d_in, d_out = self._fixed_io(x)
        n = x.shape[-2]
        frames = x.size // (2 * n)
        d_exp = self.device_buffer(max(frames * 4, 16))
        try:
            self._check(self.lib.rdp_fixed_bfp_fft(self.h, d_in.ptr, frames, n, d_out.ptr, d_exp.ptr))
            self.sync()
            return d_out.download(x.shape, np.int16), d_exp.download(x.shape[:-2], np.int32)
        finally:
            d_in.free(); d_out.free(); d_exp.free()

    # -- streaming ingest (BASELINE config 5) -------------------------------------------------
    def stream_open(self, n_slots: int = 3, slot_cpis: int = 1):
        self._check(self.lib.rdp_stream_open(self.h, n_slots, slot_cpis))
        self._slot_cpis = slot_cpis

    def stream_submit(self, slot: int, cube: np.ndarray):
        """Asynchronous: `cube` (ideally from pinned_empty) must stay untouched until stream_wait(slot)."""
        cube = self._as_cube(cube)
        self._check(self.lib.rdp_stream_submit(self.h, slot, _ptr(cube), cube.shape[0]))
        return cube

    def stream_wait(self, slot: int):
        """-> (detections, n_dets, status) views of the slot's page-locked result buffers."""
        pd, pn, ps, tot = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int32()
        self._check(self.lib.rdp_stream_wait(self.h, slot, C.byref(pd), C.byref(pn), C.byref(ps), C.byref(tot)))
        n = self._slot_cpis
        counts = np.frombuffer((C.c_int32 * n).from_address(pn.value), dtype=np.int32)
        status = np.frombuffer((C.c_uint32 * n).from_address(ps.value), dtype=np.uint32)
        total = int(tot.value)
        dets = np.frombuffer((C.c_ubyte * (max(total, 1) * DET_DTYPE.itemsize)).from_address(pd.value), dtype=DET_DTYPE)[:total]
        return dets, counts, status

    def stream_close(self):
        self._check(self.lib.rdp_stream_close(self.h))

    def cfar_stage_fixed(self, mag_map: np.ndarray, want_dense: bool = True):
        """Fixed-point CFAR stage (rdp_cfar_stage_u32): unsigned integer magnitudes in, decisions by the
        bit-exact integer arithmetic of rtl/src/os_cfar_2d.vhd.  -> (detections, n_dets, status, cfar_out uint32)."""
        mag_map = np.ascontiguousarray(mag_map, dtype=np.uint32).reshape(-1, self.map_channels, self.n_range, self.n_doppler)
        n_cpi = mag_map.shape[0]
        d_map = self.device_buffer(mag_map.nbytes).upload(mag_map)
        d_dense = self.device_buffer(mag_map.nbytes) if want_dense else None
        d_dets = self.device_buffer(n_cpi * self.max_dets_per_cpi * DET_DTYPE.itemsize)
        d_counts = self.device_buffer(n_cpi * 4)
        d_status = self.device_buffer(n_cpi * 4)
        try:
            self._check(self.lib.rdp_cfar_stage_u32(self.h, d_map.ptr, n_cpi, d_dense.ptr if want_dense else None,
                                                    d_dets.ptr, d_counts.ptr, d_status.ptr))
            self.sync()
            dets, counts = self.unpack_device_dets(d_dets, d_counts, n_cpi)
            status = d_status.download((n_cpi,), np.uint32)
            dense = d_dense.download(mag_map.shape, np.uint32) if want_dense else None
            return dets, counts, status, dense
        finally:
            for b in (d_map, d_dense, d_dets, d_counts, d_status):
                if b is not None:
                    b.free()

    # -- measurement ------------------------------------------------------------------------
    def timer_start(self):
        self._check(self.lib.rdp_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self.lib.rdp_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def set_stage_timing(self, on: bool):
        self._check(self.lib.rdp_set_stage_timing(self.h, 1 if on else 0))

    def stage_timing(self):
        ms = (C.c_float * 4)()
        n = (C.c_int64 * 4)()
        self._check(self.lib.rdp_get_stage_timing(self.h, ms, n))
        names = ("range", "doppler", "cfar", "compact")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}

    def flush_l2(self):
        self._check(self.lib.rdp_flush_l2(self.h))


# ------------------------------------------------------------------------------------------------
# Frame-in / map + detection-list-out functions, same signatures as oracle.float_path
# ------------------------------------------------------------------------------------------------
def _core_for(cubes: np.ndarray, n_cpi_axis: bool, kw: dict) -> RadarCore:
    cubes = np.asarray(cubes)
    cplx = np.iscomplexobj(cubes)
    core_shape = cubes.shape if cplx else cubes.shape[:-1]
    n_doppler, n_range = core_shape[-2], core_shape[-1]
    lead = core_shape[:-2]
    n_channels = lead[-1] if len(lead) >= 1 else 1
    n_cpi = lead[0] if (n_cpi_axis and len(lead) == 2) else 1
    return RadarCore(n_range=n_range, n_doppler=n_doppler, n_channels=n_channels,
                     input_format="complex64" if cplx else "int16", max_cpis=max(1, n_cpi), **kw)


def process_cpi(frame: np.ndarray, **kw):
    """One CPI: frame [chirp, sample(,2)] or [channel, chirp, sample(,2)] -> (rd_map float32
    [channel, range, doppler], detections sorted by (channel, range, doppler))."""
    frame = np.asarray(frame)
    if np.iscomplexobj(frame):
        frame = frame.astype(np.complex64, copy=False)
    with _core_for(frame, False, kw) as core:
        m, dets, _, _ = core.process(frame)
        return m[0].copy(), dets.copy()


def process_batch(cubes: np.ndarray, **kw):
    """[cpi, channel, chirp, sample(,2)] -> (maps, detections for all CPIs, per-CPI counts)."""
    cubes = np.asarray(cubes)
    if np.iscomplexobj(cubes):
        cubes = cubes.astype(np.complex64, copy=False)
    with _core_for(cubes, True, kw) as core:
        m, dets, counts, _ = core.process(cubes)
        return m.copy(), dets.copy(), counts.copy()
