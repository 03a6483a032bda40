"""The C-ABI library loads and exports every symbol include/rdp.h declares; struct layouts
agree between the header, the library and the ctypes mirror; without a GPU the product fails
loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import fpga_fmcw_radar_processor_b200 as rdp
from fpga_fmcw_radar_processor_b200 import rdp as host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rdp.h")
EMU = os.path.join(ROOT, "tests", "emu", "librdp_emu.so")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rdp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_match_ctypes_table():
    assert _declared_symbols() == sorted(host.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(host.DEFAULT_LIBRARY)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/rdp.h but not exported by librdp.so"
    assert not hasattr(lib, "rdp_emu_marker")          # the product library is not the emulator


def test_struct_layouts():
    lib = rdp.load_library()
    cfg = host.RdpConfig()
    lib.rdp_default_config(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(host.RdpConfig) == 100
    # reference defaults: rtl/src/radar_core.vhd:12-20, :380
    assert (cfg.n_samples, cfg.n_chirps) == (1024, 128)
    assert (cfg.cfar_ref_r, cfg.cfar_ref_d, cfg.cfar_guard_r, cfg.cfar_guard_d) == (4, 4, 2, 1)
    assert (cfg.rank_pct, cfg.scale_min, cfg.scale_nom, cfg.scale_max, cfg.scale_override) == (75, 2, 4, 6, 0)
    assert rdp.DET_DTYPE.itemsize == 20


def test_invalid_config_is_rejected_with_message():
    lib = rdp.load_library()
    cfg = host.RdpConfig()
    lib.rdp_default_config(C.byref(cfg))
    cfg.n_samples = 1000                                  # not a power of two
    h = C.c_void_p()
    assert lib.rdp_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"n_samples" in lib.rdp_last_error(None)
    cfg.n_samples = 1024
    cfg.arith = 1                                         # FIXED16 needs int16 input, AMB magnitude, unit window gain
    cfg.mag_mode = 2
    assert lib.rdp_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"FIXED16" in lib.rdp_last_error(None)
    cfg.arith, cfg.mag_mode = 0, 0
    cfg.struct_size = 4
    assert lib.rdp_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"ABI" in lib.rdp_last_error(None)


def _has_gpu():
    try:
        lib = C.CDLL("libcuda.so.1")
        n = C.c_int(0)
        return lib.cuInit(0) == 0 and lib.cuDeviceGetCount(C.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    with pytest.raises(rdp.RdpError) as e:
        rdp.RadarCore(n_range=64, n_doppler=32)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    import numpy as np
    with pytest.raises(rdp.RdpError):
        rdp.process_cpi(np.zeros((32, 64, 2), np.int16))


def test_product_loader_refuses_the_emulator():
    if not os.path.exists(EMU):
        pytest.skip("emulation harness not built")
    with pytest.raises(rdp.RdpError):
        rdp.load_library(EMU)
    with pytest.raises(rdp.RdpError):
        rdp.load_library(os.path.join(ROOT, "does_not_exist.so"))


def test_header_is_plain_c99(tmp_path):
    """include/rdp.h is the drop-in boundary: a C caller (no C++, no CUDA headers) must be able to include it."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not installed")
    src = tmp_path / "use_rdp.c"
    src.write_text('#include "rdp.h"\n'
                   "int main(void) { rdp_config c; rdp_det d; rdp_info i; rdp_scatter s; rdp_track t; rdp_tracker_config k;\n"
                   "  (void)c; (void)d; (void)i; (void)s; (void)t; (void)k; return (int)sizeof(rdp_config) == 0; }\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                          "-c", str(src), "-o", str(tmp_path / "use_rdp.o")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_c_example_builds_and_links():
    """examples/rdp_frame.c -- the plain-C caller of the ABI -- compiles and links against librdp.so (make example)."""
    import subprocess
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "fpga-fmcw-radar-processor_b200"), "example"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert os.access(os.path.join(ROOT, "examples", "rdp_frame"), os.X_OK)
