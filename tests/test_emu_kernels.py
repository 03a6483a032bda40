"""Kernel LOGIC under host emulation (CPU, no GPU): the same kernel sources as librdp.so,
compiled with g++ against tests/emu/cuda_emu.h (one OS thread per CUDA thread), driven through
the same C ABI and compared with the oracle at tiny sizes.  This is a development aid that
catches indexing / counting mistakes before GPU time is spent; it is not a product path
(the product loader refuses this library) and proves nothing about the real device -- the GPU
parity tests in test_gpu_parity.py do that."""
import os
import subprocess

import numpy as np
import pytest

import fpga_fmcw_radar_processor_b200 as rdp
import parity_cases as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "librdp_emu.so")


@pytest.fixture(scope="module")
def make_core():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "fpga-fmcw-radar-processor_b200"), "emu"])

    def factory(**kw):
        return rdp.RadarCore(library=EMU, _test_allow_emulator=True, **kw)
    return factory


@pytest.mark.parametrize("ns,nc,fmt", [(64, 32, "int16"), (128, 16, "complex64"), (16, 8, "int16"), (512, 16, "int16")])
def test_range_stage(make_core, ns, nc, fmt):
    pc.case_range_stage(make_core, ns, nc, fmt=fmt)


def test_corner_turn_order(make_core):
    pc.case_corner_turn_order(make_core)


@pytest.mark.parametrize("mti,mag", [(0, "amb"), (2, "abs"), (3, "pow"), (0, "logpow")])
def test_doppler_stage(make_core, mti, mag):
    pc.case_doppler_stage(make_core, 16, 64, mti, mag)


def test_doppler_stage_256(make_core):
    pc.case_doppler_stage(make_core, 8, 256, 2, "amb", window_gain=2.0)


@pytest.mark.parametrize("nr,nd,kw", [
    (64, 128, dict()),                                          # radar_core defaults, 2 x 2 tiles
    (16, 32, dict(cfar_edges="raster")),                        # partial tile
    (64, 32, dict(cfar_axes="as_named", scale_override=3)),
    (32, 64, dict(cfar_ref_r=3, cfar_ref_d=2, cfar_guard_r=1, cfar_guard_d=1)),
    (32, 32, dict(cfar_ref_r=1, cfar_ref_d=2, cfar_guard_r=0, cfar_guard_d=1, cfar_edges="raster")),   # generic kernel
    (32, 64, dict(rank_pct=50, scale_min=3, scale_nom=5, scale_max=7)),
    (32, 64, dict(cfar_ref_r=8, cfar_ref_d=8, cfar_guard_r=4, cfar_guard_d=2)),   # BASELINE config 4 "wide"
])
def test_cfar_stage_bit_exact(make_core, nr, nd, kw):
    pc.case_cfar_stage(make_core, nr, nd, **kw)


def test_cfar_tb_map(make_core):
    pc.case_cfar_tb_map(make_core)


@pytest.mark.parametrize("nr,nd,kw", [(32, 64, dict()), (64, 128, dict(cfar_edges="raster", cfar_axes="as_named")),
                                      (32, 64, dict(scale_override=2))])
def test_cfar_signed_map(make_core, nr, nd, kw):
    pc.case_cfar_signed_map(make_core, nr, nd, **kw)


def test_cfar_tiny_magnitudes(make_core):
    pc.case_cfar_tiny_magnitudes(make_core)
    pc.case_cfar_tiny_magnitudes(make_core, scale_override=3, cfar_edges="raster")


def test_cfar_crowded_tile(make_core):
    pc.case_cfar_crowded_tile(make_core)
    pc.case_cfar_crowded_tile(make_core, scale_override=2)


def test_end_to_end_mag_modes(make_core):
    pc.case_end_to_end_mag_modes(make_core)


def test_end_to_end(make_core):
    pc.case_end_to_end(make_core, 2, 2, 32, 64)


def test_overflow_flag(make_core):
    pc.case_overflow_flag(make_core)


def test_chunking_invariance(make_core):
    pc.case_chunking_invariance(make_core)


@pytest.mark.parametrize("kw", [dict(), dict(cfar_edges="raster"), dict(scale_override=3), dict(cfar_axes="as_named"),
                                dict(rank_pct=50, scale_min=1, scale_nom=3, scale_max=5)])
def test_cfar_fixed_point_bit_exact(make_core, kw):
    pc.case_cfar_fixed_point(make_core, 32, 64, **kw)


def test_cfar_fixed_tb_map(make_core):
    pc.case_cfar_fixed_tb_map(make_core)


def test_streaming(make_core):
    pc.case_streaming(make_core)
    pc.case_streaming(make_core, spec=2)


def test_fixed_stages(make_core):
    pc.case_fixed_stages(make_core)


@pytest.mark.parametrize("seed", range(6))
def test_cfar_random_config(make_core, seed):
    pc.case_cfar_random_config(make_core, 1000 + seed)


def test_tracker(make_core):
    pc.case_tracker(make_core)


@pytest.fixture(scope="module")
def make_sharded(make_core):
    def factory(**kw):
        return rdp.ShardedCpi(library=EMU, _test_allow_emulator=True, **kw)
    return factory


@pytest.mark.parametrize("world,kw", [(4, dict()), (2, dict(cfar_axes="as_named", mti_mode=2, input_format="int16"))])
def test_sharded_cpi(make_sharded, make_core, world, kw):
    pc.case_sharded_cpi(make_sharded, make_core, world=world, **kw)


@pytest.mark.parametrize("seed", range(6))
def test_cfar_adaptive_stress(make_core, seed):
    pc.case_cfar_adaptive_stress(make_core, seed)


def test_bfp_fft(make_core):
    pc.case_bfp_fft(make_core, sizes=(2, 8, 64, 256))


@pytest.mark.parametrize("mti,kw", [(0, dict()), (2, dict(cfar_axes="as_named")), (3, dict(scale_override=3))])
def test_fixed_end_to_end(make_core, mti, kw):
    pc.case_fixed_end_to_end(make_core, mti_mode=mti, **kw)


def test_fixed_end_to_end_saturating(make_core):
    # full-scale input: the as-built x2 window gain clips (window_multiplier.vhd:146-158) -> RDP_STATUS_WINDOW_SAT
    assert pc.case_fixed_end_to_end(make_core, n_cpi=2, n_ch=1, n_chirps=16, n_samples=64, amp_scale=40.0) > 0


def test_adc_full_scale_flag(make_core):
    pc.case_adc_full_scale_flag(make_core)


@pytest.mark.parametrize("fmt,mti,kw", [("int16", 0, dict()), ("complex64", 2, dict(mag_mode="pow", cfar_edges="raster"))])
def test_combine_and_group(make_core, fmt, mti, kw):
    pc.case_combine_and_group(make_core, fmt=fmt, mti_mode=mti, **kw)


def test_cfar_fixed_range_flag(make_core):
    pc.case_cfar_fixed_range_flag(make_core)
