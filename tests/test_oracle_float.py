"""Oracle (float path) against the reference's two data files (structural pins only --
SURVEY.md findings 0.2/0.3: no bit-exact pin exists for the FFT stages) and self-consistency.
CPU only; reads the committed fixtures under tests/golden/, never /root/reference."""
import os

import numpy as np
import pytest

from oracle import float_path as fp
from oracle import rtl_fixed as fx
from oracle import stimulus as st


@pytest.fixture(scope="module")
def golden_chirp(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_input_chirp_1024.npy"))


@pytest.fixture(scope="module")
def radar_output(golden_dir):
    return np.load(os.path.join(golden_dir, "radar_output_summary.npz"))


def test_golden_chirp_is_a_single_tone(golden_chirp):
    z = golden_chirp[:, 0].astype(np.float64) + 1j * golden_chirp[:, 1]
    assert np.all(np.abs(np.abs(z) - 16384) < 64)      # |z| ~ 16384 with small additive noise
    step = np.angle(z[1:] * np.conj(z[:-1]))
    assert abs(np.degrees(np.angle(np.mean(np.exp(1j * step)))) - 102.0) < 0.05      # 0.283333 cycles / sample


def test_golden_chirp_range_peak_bin_290(golden_chirp):
    """data/golden_input_chirp.txt[0:1024] -> Hamming -> forward DFT peaks at bin 290
    (0.283333 * 1024 = 290.13): pins direction (forward) and ordering (natural)."""
    frame = np.tile(golden_chirp[None], (8, 1, 1))             # 8 identical chirps
    x_rc = fp.range_stage(frame)
    assert x_rc.shape == (1024, 8)
    assert int(np.argmax(np.abs(x_rc[:, 0]))) == 290
    m = fp.rd_map(frame)
    r, d = np.unravel_index(np.argmax(m), m.shape)
    assert (r, d) == (290, 0)                                  # identical chirps => zero Doppler


def test_radar_output_structure_vs_oracle(radar_output):
    """rtl/old/tb_radar_core.vhd stimulus -> oracle; compare with the STRUCTURE of
    data/radar_output.txt: energy in range rows {99,100,101} and
    {499,500,501}; Doppler residues 5 = 1 and 118 = 2 (mod 4); background level."""
    ro = radar_output
    top6 = set(np.argsort(ro["row_max"])[::-1][:6].tolist())
    assert top6 == {99, 100, 101, 499, 500, 501}
    cube = st.old_tb_radar_core_cube(rng=np.random.default_rng(1))
    m = fp.rd_map(cube)
    ours6 = set(np.argsort(m.max(axis=1))[::-1][:6].tolist())
    assert ours6 == top6
    # oracle peaks: target 1 at (100, +5), target 2 at (500, -10 = 118)
    assert int(np.argmax(m[100])) == 5 and int(np.argmax(m[500])) == 118
    # the legacy file shows a comb on the Doppler axis with the same residues mod 4
    rows = ro["target_rows"].tolist()
    v100 = ro["target_row_values"][rows.index(100)]
    v500 = ro["target_row_values"][rows.index(500)]
    assert int(np.argmax(v100)) % 4 == 5 % 4 and int(np.argmax(v500)) % 4 == 118 % 4
    assert int(ro["median"]) == 104 and int(ro["n_zero"]) == 1


def test_float_matches_fixed_stages_up_to_quantisation():
    """Float window / MTI / magnitude track the bit-exact integer stages (gain x2 of the
    as-built window included) to within the quantisation step."""
    rng = np.random.default_rng(5)
    x = rng.integers(-12000, 12000, size=(4, 256))
    yf = 2.0 * x * fp.hamming(256).astype(np.float64)
    yi, sat = fx.window(x, 256)
    assert not sat.any()
    assert np.max(np.abs(yi - yf)) <= 3.0      # +1 bias, rounding, Q15 coefficient quantisation
    z = rng.integers(-20000, 20000, size=(100, 2))
    mf = fp.magnitude_float(z[:, 0] + 1j * z[:, 1], "amb")
    mi = fx.magnitude(z[:, 0], z[:, 1])
    assert np.max(np.abs(mi - mf)) <= 2.0
    w = rng.integers(-8000, 8000, size=(3, 64))
    assert np.array_equal(fx.mti(w, 2), fp.mti_float(w.astype(np.float64), 2).astype(np.int64))
    assert np.array_equal(fx.mti(w, 3), fp.mti_float(w.astype(np.float64), 3).astype(np.int64))


def test_float_cfar_equals_integer_cfar_on_integer_maps():
    """On integer-valued maps with even means the float32 CFAR equals the integer CFAR
    wherever the integer floor/shift do not bite: use override scale (no mean involved)."""
    rng = np.random.default_rng(9)
    m = rng.integers(10, 500, size=(48, 64)).astype(np.uint32)
    m[20, 30] = 9000
    for edges in ("zero", "raster"):
        for ovr in (2, 4, 6):
            a = fx.os_cfar_2d(m, scale_override=ovr, edges=edges)
            b = fp.cfar_stage(m.astype(np.float32), scale_override=ovr, cfar_edges=edges)
            assert np.array_equal(a, b.astype(np.uint32))


def test_process_cpi_detects_config2_targets():
    """BASELINE config 2 (RTL-native shape, 3 targets): detections cluster on the targets."""
    tg = [(100, 5, 8000.0), (500, 118, 5000.0), (750, 40, 3000.0)]
    cube, _ = st.synthetic_cube(1, 1, 128, 1024, seed=1234, sigma=12.0, fixed_targets=tg)
    m, dets = fp.process_cpi(cube[0])
    assert m.shape == (1, 1024, 128) and m.dtype == np.float32
    cells = {(int(r), int(d)) for r, d in zip(dets["range"], dets["doppler"])}
    for (r, d, _) in tg:
        assert (r, d) in cells
    # canonical order
    key = dets["channel"].astype(np.int64) << 40 | dets["range"].astype(np.int64) << 20 | dets["doppler"]
    assert np.all(np.diff(key) > 0)


def test_ambiguous_cells_cover_perturbation_flips():
    rng = np.random.default_rng(21)
    m = rng.rayleigh(100.0, size=(64, 64)).astype(np.float32)
    m[30, 30] = 3000.0
    base = fp.cfar_stage(m) != 0
    eps = 0.05
    amb = fp.ambiguous_cells(m, eps)
    for trial in range(5):
        pert = (m + rng.uniform(-eps, eps, m.shape)).astype(np.float32)
        flipped = (fp.cfar_stage(pert) != 0) != base
        assert not np.any(flipped & ~amb)
    assert amb.mean() < 0.05
