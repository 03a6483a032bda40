"""Oracle (integer stages) against every known answer the reference's own testbenches hold.
CPU only.  Citations are into /root/reference."""
import os

import numpy as np
import pytest

from oracle import rtl_fixed as fx
from oracle import stimulus as st


# --- magnitude: rtl/src/tb_magnitude_calc.vhd:49-73 (17 vectors), amb_ref :32-40 -------------
MAG_VECTORS = [(1000, 0), (0, 1000), (-1000, 0), (0, -1000), (1000, 1000), (-1000, 1000),
               (-1000, -1000), (1000, -1000), (5000, 3000), (-8000, 6000), (100, 100),
               (32000, 32000), (0, 0), (1, 0), (1, 1), (30000, 100), (100, 30000)]
# literal evaluation of amb_ref (mx + mn/4 + mn/8, integer division) for those vectors
MAG_EXPECTED = [1000, 1000, 1000, 1000, 1375, 1375, 1375, 1375, 6125, 10250, 137, 44000, 0, 1, 1,
                30037, 30037]


def test_magnitude_known_answers():
    i = np.array([v[0] for v in MAG_VECTORS])
    q = np.array([v[1] for v in MAG_VECTORS])
    got = fx.magnitude(i, q)
    assert got.tolist() == MAG_EXPECTED          # exact (the TB allows +-1 LSB, :156-158)


def test_magnitude_extremes():
    assert int(fx.magnitude(-32768, -32768)) == 32768 + 8192 + 4096   # magnitude_calc.vhd:57-58
    assert int(fx.magnitude(-32768, 0)) == 32768


# --- corner turn: rtl/src/tb_corner_turner.vhd:36-39,146-186 (16 x 8, value = chirp*256+sample)
def test_corner_turn_kat():
    n_range, n_doppler = 16, 8
    chirp = np.arange(n_doppler)[:, None]
    sample = np.arange(n_range)[None, :]
    x = chirp * 256 + sample
    y = fx.corner_turn(x)
    assert y.shape == (n_range, n_doppler)
    flat = y.reshape(-1)                 # output order: range slow, Doppler fast
    assert np.array_equal(flat // 256, np.tile(np.arange(n_doppler), n_range))
    assert np.array_equal(flat % 256, np.repeat(np.arange(n_range), n_doppler))


# --- window: rtl/src/tb_window_multiplier.vhd:83-108,185-240 (N = 64) -------------------------
def test_window_tb_checks():
    n = 64
    y, sat = fx.window(np.full(n, 16000), n)
    assert abs(y[0]) <= 3000 and abs(y[-1]) <= 3000            # endpoints  (:185-198)
    assert y[n // 2] >= 10000 and y[n // 2 - 1] >= 10000       # centre
    assert np.all(np.abs(y - y[::-1]) <= 1)                    # symmetry  (:209-222)
    assert not sat.any()
    y, sat = fx.window(np.full(n, 32767), n)                   # full scale sets the flag (:102-108)
    assert sat.any() and y.max() == 32767
    # impulse at index 10: only that sample is "large"
    x = np.zeros(n, dtype=np.int64)
    x[10] = 20000
    y, _ = fx.window(x, n)
    assert y[10] == (20000 * int(fx.hamming_rom(n)[10]) + 16384) >> 14
    # KNOWN REFERENCE INCONSISTENCY: the TB expects zero in -> zero out (:200-207), the
    # as-built rounding constant 2^14 at a >>14 shift gives 1 (window_multiplier.vhd:146-149).
    y, _ = fx.window(np.zeros(n, dtype=np.int64), n)
    assert np.all(y == 1)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 4096])
def test_hamming_rom_no_rounding_tie(n):
    i = np.arange(n // 2, dtype=np.float64)
    v = (0.54 - 0.46 * np.cos(2 * np.pi * i / (n - 1))) * 32767.0
    frac = v - np.floor(v)
    assert np.all(np.abs(frac - 0.5) > 1e-9)      # integer(real) tie rule never matters
    rom = fx.hamming_rom(n)
    assert rom[0] == 2621 and rom.max() <= 32767 and np.all(np.diff(rom) >= 0)


# --- MTI: rtl/src/tb_doppler_notch.vhd:117-180 -------------------------------------------------
def test_mti_tb_checks():
    n = 32
    dc = np.full(n, 5000)
    y = fx.mti(dc, 2)
    assert y[0] == 5000 and np.all(y[1:] == 0)                 # DC rejected after first sample
    k = np.arange(n)
    tone = np.round(8000 * np.cos(2 * np.pi * 8 * k / n)).astype(np.int64)
    assert np.mean(np.abs(fx.mti(tone, 2))) >= 1000            # f = 8/32 passes
    assert np.mean(np.abs(fx.mti(dc, 2, bypass=True))) >= 5000  # bypass
    y3 = fx.mti(dc, 3)
    assert y3[0] == 5000 and y3[1] == -5000 and np.all(y3[2:] == 0)
    # tlast clears the state: two lines are independent
    two = np.stack([dc, dc])
    assert np.array_equal(fx.mti(two, 2)[1], y)
    # saturation
    assert fx.mti(np.array([-32768, 32767]), 2)[1] == 32767


# --- CFAR: rtl/src/tb_os_cfar_2d.vhd:42-75,123-209 --------------------------------------------
def test_cfar_tb_geometry():
    g = fx.CfarGeometry(3, 2, 1, 1)
    assert (g.n_ref, g.rank_idx) == (54, 40)
    g = fx.CfarGeometry(4, 4, 2, 1)                            # radar_core.vhd:376-382 defaults
    assert (g.win_lines, g.win_cols, g.half_lines, g.half_cols) == (11, 13, 5, 6)
    assert (g.n_ref, g.rank_idx) == (128, 96)
    g = fx.CfarGeometry(8, 8, 4, 2)                            # BASELINE config 4 "wide"
    assert (g.n_ref, g.rank_idx) == (480, 360)
    g = fx.CfarGeometry(4, 4, 2, 1, axes="as_named")
    assert (g.win_lines, g.win_cols) == (13, 11)


@pytest.mark.parametrize("edges", ["zero", "raster"])
def test_cfar_tb_two_targets(edges):
    m = st.cfar_tb_map()
    out = fx.os_cfar_2d(m, 3, 2, 1, 1, edges=edges)
    det = fx.detections_from_cfar(out)
    assert det.shape[0] >= 2                                   # T1 (:123-134)
    cells = {(int(r), int(d)) for r, d, _ in det}
    assert (30, 16) in cells and (50, 8) in cells
    assert out[30, 16] == 5000 and out[50, 8] == 5000
    # every detection belongs to one of the two 3x3 target blobs
    assert all(max(abs(r - 30), abs(d - 16)) <= 1 or max(abs(r - 50), abs(d - 8)) <= 1
               for r, d in cells)


def test_cfar_tb_noise_only_false_alarms():
    r = np.arange(64)[:, None]
    d = np.arange(32)[None, :]
    noise = (100 + ((r * 7 + d * 13) % 30)).astype(np.uint32)  # make_map without targets
    out = fx.os_cfar_2d(noise, 3, 2, 1, 1, edges="raster")
    interior = out[3:-3, 4:-4]
    assert np.count_nonzero(interior) <= 3                     # T2 (:145-156): <= 3 false alarms


def test_cfar_tb_fixed_scale():
    m = st.cfar_tb_map()
    for ovr in (1, 2, 7):                                      # scale_override (:191-192)
        out, dbg = fx.os_cfar_2d(m, 3, 2, 1, 1, scale_override=ovr, return_debug=True)
        assert np.all(dbg["scale"] == ovr)
        assert out[30, 16] == 5000


def test_cfar_stream_model_matches_closed_form():
    """The literal clock-by-clock model of os_cfar_2d.vhd equals the closed-form raster
    window, and output beat k is the decision for flat cell k+3."""
    m = st.cfar_tb_map()
    for cfg in ((3, 2, 1, 1), (2, 2, 1, 1), (1, 2, 1, 0)):
        beats = np.array(fx.os_cfar_2d_stream(m.reshape(-1), 32, *cfg, flush=32 * 8))
        ras = fx.os_cfar_2d(m, *cfg, edges="raster").reshape(-1)
        n = min(beats.size, ras.size - 3)
        assert n > 1900
        assert np.array_equal(beats[:n], ras[3:3 + n])


def test_cfar_stream_model_random_maps():
    rng = np.random.default_rng(7)
    for cfg in ((2, 1, 1, 1), (1, 1, 0, 0)):
        m = rng.integers(0, 2000, size=(20, 16)).astype(np.uint32)
        m[10, 8] = 60000
        beats = np.array(fx.os_cfar_2d_stream(m.reshape(-1), 16, *cfg, flush=16 * 8))
        ras = fx.os_cfar_2d(m, *cfg, edges="raster").reshape(-1)
        n = min(beats.size, ras.size - 3)
        assert np.array_equal(beats[:n], ras[3:3 + n])


def test_cfar_zero_and_raster_agree_in_interior():
    rng = np.random.default_rng(3)
    m = rng.integers(50, 400, size=(40, 48)).astype(np.uint32)
    a = fx.os_cfar_2d(m, edges="zero")
    b = fx.os_cfar_2d(m, edges="raster")
    assert np.array_equal(a[5:-5, 6:-6], b[5:-5, 6:-6])


def test_cfar_sort_free_identities():
    """The counting forms the CUDA kernel uses equal the sorted form."""
    rng = np.random.default_rng(11)
    m = rng.integers(0, 300, size=(24, 32)).astype(np.int64)
    g = fx.CfarGeometry()
    refs = fx.cfar_refs(m, g, "zero")
    xk = np.sort(refs, axis=-1)[..., g.rank_idx]
    for a in (0, 50, 150, 299):
        assert np.array_equal(xk > a, (refs <= a).sum(-1) <= g.rank_idx)
        assert np.array_equal(xk < a, (refs < a).sum(-1) >= g.rank_idx + 1)
    for s in (2, 4, 6):
        assert np.array_equal(m > s * xk, (refs * s < m[..., None]).sum(-1) >= g.rank_idx + 1)


# --- PRNG ------------------------------------------------------------------------------------
def test_vhdl_uniform_basic():
    u = st.VhdlUniform(1234, 1234)
    a = [u() for _ in range(4)]
    assert all(0.0 < v < 1.0 for v in a)
    u2 = st.VhdlUniform(1234, 1234)
    assert np.allclose(u2.draw(4), a)
    frames, targets = st.tb_radar_core_frames(n_frames=1, n_range=64, n_doppler=8)
    tr, td = targets[0]
    assert 50.0 <= tr <= 950.0 and 0.0 <= td <= 100.0          # tb_radar_core.vhd:87-88
    assert frames.shape == (1, 8, 64, 2) and frames.dtype == np.int16
