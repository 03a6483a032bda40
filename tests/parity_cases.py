"""Parity cases shared by the CPU emulation tests (tiny sizes, kernel logic) and the GPU tests
(the parity tests proper, through the C ABI of librdp.so).  Each case takes `make_core`, a
factory returning a RadarCore, and compares the CUDA path with the oracle on seeded inputs.

Tolerances (north_star): RD map max |err| <= 1e-4 of the map peak (we assert 2e-6, float32
FFT round-off); CFAR on an identical map: bit-exact; end-to-end detection list: identical
except cells the oracle enumerates as within the map error of the threshold."""
import numpy as np

from oracle import float_path as fp
from oracle import rtl_fixed as fx
from oracle import stimulus as st

MAP_TOL = 2e-6          # of map / stage peak; north_star allows 1e-4


def rel_err(a, b):
    return float(np.abs(np.asarray(a, np.complex128) - np.asarray(b, np.complex128)).max() / np.abs(b).max())


def det_cells(d):
    return {(int(x["cpi"]), int(x["channel"]), int(x["range"]), int(x["doppler"])) for x in d}


def expected_status(cube):
    """Status word per CPI of a cube the detection list does not overflow on: RDP_STATUS_ADC_FULL_SCALE (4) for an int16
    CPI with a sample on a rail, else 0."""
    cube = np.asarray(cube)
    if cube.dtype != np.int16:
        return [0] * cube.shape[0]
    return [4 if bool((c == 32767).any() or (c == -32768).any()) else 0 for c in cube]


def check_canonical(dets):
    key = (dets["cpi"].astype(np.int64) << 48) | (dets["channel"].astype(np.int64) << 40) | \
          (dets["range"].astype(np.int64) << 20) | dets["doppler"].astype(np.int64)
    assert np.all(np.diff(key) > 0), "detection list not in canonical (cpi, channel, range, doppler) order"


def case_range_stage(make_core, n_samples, n_chirps, n_ch=2, fmt="int16", seed=3):
    cube, _ = st.synthetic_cube(2, n_ch, n_chirps, n_samples, seed=seed, dtype=fmt, n_targets=(2, 4))
    with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=n_ch, input_format=fmt, max_cpis=2) as c:
        got = c.range_stage(cube)
    ref = fp.range_stage(cube)
    assert got.shape == ref.shape == (2, n_ch, n_samples, n_chirps)
    assert rel_err(got, ref) <= MAP_TOL


def case_corner_turn_order(make_core, n_samples=16, n_chirps=8):
    """tb_corner_turner.vhd:146-186 analogue: an impulse of amplitude (chirp+1) at sample 0 makes the
    range spectrum of chirp c the constant (c+1)*w[0]; out[r][d] must be in[chirp d], range slow."""
    cube = np.zeros((n_chirps, n_samples, 2), np.int16)
    cube[:, 0, 0] = np.arange(1, n_chirps + 1) * 100
    with make_core(n_range=n_samples, n_doppler=n_chirps) as c:
        got = c.range_stage(cube)[0, 0]
    w0 = float(fp.hamming(n_samples)[0])
    exp = np.tile((np.arange(1, n_chirps + 1) * 100 * w0)[None, :], (n_samples, 1))
    assert np.allclose(got.real, exp, rtol=1e-6) and np.abs(got.imag).max() <= 1e-3


def case_doppler_stage(make_core, n_samples, n_chirps, mti_mode=0, mag_mode="amb", seed=4, window_gain=1.0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((1, 2, n_samples, n_chirps)) + 1j * rng.standard_normal((1, 2, n_samples, n_chirps)))
    x = (x * 1000).astype(np.complex64)
    x[0, 0, n_samples // 3] += 40000 * np.exp(2j * np.pi * 5 * np.arange(n_chirps) / n_chirps).astype(np.complex64)
    with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=2, mti_mode=mti_mode, mag_mode=mag_mode,
                   window_gain=window_gain) as c:
        got = c.doppler_stage(x)
    ref = fp.doppler_stage(x, mti_mode, mag_mode, window_gain)
    if mag_mode == "logpow":                              # compare in the linear domain (dB of a null is ill-conditioned)
        assert rel_err(10.0 ** (got.astype(np.float64) / 10.0), 10.0 ** (ref / 10.0)) <= 1e-5
        strong = ref >= ref.max() - 60.0
        assert np.abs(got - ref)[strong].max() <= 1e-2
    elif mag_mode == "pow":
        assert rel_err(got, ref) <= 4 * MAP_TOL
    else:
        assert rel_err(got, ref) <= MAP_TOL


def rayleigh_map(n_range, n_doppler, seed=1, n_maps=2, targets=3):
    rng = np.random.default_rng(seed)
    m = rng.rayleigh(100.0, size=(n_maps, n_range, n_doppler)).astype(np.float32)
    for k in range(targets):
        r, d = int(rng.integers(0, n_range)), int(rng.integers(0, n_doppler))
        m[k % n_maps, r, d] = 5000.0
        m[k % n_maps, max(r - 1, 0), d] = 1800.0
    m[0, 0, 0] = 4000.0            # corner and edge cells exercise the edge rule
    m[0, n_range - 1, n_doppler - 1] = 4000.0
    m[0, n_range // 2, 0] = 3000.0
    return m


def case_cfar_stage(make_core, n_range, n_doppler, seed=1, **kw):
    """CFAR stage on an identical float32 map: the dense RTL-style output and the detection list
    must be BIT-EXACT against the oracle."""
    m = rayleigh_map(n_range, n_doppler, seed)
    kw.setdefault("max_dets_per_cpi", 1 << 16)
    with make_core(n_range=n_range, n_doppler=n_doppler, n_channels=2, **kw) as c:
        dets, counts, status, dense = c.cfar_stage(m[None])
    okw = {k: v for k, v in kw.items() if k in fp.DEFAULTS}
    exp = np.stack([fp.cfar_stage(m[i], **okw) for i in range(2)])
    assert np.array_equal(dense[0], exp)
    ref = np.concatenate([fp.detections(exp[i], 0, i) for i in range(2)])
    assert counts.tolist() == [len(ref)] and status.tolist() == [0]
    assert len(ref) > 0
    for f in ("cpi", "channel", "range", "doppler", "mag"):
        assert np.array_equal(dets[f], ref[f]), f
    check_canonical(dets)


def case_cfar_tb_map(make_core):
    """rtl/src/tb_os_cfar_2d.vhd:42-75 map, (3,2,1,1): the float CUDA CFAR with a fixed scale equals
    the bit-exact integer restatement of os_cfar_2d.vhd on this integer-valued map."""
    m = st.cfar_tb_map()
    for edges in ("zero", "raster"):
        for ovr in (2, 4):
            with make_core(n_range=64, n_doppler=32, cfar_ref_r=3, cfar_ref_d=2, cfar_guard_r=1, cfar_guard_d=1,
                           scale_override=ovr, cfar_edges=edges) as c:
                dets, counts, status, dense = c.cfar_stage(m.astype(np.float32)[None, None])
            exp = fx.os_cfar_2d(m, 3, 2, 1, 1, scale_override=ovr, edges=edges)
            assert np.array_equal(dense[0, 0].astype(np.uint32), exp)
            cells = {(int(r), int(d)) for r, d in zip(dets["range"], dets["doppler"])}
            assert (30, 16) in cells and (50, 8) in cells           # T1: both targets found (:123-134)


def case_end_to_end(make_core, n_cpi, n_ch, n_chirps, n_samples, seed=11, fmt="int16", **kw):
    cube, _ = st.synthetic_cube(n_cpi, n_ch, n_chirps, n_samples, seed=seed, dtype=fmt)
    kw.setdefault("max_dets_per_cpi", 1 << 18)
    with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=n_ch, input_format=fmt, max_cpis=n_cpi, **kw) as c:
        rd_map, dets, n_dets, status = (x.copy() for x in c.process(cube))
        launches = c.info()["kernel_launches"]
    okw = {k: v for k, v in kw.items() if k in fp.DEFAULTS}
    ref_map, ref_dets, ref_counts = fp.process_batch(cube, **okw)
    err = float(np.abs(rd_map - ref_map).max())
    assert err <= MAP_TOL * float(ref_map.max())
    assert status.tolist() == expected_status(cube) and launches > 0
    check_canonical(dets)
    assert int(n_dets.sum()) == len(dets)
    eps = max(4.0 * err, 1e-7 * float(ref_map.max()))
    diff = det_cells(dets) ^ det_cells(ref_dets)
    amb_maps = {}
    for (i, ch, r, d) in diff:
        if (i, ch) not in amb_maps:
            amb_maps[(i, ch)] = fp.ambiguous_cells(ref_map[i, ch], eps, **okw)
        assert amb_maps[(i, ch)][r, d], f"detection mismatch at {(i, ch, r, d)} is not near-threshold"
    assert len(diff) <= max(2, len(ref_dets) // 500)                    # observed on B200: 0 (bench parity_spot, smoke)
    assert len(ref_dets) > 0
    return len(dets), len(diff)


def case_cfar_signed_map(make_core, n_range=64, n_doppler=64, seed=5, **kw):
    """Maps with NEGATIVE words (a dB map below 0 dB, a signed caller map): the packed pre-filters assume
    magnitudes >= 0, so such tiles take the exact sign-safe path -- the dense output and the list must still be
    bit-exact against the oracle, including tiles that only see a negative word in their halo."""
    rng = np.random.default_rng(seed)
    m = np.empty((2, n_range, n_doppler), np.float32)
    m[0] = (20.0 * np.log10(rng.rayleigh(1e-4, size=(n_range, n_doppler)))).astype(np.float32)   # all below 0 dB
    m[0, n_range // 3, n_doppler // 3] = -20.0
    m[1] = rng.rayleigh(60.0, size=(n_range, n_doppler)).astype(np.float32)          # ordinary map ...
    m[1, 5, 7] = 900.0
    m[1, n_range - 9, n_doppler - 4] = 1500.0
    m[1, n_range // 2, n_doppler // 2] = -3.0                                          # ... with one negative word,
    m[1, 2, 3] = -0.0                                                                  # a negative zero
    m[1, 10:14, 20] = 0.0                                                              # and exact zeros
    kw.setdefault("max_dets_per_cpi", 2 * n_range * n_doppler)                          # nearly every cell of the dB map is a "detection"
    with make_core(n_range=n_range, n_doppler=n_doppler, n_channels=2, **kw) as c:
        dets, counts, status, dense = c.cfar_stage(m[None])
    okw = {k: v for k, v in kw.items() if k in fp.DEFAULTS}
    exp = np.stack([fp.cfar_stage(m[i], **okw) for i in range(2)])
    assert np.array_equal(dense[0], exp)
    ref = np.concatenate([fp.detections(exp[i], 0, i) for i in range(2)])
    assert counts.tolist() == [len(ref)] and len(ref) > 0
    for f in ("channel", "range", "doppler", "mag"):
        assert np.array_equal(dets[f], ref[f]), f
    assert int((ref["mag"] < 0).sum()) > 0                                             # negative detections exist and match


def case_cfar_tiny_magnitudes(make_core, n_range=64, n_doppler=64, seed=13, **kw):
    """Magnitudes in the bfloat16 / float32 SUBNORMAL range and below bfloat16 resolution (< 9.2e-41): the packed
    pre-filters work on bfloat16 truncations (which honour subnormals, tools/microbench/mb_bf16_subnormal.cu) and
    take their cuts from them; truncations of 0 fall back to the exact words.  Bit-exact against the oracle."""
    rng = np.random.default_rng(seed)
    m = np.zeros((2, n_range, n_doppler), np.float32)
    m[0] = (rng.rayleigh(60.0, size=(n_range, n_doppler)) * 1e-41).astype(np.float32)     # ~6e-40: subnormal everywhere
    m[0, 20, 30] = np.float32(5e-38)
    m[0, 21, 30] = np.float32(2e-38)
    m[0, 40, 5] = np.float32(3e-39)
    m[1, 10, 10] = np.float32(3e-44)                                  # below bfloat16 resolution, alone among zeros: a detection
    m[1, 30:34, 40:44] = np.float32(1.4e-45)                          # the smallest float32 magnitudes
    m[1, 50, 20] = np.float32(7.0)                                    # and an ordinary cell on the same map
    kw.setdefault("max_dets_per_cpi", 1 << 14)
    with make_core(n_range=n_range, n_doppler=n_doppler, n_channels=2, **kw) as c:
        dets, counts, status, dense = c.cfar_stage(m[None])
    okw = {k: v for k, v in kw.items() if k in fp.DEFAULTS}
    exp = np.stack([fp.cfar_stage(m[i], **okw) for i in range(2)])
    assert np.array_equal(dense[0], exp)
    assert counts[0] == np.count_nonzero(exp) and exp[1, 10, 10] != 0 and exp[0, 20, 30] != 0


def case_cfar_crowded_tile(make_core, n_range=64, n_doppler=128, seed=9, **kw):
    """More surviving strips in one 64 x 64 tile than one phase-1 round takes (128) and more survivors than the
    list holds at once: a lattice of strong isolated cells.  Bit-exact against the oracle."""
    rng = np.random.default_rng(seed)
    m = rng.rayleigh(40.0, size=(2, n_range, n_doppler)).astype(np.float32)
    m[0, ::4, ::3] = rng.uniform(2000.0, 9000.0, size=m[0, ::4, ::3].shape).astype(np.float32)
    m[1, ::2, ::2] *= 40.0
    kw.setdefault("max_dets_per_cpi", 1 << 16)
    with make_core(n_range=n_range, n_doppler=n_doppler, n_channels=2, **kw) as c:
        dets, counts, status, dense = c.cfar_stage(m[None])
    okw = {k: v for k, v in kw.items() if k in fp.DEFAULTS}
    exp = np.stack([fp.cfar_stage(m[i], **okw) for i in range(2)])
    assert np.array_equal(dense[0], exp)
    assert counts[0] == np.count_nonzero(exp) and counts[0] > 300


def case_end_to_end_mag_modes(make_core, n_ch=2, n_chirps=32, n_samples=64, seed=23):
    """Whole path with mag_mode pow and logpow (north_star: |X|^2 and the log magnitude), CFAR included, against the
    oracle -- once on an int16 cube (dB values far above 0) and once on a complex64 cube so weak that |X|^2 < 1
    (every dB value negative: the sign-safe CFAR path).  CFAR on a dB map keeps the reference's MULTIPLICATIVE rule
    (os_cfar_2d.vhd:204,213 knows no other); it is a defined function of the map, not a sensible detector."""
    out = {}
    for fmt, scale in (("int16", 1.0), ("complex64", 1e-9)):
        cube, _ = st.synthetic_cube(2, n_ch, n_chirps, n_samples, seed=seed, dtype=fmt, sigma=30.0, n_targets=(2, 3))
        if fmt == "complex64":
            cube = (cube * np.float32(scale)).astype(np.complex64)
        for mode in ("pow", "logpow"):
            with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=n_ch, input_format=fmt, max_cpis=2,
                           mag_mode=mode, max_dets_per_cpi=n_ch * n_chirps * n_samples) as c:   # a dB map below 0 dB "detects" nearly everywhere
                rd_map, dets, n_dets, status = (x.copy() for x in c.process(cube))
            ref_map, ref_dets, ref_counts = fp.process_batch(cube, mag_mode=mode)
            if mode == "logpow":
                # compare in the linear domain (the dB value of a near-null cell is ill-conditioned) ...
                assert rel_err(10.0 ** (rd_map.astype(np.float64) / 10.0), 10.0 ** (ref_map.astype(np.float64) / 10.0)) <= 1e-5
                err = float(np.abs(rd_map - ref_map).max())                            # ... and bound the dB error itself
                assert err <= 0.5, err
                strong = ref_map >= ref_map.max() - 60.0
                assert float(np.abs(rd_map - ref_map)[strong].max()) <= 1e-2
                if scale != 1.0:
                    assert float(ref_map.max()) < 0.0                                  # the all-negative case really is one
            else:
                err = float(np.abs(rd_map - ref_map).max())
                assert err <= 4 * MAP_TOL * float(ref_map.max())
            # the CFAR stage on the GPU's own map is bit-exact against the oracle's CFAR on that same map
            exp = np.concatenate([fp.detections(fp.cfar_stage(rd_map[i, ch]), i, ch) for i in range(2) for ch in range(n_ch)])
            assert len(exp) == len(dets) and status.tolist() == expected_status(cube), (fmt, mode)
            # (a dB map far above 0 dB has no cell 2..6 TIMES its neighbours: the multiplicative rule finds nothing there)
            assert len(dets) > 0 or (mode == "logpow" and scale == 1.0), (fmt, mode)
            for f in ("cpi", "channel", "range", "doppler", "mag"):
                assert np.array_equal(dets[f], exp[f]), (fmt, mode, f)
            # and the list from the oracle's map differs only in near-threshold cells
            eps = max(4.0 * err, 1e-7 * float(np.abs(ref_map).max()))
            diff = det_cells(dets) ^ det_cells(ref_dets)
            amb = {}
            for (i, ch, r, d) in diff:
                if (i, ch) not in amb:
                    amb[(i, ch)] = fp.ambiguous_cells(ref_map[i, ch], eps, mag_mode=mode)
                assert amb[(i, ch)][r, d], (fmt, mode, i, ch, r, d)
            out[(fmt, mode)] = (len(dets), len(diff))
    return out


def case_overflow_flag(make_core):
    m = rayleigh_map(64, 64, seed=5, n_maps=1)
    m[0, 10:30, 10:30:3] = 9000.0                          # many isolated strong cells
    with make_core(n_range=64, n_doppler=64, max_dets_per_cpi=8, scale_override=1) as c:
        dets, counts, status, _ = c.cfar_stage(m[None], want_dense=False)
    assert counts[0] > 8 and status[0] & 1 and len(dets) == 8
    check_canonical(dets)
    ref = fp.detections(fp.cfar_stage(m[0], scale_override=1))
    assert counts[0] == len(ref)
    for f in ("range", "doppler", "mag"):                 # truncation keeps the FIRST records in canonical order
        assert np.array_equal(dets[f], ref[f][:8])


def case_chunking_invariance(make_core, n_samples=64, n_chirps=32):
    cube, _ = st.synthetic_cube(3, 2, n_chirps, n_samples, seed=21, sigma=25.0, n_targets=(1, 3))
    out = []
    for chunk in (1, 2, 3):
        with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=2, max_cpis=3, chunk_cpis=chunk) as c:
            assert c.info()["chunk_cpis"] == chunk
            out.append(tuple(x.copy() for x in c.process(cube)))
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0])
        assert np.array_equal(o[1], out[0][1]) and np.array_equal(o[2], out[0][2])


def case_golden_chirp(make_core, golden_chirp_1024, n_chirps=16):
    """data/golden_input_chirp.txt[0:1024] replicated over the chirps with a Doppler ramp of bin 3:
    the RD map must peak at range bin 290 (0.283333 * 1024), Doppler bin 3."""
    base = golden_chirp_1024[:, 0].astype(np.float64) + 1j * golden_chirp_1024[:, 1]
    ramp = np.exp(2j * np.pi * 3 * np.arange(n_chirps) / n_chirps)[:, None]
    z = 0.5 * base[None, :] * ramp
    cube = np.stack([np.round(z.real), np.round(z.imag)], axis=-1).astype(np.int16)
    with make_core(n_range=1024, n_doppler=n_chirps) as c:
        rd_map, dets, n_dets, status = (x.copy() for x in c.process(cube))
    ref_map, ref_dets = fp.process_cpi(cube)
    r, d = np.unravel_index(np.argmax(rd_map[0, 0]), rd_map[0, 0].shape)
    assert (int(r), int(d)) == (290, 3)
    assert float(np.abs(rd_map[0] - ref_map).max()) <= MAP_TOL * float(ref_map.max())
    assert (0, 0, 290, 3) in det_cells(dets)


def case_cfar_fixed_point(make_core, n_range=64, n_doppler=64, seed=3, **kw):
    """Fixed-point CFAR stage (rdp_cfar_stage_u32) against the bit-exact integer restatement of
    rtl/src/os_cfar_2d.vhd (oracle.rtl_fixed.os_cfar_2d): adaptive scale, floor division, 17-bit
    wrap and all -- the dense RTL output word and the detection list must be identical."""
    rng = np.random.default_rng(seed)
    m = rng.integers(20, 600, size=(2, n_range, n_doppler)).astype(np.uint32)
    m[0, n_range // 2, n_doppler // 2] = 45056                   # largest 17-bit magnitude (magnitude_calc.vhd:78-81)
    m[0, n_range // 2, n_doppler // 2 + 1] = 20000
    m[1, 3, 5] = 9000
    m[1, :, 7] += 2500                                           # a ridge: windows whose mean is dominated by a few cells
    m[0, 0, 0] = 7000
    m[0, n_range - 1, n_doppler - 1] = 7000
    kw.setdefault("max_dets_per_cpi", 1 << 16)
    with make_core(n_range=n_range, n_doppler=n_doppler, n_channels=2, **kw) as c:
        dets, counts, status, dense = c.cfar_stage_fixed(m[None])
    names = dict(cfar_ref_r="ref_range", cfar_ref_d="ref_doppler", cfar_guard_r="guard_range", cfar_guard_d="guard_doppler",
                 rank_pct="rank_pct", scale_min="scale_min", scale_nom="scale_nom", scale_max="scale_max",
                 scale_override="scale_override", cfar_edges="edges", cfar_axes="axes")
    okw = {names[k]: v for k, v in kw.items() if k in names}
    exp = np.stack([fx.os_cfar_2d(m[i], **okw) for i in range(2)])
    assert np.array_equal(dense[0], exp)
    n_exp = int(np.count_nonzero(exp))
    assert counts.tolist() == [n_exp] and n_exp > 0 and status.tolist() == [0]
    ref = np.concatenate([fx.detections_from_cfar(exp[i]) for i in range(2)])
    assert np.array_equal(dets["range"], ref[:, 0]) and np.array_equal(dets["doppler"], ref[:, 1])
    assert np.array_equal(dets["mag"].astype(np.int64), ref[:, 2])
    check_canonical(dets)


def case_cfar_fixed_range_flag(make_core):
    """rdp_cfar_stage_u32's hard precondition (words < 2^24) is checked on the device: RDP_STATUS_MAP_RANGE for exactly
    the CPIs that break it."""
    rng = np.random.default_rng(2)
    m = rng.integers(0, 5000, size=(3, 1, 64, 64)).astype(np.uint32)
    m[1, 0, 40, 7] = 1 << 24
    m[2, 0, 0, 0] = (1 << 24) - 1
    with make_core(n_range=64, n_doppler=64, max_cpis=3, max_dets_per_cpi=4096) as c:
        _, _, status, _ = c.cfar_stage_fixed(m)
    assert [int(x) & 8 for x in status] == [0, 8, 0]


def case_cfar_fixed_tb_map(make_core):
    """rtl/src/tb_os_cfar_2d.vhd:42-75 map, (3,2,1,1), adaptive scale: bit-exact against the integer oracle,
    both targets detected (T1, :123-134)."""
    m = st.cfar_tb_map()
    for edges in ("zero", "raster"):
        with make_core(n_range=64, n_doppler=32, cfar_ref_r=3, cfar_ref_d=2, cfar_guard_r=1, cfar_guard_d=1,
                       cfar_edges=edges) as c:
            dets, counts, status, dense = c.cfar_stage_fixed(m[None, None])
        exp = fx.os_cfar_2d(m, 3, 2, 1, 1, edges=edges)
        assert np.array_equal(dense[0, 0], exp)
        cells = {(int(r), int(d)) for r, d in zip(dets["range"], dets["doppler"])}
        assert (30, 16) in cells and (50, 8) in cells and counts[0] >= 2


def case_streaming(make_core, n_samples=64, n_chirps=32, n_sub=7, n_slots=3, spec=None):
    """Streaming ingest (BASELINE config 5): results of pipelined single-CPI submissions equal the batch API's.
    `spec` shrinks the speculatively copied head of the list so that the second copy in rdp_stream_wait runs."""
    import os
    if spec is not None:
        os.environ["RDP_STREAM_SPEC"] = str(spec)
    try:
        _streaming(make_core, n_samples, n_chirps, n_sub, n_slots)
    finally:
        os.environ.pop("RDP_STREAM_SPEC", None)


def _streaming(make_core, n_samples, n_chirps, n_sub, n_slots):
    cube, _ = st.synthetic_cube(n_sub, 2, n_chirps, n_samples, seed=41, sigma=25.0, n_targets=(1, 3))
    with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=2, max_cpis=n_sub, max_dets_per_cpi=4096) as c:
        ref, ref_counts = (x.copy() for x in c.process(cube, want_map=False)[1:3])
        c.stream_open(n_slots, 1)
        bufs = [c.pinned_empty(cube[0:1].shape, cube.dtype) for _ in range(n_slots)]
        got = [None] * n_sub
        for i in range(n_sub + n_slots):
            slot = i % n_slots
            if i >= n_slots:
                d, n, s = c.stream_wait(slot)
                got[i - n_slots] = (d.copy(), int(n[0]), int(s[0]))
            if i < n_sub:
                bufs[slot][...] = cube[i:i + 1]
                c.stream_submit(slot, bufs[slot])
        c.stream_close()
    for i in range(n_sub):
        d, n, s = got[i]
        exp = ref[ref["cpi"] == i]
        assert n == ref_counts[i] == len(exp) and s == 0
        for f in ("channel", "range", "doppler", "mag"):
            assert np.array_equal(d[f], exp[f])
        assert np.all(d["cpi"] == 0)


def case_fixed_stages(make_core):
    """16-bit fixed-point stage taps against the bit-exact oracle AND the reference's own testbench vectors."""
    from test_oracle_fixed import MAG_VECTORS, MAG_EXPECTED
    rng = np.random.default_rng(17)
    with make_core(n_range=64, n_doppler=32) as c:
        # magnitude: the 17 known answers of rtl/src/tb_magnitude_calc.vhd:49-73, exact (+ extremes)
        vec = np.array(MAG_VECTORS + [(-32768, -32768), (-32768, 0)], dtype=np.int16)
        got = c.fixed_magnitude(vec)
        assert got.tolist() == MAG_EXPECTED + [32768 + 8192 + 4096, 32768]
        z = rng.integers(-32768, 32768, size=(1000, 2)).astype(np.int16)
        assert np.array_equal(c.fixed_magnitude(z), fx.magnitude(z[:, 0], z[:, 1]))
        # corner turn: rtl/src/tb_corner_turner.vhd:36-39,146-186 (16 x 8, value = chirp * 256 + sample)
        chirp, sample = np.arange(8)[:, None], np.arange(16)[None, :]
        val = (chirp * 256 + sample).astype(np.int16)
        x = np.stack([val, -val], axis=-1)[None]
        y = c.fixed_corner_turn(x)
        assert y.shape == (1, 16, 8, 2)
        flat = y[0, :, :, 0].reshape(-1).astype(np.int64)
        assert np.array_equal(flat // 256, np.tile(np.arange(8), 16)) and np.array_equal(flat % 256, np.repeat(np.arange(16), 8))
        big = rng.integers(-32768, 32768, size=(3, 70, 45, 2)).astype(np.int16)
        assert np.array_equal(c.fixed_corner_turn(big), np.swapaxes(big, 1, 2))
        # window: rtl/src/tb_window_multiplier.vhd:83-108,185-240 (N = 64) + random rows vs the oracle
        dc = np.full((1, 64, 2), 16000, np.int16)
        y, sat = c.fixed_window(dc)
        assert not sat and abs(int(y[0, 0, 0])) <= 3000 and int(y[0, 32, 0]) >= 10000
        assert np.all(np.abs(y[0, :, 0].astype(int) - y[0, ::-1, 0].astype(int)) <= 1)
        y, sat = c.fixed_window(np.full((1, 64, 2), 32767, np.int16))
        assert sat and int(y.max()) == 32767
        y, sat = c.fixed_window(np.zeros((1, 64, 2), np.int16))
        assert np.all(y == 1)                                # as-built rounding constant (window_multiplier.vhd:146-149)
        for n in (64, 256, 1024):
            w = rng.integers(-32768, 32768, size=(5, n, 2)).astype(np.int16)
            y, sat = c.fixed_window(w)
            ei, si = fx.window(w[..., 0], n)
            eq, sq = fx.window(w[..., 1], n)
            assert np.array_equal(y[..., 0], ei) and np.array_equal(y[..., 1], eq) and sat == bool(si.any() or sq.any())
        # MTI: rtl/src/tb_doppler_notch.vhd:117-180 + random lines vs the oracle
        line = np.full((2, 32, 2), 5000, np.int16)
        y = c.fixed_mti(line, 2)
        assert np.all(y[:, 0] == 5000) and np.all(y[:, 1:] == 0)
        for mode in (0, 2, 3):
            w = rng.integers(-32768, 32768, size=(7, 33, 2)).astype(np.int16)
            y = c.fixed_mti(w, mode)
            assert np.array_equal(y[..., 0], fx.mti(w[..., 0], mode)) and np.array_equal(y[..., 1], fx.mti(w[..., 1], mode))


def case_cfar_random_config(make_core, seed):
    """Randomised CFAR generics (window extents, guard, rank, scales, override, edge rule, axis convention) on a
    random map: float CUDA CFAR bit-exact against the float oracle, fixed-point CUDA CFAR bit-exact against the
    integer oracle whenever the window shape has a tiled kernel."""
    rng = np.random.default_rng(seed)
    tiled = [(4, 4, 2, 1), (3, 2, 1, 1), (2, 2, 1, 1), (8, 8, 4, 2)]
    if rng.random() < 0.5:
        rr, rd, gr, gd = tiled[int(rng.integers(0, len(tiled)))]
    else:
        rr, rd, gr, gd = int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(0, 3)), int(rng.integers(0, 3))
    smin = int(rng.integers(1, 4))
    kw = dict(cfar_ref_r=rr, cfar_ref_d=rd, cfar_guard_r=gr, cfar_guard_d=gd, rank_pct=int(rng.integers(10, 100)),
              scale_min=smin, scale_nom=smin + int(rng.integers(0, 3)), scale_max=smin + int(rng.integers(2, 5)),
              scale_override=int(rng.choice([0, 0, 0, 1, 3, 7])), cfar_edges=str(rng.choice(["zero", "raster"])),
              cfar_axes=str(rng.choice(["as_built", "as_named"])))
    nr, nd = int(rng.choice([8, 16, 32, 64])), int(rng.choice([8, 16, 32, 64, 128]))
    m = rng.rayleigh(60.0, size=(2, nr, nd)).astype(np.float32)
    for _ in range(3):
        m[int(rng.integers(0, 2)), int(rng.integers(0, nr)), int(rng.integers(0, nd))] = float(rng.uniform(500, 5000))
    m[0, 0, :] *= 3.0
    with make_core(n_range=nr, n_doppler=nd, n_channels=2, max_dets_per_cpi=1 << 15, **kw) as c:
        dets, counts, status, dense = c.cfar_stage(m[None])
        exp = np.stack([fp.cfar_stage(m[i], **kw) for i in range(2)])
        assert np.array_equal(dense[0], exp), kw
        assert counts[0] == np.count_nonzero(exp)
        if (rr, rd, gr, gd) in tiled:
            mi = np.minimum(m * 8.0, 131071.0).astype(np.uint32)
            d2, c2, s2, dense_i = c.cfar_stage_fixed(mi[None])
            names = dict(cfar_ref_r="ref_range", cfar_ref_d="ref_doppler", cfar_guard_r="guard_range",
                         cfar_guard_d="guard_doppler", cfar_edges="edges", cfar_axes="axes")
            okw = {names.get(k, k): v for k, v in kw.items()}
            exp_i = np.stack([fx.os_cfar_2d(mi[i], **okw) for i in range(2)])
            assert np.array_equal(dense_i[0], exp_i), kw


def case_cfar_adaptive_stress(make_core, seed):
    """The adaptive pre-filter threshold (window-sum bound, half-way exit) must never lose a detection: maps full
    of ridges and isolated strong cells (windows whose mean is dominated by a few references, i.e. SCALE_MIN
    territory), scales in any order, tiny / zero / huge magnitudes.  Bit-exact against the float oracle."""
    rng = np.random.default_rng(1000 + seed)
    geo = [(4, 4, 2, 1), (3, 2, 1, 1), (2, 2, 1, 1), (8, 8, 4, 2)][seed % 4]
    scales = [int(v) for v in rng.permutation([int(rng.integers(1, 3)), int(rng.integers(3, 5)), int(rng.integers(5, 8))])]
    kw = dict(cfar_ref_r=geo[0], cfar_ref_d=geo[1], cfar_guard_r=geo[2], cfar_guard_d=geo[3],
              rank_pct=int(rng.choice([25, 50, 75, 90])), scale_min=scales[0], scale_nom=scales[1], scale_max=scales[2],
              cfar_axes=str(rng.choice(["as_built", "as_named"])), cfar_edges=str(rng.choice(["zero", "raster"])))
    nr, nd = int(rng.choice([32, 64, 128])), int(rng.choice([64, 128, 256]))
    m = rng.rayleigh(30.0, size=(2, nr, nd)).astype(np.float32)
    for _ in range(6):                                               # ridges along either axis, of very different strength
        a = float(10.0 ** rng.uniform(1.5, 4.5))
        if rng.random() < 0.5:
            m[int(rng.integers(0, 2)), int(rng.integers(0, nr)), :] += a * rng.random(nd).astype(np.float32)
        else:
            m[int(rng.integers(0, 2)), :, int(rng.integers(0, nd))] += a * rng.random(nr).astype(np.float32)
    for _ in range(12):                                              # isolated strong cells and small plateaus
        ch, r, d = int(rng.integers(0, 2)), int(rng.integers(0, nr)), int(rng.integers(0, nd))
        m[ch, r:r + int(rng.integers(1, 4)), d:d + int(rng.integers(1, 4))] = float(10.0 ** rng.uniform(2.5, 6.0))
    if seed % 3 == 0:
        m[0] *= np.float32(1e-36)                                    # down into the bfloat16 subnormal range
        m[1, : nr // 2] = 0.0                                        # a block of exact zeros
    with make_core(n_range=nr, n_doppler=nd, n_channels=2, max_dets_per_cpi=1 << 15, **kw) as c:
        dets, counts, status, dense = c.cfar_stage(m[None])
        exp = np.stack([fp.cfar_stage(m[i], **kw) for i in range(2)])
        assert np.array_equal(dense[0], exp), kw
        assert counts[0] == np.count_nonzero(exp) and counts[0] > 0


def _dets_from_scans(scans_per_channel):
    """scans_per_channel[cpi][channel] = list of (r, d, mag) in (range, doppler) order -> canonical list + counts."""
    import fpga_fmcw_radar_processor_b200 as rdp
    rows, counts = [], []
    for cpi, chans in enumerate(scans_per_channel):
        n = 0
        for ch, lst in enumerate(chans):
            for (r, d, mag) in lst:
                rows.append((cpi, ch, r, d, float(mag)))
                n += 1
        counts.append(n)
    dets = np.zeros(len(rows), dtype=rdp.DET_DTYPE)
    for i, r in enumerate(rows):
        dets[i] = r
    return dets, np.asarray(counts, np.int32)


def case_tracker(make_core):
    """tws_tracker on the device detection list against the bit-exact oracle (oracle/tracker.py): the scripted
    scenario of rtl/src/tb_tws_tracker.vhd (with its assertions), then random multi-channel scans including more
    than 64 detections per scan (the 6-bit det_count wrap) and state carried across two calls."""
    from oracle import tracker as tk
    from test_oracle_tracker import tb_scan_detections
    # --- the reference testbench scenario, generics of tb_tws_tracker.vhd:11-18
    scans = [[sorted(tb_scan_detections(s))] for s in range(1, 13)]
    dets, counts = _dets_from_scans(scans)
    with make_core(n_range=1024, n_doppler=128, max_cpis=12) as c:
        c.tracker_reset(max_tracks=16, coast_max=3)
        trk, nt, act = c.track(dets, counts)
    ref = tk.TwsTracker(max_tracks=16, coast_max=3)
    exp_rows, exp_act = [], []
    for cpi, s in enumerate(scans):
        out, a = ref.scan(s[0])
        exp_act.append(a)
        exp_rows += [(cpi, 0, t["id"], t["range_q2"], t["doppler_q2"], t["vel_r"], t["vel_d"], t["quality"], t["status"]) for t in out]
    assert act[:, 0].tolist() == exp_act
    assert act[1, 0] >= 2 and act[2, 0] >= 2 and act[5, 0] >= 3                 # tb_tws_tracker.vhd:134-153
    assert [tuple(int(v) for v in t) for t in trk] == exp_rows and len(exp_rows) > 0
    # --- random scans, 2 channels, some with > 64 detections, two calls sharing the track files
    rng = np.random.default_rng(12)
    n_scan, n_ch = 10, 2
    scans = []
    for s in range(n_scan):
        chans = []
        for ch in range(n_ch):
            n = int(rng.choice([0, 1, 3, 8, 70, 130]))
            base = [(100 + 3 * s + 11 * k, 10 + k, 500 + k) for k in range(4)]       # slowly moving targets in range < 512, Doppler < 64
            rnd = [(int(rng.integers(0, 1024)), int(rng.integers(0, 128)), int(rng.integers(1, 60000))) for _ in range(n)]
            chans.append(sorted(set(base + rnd)))
        scans.append(chans)
    dets, counts = _dets_from_scans(scans)
    with make_core(n_range=1024, n_doppler=128, n_channels=n_ch, max_cpis=n_scan, max_dets_per_cpi=512) as c:
        c.tracker_reset()
        half = n_scan // 2
        cut = int(counts[:half].sum())
        d1 = dets[:cut].copy()
        d2 = dets[cut:].copy()
        d2["cpi"] -= half
        t1, nt1, a1 = c.track(d1, counts[:half])
        t2, nt2, a2 = c.track(d2, counts[half:])
    t2 = t2.copy()
    t2["cpi"] += half
    got = [tuple(int(v) for v in t) for t in np.concatenate([t1, t2])]
    act = np.concatenate([a1, a2])
    refs = [tk.TwsTracker() for _ in range(n_ch)]
    exp_rows = []
    for cpi, chans in enumerate(scans):
        for ch, lst in enumerate(chans):
            out, a = refs[ch].scan(lst)
            assert act[cpi, ch] == a
            exp_rows += [(cpi, ch, t["id"], t["range_q2"], t["doppler_q2"], t["vel_r"], t["vel_d"], t["quality"], t["status"]) for t in out]
    assert got == exp_rows and len(exp_rows) > 10


def case_sharded_cpi(make_sharded, make_core, world=4, n_chirps=64, n_samples=128, n_ch=2, n_cpi=2, **kw):
    """BASELINE config 4 in miniature: one CPI sharded over `world` ranks (simulated in ONE process: the ranks'
    phases are run in lockstep, peer buffers are plain pointers) -- chirp-sharded range FFT with the corner turn
    fused into peer stores, range-sharded Doppler FFT, CFAR with halo lines from the neighbours -- must give
    exactly what the single-device path gives on the whole CPI."""
    cube, _ = st.synthetic_cube(n_cpi, n_ch, n_chirps, n_samples, seed=77, sigma=30.0, n_targets=(2, 4))
    with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=n_ch, max_cpis=n_cpi, max_dets_per_cpi=1 << 16, **kw) as c:
        ref_map, ref_dets = (x.copy() for x in c.process(cube)[:2])
    ranks = [make_sharded(n_chirps=n_chirps, n_samples=n_samples, world=world, rank=r, n_channels=n_ch, n_cpi=n_cpi, **kw)
             for r in range(world)]
    try:
        handles = [s.export_handles() for s in ranks]
        for s in ranks:
            s.connect(handles)
        cl = n_chirps // world
        for r, s in enumerate(ranks):
            s.range_scatter(cube[:, :, r * cl:(r + 1) * cl])
        for s in ranks:
            s.doppler()
        for s in ranks:
            s.halo_exchange()
        dets = [s.cfar()[0] for s in ranks]
        maps = [s.local_map() for s in ranks]
    finally:
        for s in ranks:
            s.close()
    assert np.array_equal(np.concatenate(maps, axis=2), ref_map)                   # same kernels, same data: identical
    import fpga_fmcw_radar_processor_b200 as rdp
    merged = rdp.merge_detections(dets, [0] * world)
    assert len(merged) == len(ref_dets) and len(ref_dets) > 0
    for f in ("cpi", "channel", "range", "doppler", "mag"):
        assert np.array_equal(merged[f], ref_dets[f]), f


def case_bfp_fft(make_core, sizes=(2, 8, 64, 256, 1024), seed=31):
    """rdp_fixed_bfp_fft against oracle.rtl_fixed.bfp_fft word for word (mantissas and block exponents): full-scale
    random frames (every stage shifts), small ones (no stage shifts), an impulse, DC at the rail, all zeros."""
    rng = np.random.default_rng(seed)
    with make_core(n_range=64, n_doppler=32) as c:
        for n in sizes:
            frames = [rng.integers(-32768, 32768, size=(3, n, 2)), rng.integers(-40, 41, size=(2, n, 2)),
                      np.zeros((1, n, 2), np.int64), np.full((1, n, 2), 32767), np.full((1, n, 2), -32768)]
            imp = np.zeros((1, n, 2), np.int64)
            imp[0, 1 % n, 0] = 32767
            x = np.concatenate(frames + [imp]).astype(np.int16)
            y, e = c.fixed_bfp_fft(x)
            er, ei, ee = fx.bfp_fft(x[..., 0], x[..., 1])
            assert np.array_equal(y[..., 0], er) and np.array_equal(y[..., 1], ei), n
            assert np.array_equal(e, ee), n
            # and it IS a DFT: X * 2^e within the rounding noise of log2(n) 16-bit stages
            ref = np.fft.fft(x[..., 0].astype(np.float64) + 1j * x[..., 1].astype(np.float64), axis=-1)
            got = (y[..., 0].astype(np.float64) + 1j * y[..., 1].astype(np.float64)) * (2.0 ** e)[:, None]
            # rounding noise: <= 1/2 LSB (of the stage's scale) per butterfly, accumulated over the n terms of an output
            assert np.all(np.abs(got - ref).max(axis=-1) <= (2.0 ** e) * (2.0 * np.sqrt(n) + 4.0)), n


def case_fixed_end_to_end(make_core, n_cpi=2, n_ch=2, n_chirps=32, n_samples=64, mti_mode=0, seed=41, amp_scale=1.0, **kw):
    """arith="fixed": ADC words to detection list through the 16-bit chain, bit-exact against
    oracle.rtl_fixed.process_cpi_fixed -- magnitude words, both block exponents, the detection list, the status word
    (window saturation, ADC full scale) -- through rdp_process_fixed_device AND the host path (process)."""
    cube, _ = st.synthetic_cube(n_cpi, n_ch, n_chirps, n_samples, seed=seed, dtype="int16", n_targets=(2, 4))
    if amp_scale != 1.0:
        cube = np.clip(cube.astype(np.float64) * amp_scale, -32768, 32767).astype(np.int16)
    okw = {k: v for k, v in kw.items() if k not in ("max_dets_per_cpi", "chunk_cpis")}
    with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=n_ch, max_cpis=n_cpi, mti_mode=mti_mode, arith="fixed",
                   max_dets_per_cpi=kw.pop("max_dets_per_cpi", 1 << 14), **kw) as c:
        mag, er, ed, dets, counts, status = c.process_fixed(cube)
        m2, dets2, counts2, status2 = (x.copy() for x in c.process(cube))
    assert m2.dtype == np.uint32 and np.array_equal(m2, mag)
    assert np.array_equal(dets2, dets) and np.array_equal(counts2, counts) and np.array_equal(status2, status)
    check_canonical(dets)
    n_sat = 0
    for i in range(n_cpi):
        e_mag, e_dets, e_sat, e_er, e_ed = fx.process_cpi_fixed(cube[i], mti_mode=mti_mode, **okw)
        assert np.array_equal(mag[i], e_mag), i
        assert np.array_equal(er[i], e_er) and np.array_equal(ed[i], e_ed), i
        mine = dets[dets["cpi"] == i]
        assert len(mine) == len(e_dets) == counts[i], (i, len(mine), len(e_dets))
        assert np.array_equal(np.stack([mine["channel"], mine["range"], mine["doppler"]], axis=1), e_dets[:, :3])
        assert np.array_equal(mine["mag"].astype(np.int64), e_dets[:, 3])
        rail = bool((cube[i] == 32767).any() or (cube[i] == -32768).any())
        assert int(status[i]) == (2 if e_sat else 0) | (4 if rail else 0), (i, int(status[i]), e_sat, rail)
        n_sat += int(e_sat)
    return n_sat


def case_adc_full_scale_flag(make_core, n_samples=64, n_chirps=32):
    """float32 arithmetic, int16 input: RDP_STATUS_ADC_FULL_SCALE is raised for exactly the CPIs that hold a sample on
    a rail (either rail, I or Q, any channel), and never for complex64 input."""
    cube, _ = st.synthetic_cube(4, 2, n_chirps, n_samples, seed=5, dtype="int16", n_targets=(1, 2))
    cube = np.clip(cube, -20000, 20000)
    cube[1, 0, 3, 7, 0] = 32767
    cube[2, 1, n_chirps - 1, n_samples - 1, 1] = -32768
    with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=2, max_cpis=4, max_dets_per_cpi=1 << 14) as c:
        _, _, _, status = c.process(cube, want_map=False)
        assert [int(s) & 4 for s in status] == [0, 4, 4, 0]


def case_combine_and_group(make_core, n_cpi=2, n_ch=4, n_chirps=32, n_samples=64, fmt="int16", mti_mode=0, seed=51, **kw):
    """Optional stages in front of / behind the CFAR: non-coherent integration over the channels
    (one combined map per CPI, fused into the Doppler stage), coherent beams (steered channel sums in front of the
    Doppler transform = an angle DFT over the channels, one map per beam) and peak grouping of the detections.
    Combined map within MAP_TOL of the oracle's; CFAR (+ grouping) on the GPU's own map bit-exact against the
    oracle's CFAR (+ grouping) on that same map; end-to-end list identical outside near-threshold cells."""
    cube, _ = st.synthetic_cube(n_cpi, n_ch, n_chirps, n_samples, seed=seed, dtype=fmt, n_targets=(2, 4))
    n_beams = 1
    while n_beams < n_ch:
        n_beams <<= 1
    for combine, group in (("nci", False), ("nci", True), ("none", True), ("beams", False), ("beams", True)):
        with make_core(n_range=n_samples, n_doppler=n_chirps, n_channels=n_ch, input_format=fmt, max_cpis=n_cpi,
                       mti_mode=mti_mode, combine=combine, peak_group=group, max_dets_per_cpi=1 << 16, **kw) as c:
            rd_map, dets, n_dets, status = (x.copy() for x in c.process(cube))
            tap = c.doppler_stage(c.range_stage(cube))                          # the stage tap combines too
            d_tap, c_tap, s_tap, dense = c.cfar_stage(rd_map)                    # CFAR stage tap on the combined maps
        n_maps = {"nci": 1, "beams": n_beams}.get(combine, n_ch)
        assert rd_map.shape == (n_cpi, n_maps, n_samples, n_chirps) and np.array_equal(tap, rd_map)
        ref_map, ref_dets, ref_counts = fp.process_batch(cube, mti_mode=mti_mode, combine=combine, peak_group=group, **kw)
        err = float(np.abs(rd_map - ref_map).max())
        assert err <= (2 if combine == "beams" else 1) * MAP_TOL * float(ref_map.max()), (combine, err / float(ref_map.max()))
        check_canonical(dets)
        assert int(n_dets.sum()) == len(dets) and status.tolist() == expected_status(cube)
        # bit-exact on the GPU's own map: CFAR, then grouping
        exp = []
        for i in range(n_cpi):
            for ch in range(n_maps):
                o = fp.cfar_stage(rd_map[i, ch], **kw)
                assert np.array_equal(dense[i, ch], o)                          # cfar_out stays the ungrouped CFAR output
                exp.append(fp.detections(fp.peak_group(o) if group else o, i, ch))
        exp = np.concatenate(exp)
        for f in ("cpi", "channel", "range", "doppler", "mag"):
            assert np.array_equal(dets[f], exp[f]) and np.array_equal(d_tap[f], exp[f]), (combine, group, f)
        # against the oracle's own map: identical outside near-threshold cells (grouping can move a kept peak only
        # where the CFAR decision of a neighbour differs, so it is compared on the ungrouped lists' ambiguity)
        if not group:
            eps = max(4.0 * err, 1e-7 * float(ref_map.max()))
            diff = det_cells(dets) ^ det_cells(ref_dets)
            for (i, ch, r, d) in diff:
                assert fp.ambiguous_cells(ref_map[i, ch], eps, **kw)[r, d], (combine, (i, ch, r, d))
        if group:
            assert len(exp) < sum(int(np.count_nonzero(fp.cfar_stage(rd_map[i, ch], **kw))) for i in range(n_cpi) for ch in range(n_maps))
        assert len(dets) > 0
