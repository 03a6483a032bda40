"""Bit-exact integer restatement of the hand-written RTL stages.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to
/root/reference).  "as-built" = what the VHDL text does under VHDL signal /
variable semantics, not what its comments claim.

The FFT IP (Xilinx xfft v9.1) is NOT restated here: its RTL is absent from the
reference tree, so the fixed-point FFT is PARITY UNPINNED (oracle/__init__.py).
`bfp_fft` below is a SELF-DEFINED block-floating-point transform with the IP's
configured word widths and rounding rule; `process_cpi_fixed` chains it with the
pinned stages -- the end-to-end fixed-point mode is self-consistent (librdp is
tested bit for bit against it), not Xilinx-exact.
"""
from __future__ import annotations

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

INT16_MAX = 32767
INT16_MIN = -32768


# ----------------------------------------------------------------------------
# window_multiplier  (rtl/src/window_multiplier.vhd)
# ----------------------------------------------------------------------------
def hamming_rom(n_samples: int, coef_width: int = 16) -> np.ndarray:
    """Half-length Q15 Hamming ROM, rtl/src/window_multiplier.vhd:30,34-49.

    rom[i] = clamp(integer((0.54 - 0.46 cos(2 pi i/(N-1))) * (2^(W-1)-1)), 0, 2^(W-1)-1)
    VHDL integer(real) rounds to nearest (ties away from zero; no tie occurs for
    the sizes tested -- asserted in tests/test_oracle_fixed.py).
    """
    rom_size = n_samples // 2
    i = np.arange(rom_size, dtype=np.float64)
    full = float(2 ** (coef_width - 1) - 1)
    coef_real = 0.54 - 0.46 * np.cos(2.0 * np.pi * i / float(n_samples - 1))
    coef_int = np.floor(coef_real * full + 0.5).astype(np.int64)
    return np.clip(coef_int, 0, int(full)).astype(np.int32)


def window_coef_index(n_samples: int) -> np.ndarray:
    """ROM address per sample index, rtl/src/window_multiplier.vhd:97-102."""
    rom_size = n_samples // 2
    n = np.arange(n_samples)
    a = np.where(n < rom_size, n, n_samples - 1 - n)
    return np.minimum(a, rom_size - 1)


def window(x: np.ndarray, n_samples: int | None = None, coef_width: int = 16):
    """Fixed-point window along the LAST axis of an integer array (I or Q plane).

    rtl/src/window_multiplier.vhd:126-127 (16x16 multiply), :146-149 (add
    2^(W-2), take bits [30:14] as 17-bit signed == (p + 16384) >> 14), :151-158
    (saturate to int16, flag).  As-built gain is 2*w[n] and x=0 -> y=1.

    Returns (y int16-valued int32 array, saturated bool array).
    """
    x = np.asarray(x).astype(np.int64)
    n = x.shape[-1] if n_samples is None else n_samples
    assert x.shape[-1] == n
    coef = hamming_rom(n, coef_width).astype(np.int64)[window_coef_index(n)]
    p = x * coef
    v = (p + (1 << (coef_width - 2))) >> (coef_width - 2)
    sat = (v > INT16_MAX) | (v < INT16_MIN)
    y = np.clip(v, INT16_MIN, INT16_MAX)
    return y.astype(np.int32), sat


# ----------------------------------------------------------------------------
# corner_turner  (rtl/src/corner_turner.vhd)
# ----------------------------------------------------------------------------
def corner_turn(x: np.ndarray) -> np.ndarray:
    """out[r][d] = in[chirp d][sample r]; rtl/src/corner_turner.vhd:79-80
    (wr_addr = chirp*N_RANGE + sample, rd_addr = range + doppler*N_RANGE, range
    slow / Doppler fast on the read side, :147-158).  Works on [..., chirp, sample]."""
    return np.ascontiguousarray(np.swapaxes(np.asarray(x), -1, -2))


# ----------------------------------------------------------------------------
# doppler_notch  (rtl/src/doppler_notch.vhd)
# ----------------------------------------------------------------------------
def mti(x: np.ndarray, mode: int = 2, bypass: bool = False) -> np.ndarray:
    """2- or 3-pulse canceller along the LAST axis (one range line), with int16
    saturation; delay line is zero at the start of each line
    (rtl/src/doppler_notch.vhd:70-71 bypass, :72-80 mode 2, :81-93 mode 3,
    :96-102 delay update / clear on tlast)."""
    x = np.asarray(x).astype(np.int64)
    if bypass or mode == 0:
        return x.astype(np.int32)
    d1 = np.zeros_like(x)
    d1[..., 1:] = x[..., :-1]
    if mode == 2:
        y = x - d1
    elif mode == 3:
        d2 = np.zeros_like(x)
        d2[..., 2:] = x[..., :-2]
        y = x - 2 * d1 + d2
    else:
        raise ValueError("mode must be 0 (bypass), 2 or 3")
    return np.clip(y, INT16_MIN, INT16_MAX).astype(np.int32)


# ----------------------------------------------------------------------------
# magnitude_calc  (rtl/src/magnitude_calc.vhd)
# ----------------------------------------------------------------------------
def magnitude(i: np.ndarray, q: np.ndarray) -> np.ndarray:
    """alpha-max-beta-min: max(|I|,|Q|) + (min>>2) + (min>>3), 17-bit unsigned.
    rtl/src/magnitude_calc.vhd:55-60 (abs as 16-bit unsigned, so -32768 -> 32768;
    max/min), :78-81 (sum).  Same as tb_magnitude_calc.vhd:32-40 `amb_ref`."""
    ai = np.abs(np.asarray(i).astype(np.int64))
    aq = np.abs(np.asarray(q).astype(np.int64))
    mx = np.maximum(ai, aq)
    mn = np.minimum(ai, aq)
    return (mx + (mn >> 2) + (mn >> 3)).astype(np.uint32)


# ----------------------------------------------------------------------------
# os_cfar_2d  (rtl/src/os_cfar_2d.vhd)
# ----------------------------------------------------------------------------
class CfarGeometry:
    """Window geometry in PHYSICAL axes (lines = range bins, cols = Doppler bins).

    as_built (default): the RTL generics named *_RANGE act along the Doppler
    (fast / stream) axis and *_DOPPLER along the range (line) axis, because the
    stream is range-major: rtl/src/os_cfar_2d.vhd:41-47 (constants), :140-146
    (window built from line buffer rows d and shift-register taps r), :157-159
    (guard test).  as_named swaps them to what the names say.
    """

    def __init__(self, ref_range=4, ref_doppler=4, guard_range=2, guard_doppler=1,
                 rank_pct=75, axes="as_built"):
        if axes == "as_built":
            self.half_lines = ref_doppler + guard_doppler   # CUT_D, along range
            self.guard_lines = guard_doppler
            self.half_cols = ref_range + guard_range        # CUT_R, along Doppler
            self.guard_cols = guard_range
        elif axes == "as_named":
            self.half_lines = ref_range + guard_range
            self.guard_lines = guard_range
            self.half_cols = ref_doppler + guard_doppler
            self.guard_cols = guard_doppler
        else:
            raise ValueError(axes)
        self.win_lines = 2 * self.half_lines + 1
        self.win_cols = 2 * self.half_cols + 1
        guard_area = (2 * self.guard_lines + 1) * (2 * self.guard_cols + 1)
        self.n_ref = self.win_lines * self.win_cols - guard_area        # :45-47
        self.rank_idx = min((self.n_ref * rank_pct) // 100, self.n_ref - 1)  # :181-182
        m = np.ones((self.win_lines, self.win_cols), dtype=bool)
        m[self.half_lines - self.guard_lines:self.half_lines + self.guard_lines + 1,
          self.half_cols - self.guard_cols:self.half_cols + self.guard_cols + 1] = False
        self.ref_mask = m


def _padded(mag: np.ndarray, g: CfarGeometry, edges: str) -> np.ndarray:
    n_lines, n_cols = mag.shape
    hl, hc = g.half_lines, g.half_cols
    if edges == "zero":
        return np.pad(mag, ((hl, hl), (hc, hc)))
    if edges == "raster":
        # Window taken on the flat raster F[line*N_D + col]; cells before the
        # first / after the last sample of the CPI read as 0 (reset value,
        # rtl/src/os_cfar_2d.vhd:101-102).  SURVEY.md Appendix A.6.2.
        flat = mag.reshape(-1)
        lines = np.arange(-hl, n_lines + hl)[:, None]
        cols = np.arange(-hc, n_cols + hc)[None, :]
        idx = lines * n_cols + cols
        ok = (idx >= 0) & (idx < flat.size)
        out = np.zeros(idx.shape, dtype=mag.dtype)
        out[ok] = flat[idx[ok]]
        return out
    raise ValueError(edges)


def cfar_refs(mag: np.ndarray, g: CfarGeometry, edges: str = "zero") -> np.ndarray:
    """[n_lines, n_cols, n_ref] reference cells of every CUT (guard removed)."""
    p = _padded(np.asarray(mag), g, edges)
    win = sliding_window_view(p, (g.win_lines, g.win_cols))
    return win[..., g.ref_mask]


def os_cfar_2d(mag: np.ndarray, ref_range=4, ref_doppler=4, guard_range=2, guard_doppler=1,
               rank_pct=75, scale_min=2, scale_nom=4, scale_max=6, scale_override=0,
               edges="zero", axes="as_built", return_debug=False):
    """Frame-local integer 2D OS-CFAR decision for every cell of `mag`
    ([range][doppler], unsigned).  Returns the RTL's output word: CUT or 0.

    Per cell (rtl/src/os_cfar_2d.vhd): refs = window minus guard (:152-167),
    sum (:163), ascending sort and x_k = refs[rank_idx] (:172-183),
    mean = floor(sum / N_REF) truncated to DATA_WIDTH bits (:189),
    scale = override if /= 0 (:191-192) else SCALE_MAX if x_k > mean + (mean>>1)
    (:193-194) else SCALE_MIN if x_k < (mean>>1) (:195-196) else SCALE_NOM,
    threshold = x_k * scale (intended exact product, :204), output CUT iff
    CUT > threshold (:213-217).  True coordinates (no +3 beat skew).
    """
    mag = np.asarray(mag).astype(np.int64)
    g = CfarGeometry(ref_range, ref_doppler, guard_range, guard_doppler, rank_pct, axes)
    refs = cfar_refs(mag, g, edges)
    s = refs.sum(axis=-1)
    xk = np.partition(refs, g.rank_idx, axis=-1)[..., g.rank_idx]
    mean = (s // g.n_ref) & 0x1FFFF
    hi = (mean + (mean >> 1)) & 0x1FFFF          # 17-bit unsigned add (:193)
    lo = mean >> 1
    if scale_override:
        scale = np.full(mag.shape, int(scale_override), dtype=np.int64)
    else:
        scale = np.where(xk > hi, scale_max, np.where(xk < lo, scale_min, scale_nom))
    thr = xk * scale
    out = np.where(mag > thr, mag, 0).astype(np.uint32)
    if return_debug:
        return out, dict(xk=xk, mean=mean, scale=scale, thr=thr, geometry=g)
    return out


def os_cfar_2d_stream(samples, n_doppler, ref_range=4, ref_doppler=4, guard_range=2,
                      guard_doppler=1, rank_pct=75, scale_min=2, scale_nom=4, scale_max=6,
                      scale_override=0, tlast_every=None, flush=0):
    """LITERAL clock-by-clock model of rtl/src/os_cfar_2d.vhd:83-230 with
    tvalid = tready = 1 on every clock (pure Python, small inputs only).

    `samples` is the flat range-major stream; tlast is asserted every
    `tlast_every` (default n_doppler) samples.  `flush` zero samples are
    appended so the pipeline drains.  Returns the list of m_axis_tdata words on
    the clocks where m_axis_tvalid = 1, in order (output beat k).  Used to pin
    the closed-form window membership of `os_cfar_2d(edges="raster")`: beat k is
    the decision for flat cell k + 3.
    """
    WIN_RANGE = 2 * ref_range + 2 * guard_range + 1           # :41
    WIN_DOPPLER = 2 * ref_doppler + 2 * guard_doppler + 1     # :42
    CUT_R = ref_range + guard_range                           # :43
    CUT_D = ref_doppler + guard_doppler                       # :44
    N_REF = WIN_RANGE * WIN_DOPPLER - (2 * guard_range + 1) * (2 * guard_doppler + 1)
    STARTUP = (CUT_D + 1) * n_doppler + CUT_R + 2             # :67-68
    rank_idx = min((N_REF * rank_pct) // 100, N_REF - 1)      # :181-182
    tl = n_doppler if tlast_every is None else tlast_every

    line_buffer = [[0] * n_doppler for _ in range(WIN_DOPPLER)]
    window = [[0] * WIN_RANGE for _ in range(WIN_DOPPLER)]
    doppler_cnt = row_cnt = startup_cnt = 0
    window_valid = 0
    out_valid_p1 = out_valid_p2 = 0
    out_data_p2 = 0
    m_valid = 0
    m_data = 0
    beats = []

    stream = [int(v) for v in samples] + [0] * flush
    for m, din in enumerate(stream):
        tlast = 1 if (m % tl) == tl - 1 else 0
        # the testbench checker samples m_axis on the same edge, i.e. OLD values
        if m_valid:
            beats.append(m_data)
        # ---- everything below reads OLD signal values, then commits (signal semantics)
        # 3..6: decision from the OLD window (:152-217)
        refs = []
        for d in range(WIN_DOPPLER):
            for r in range(WIN_RANGE):
                if not (abs(d - CUT_D) <= guard_doppler and abs(r - CUT_R) <= guard_range):
                    refs.append(window[d][r])
        s = sum(refs)
        refs.sort()
        ranked = refs[rank_idx]
        mean = (s // N_REF) & 0x1FFFF
        if scale_override:
            scale = scale_override
        elif ranked > ((mean + (mean >> 1)) & 0x1FFFF):
            scale = scale_max
        elif ranked < (mean >> 1):
            scale = scale_min
        else:
            scale = scale_nom
        cut = window[CUT_D][CUT_R]
        new_out_data_p2 = cut if cut > ranked * scale else 0
        new_out_valid_p1 = window_valid            # and tvalid (=1)  (:207)
        new_out_valid_p2 = out_valid_p1            # (:210)
        new_m_valid, new_m_data = out_valid_p2, out_data_p2   # (:225-227)
        # 2: window shift, new column from the OLD line buffer (:140-146)
        new_window = [[0] * WIN_RANGE for _ in range(WIN_DOPPLER)]
        for d in range(WIN_DOPPLER):
            for r in range(WIN_RANGE - 1, 0, -1):
                new_window[d][r] = window[d][r - 1]
            rd_row = (row_cnt + d) % WIN_DOPPLER
            new_window[d][0] = line_buffer[rd_row][doppler_cnt]
        # 1: line buffer write + counters (:118-135)
        line_buffer[row_cnt][doppler_cnt] = din
        if tlast or doppler_cnt == n_doppler - 1:
            doppler_cnt = 0
            row_cnt = 0 if row_cnt == WIN_DOPPLER - 1 else row_cnt + 1
        else:
            doppler_cnt += 1
        if startup_cnt < STARTUP:
            startup_cnt += 1
            window_valid = 0
        else:
            window_valid = 1
        window = new_window
        out_valid_p1, out_valid_p2, out_data_p2 = new_out_valid_p1, new_out_valid_p2, new_out_data_p2
        m_valid, m_data = new_m_valid, new_m_data
    return beats


def detections_from_cfar(cfar_out: np.ndarray) -> np.ndarray:
    """Non-zero filter + coordinates (rtl/src/radar_core.vhd:396-418), with TRUE
    coordinates.  Returns int64 [n, 3] rows (range_bin, doppler_bin, magnitude),
    range-major order -- the `r d mag` record of rtl/src/tb_radar_core.vhd:153-159."""
    r, d = np.nonzero(cfar_out)
    return np.stack([r, d, np.asarray(cfar_out)[r, d].astype(np.int64)], axis=1).astype(np.int64)


# ----------------------------------------------------------------------------
# Block-floating-point FFT for the end-to-end fixed-point mode
# ----------------------------------------------------------------------------
# The reference's two FFTs are Xilinx xfft v9.1 cores: 16-bit data, 16-bit twiddles, block floating point,
# convergent rounding (vivado_proj/.../ip/xfft_0_1/xfft_0.xci:12-27,59; xfft_doppler_1/xfft_doppler.xci:12-27),
# whose RTL is NOT in the reference tree and whose block exponent the top level discards
# (rtl/src/radar_core.vhd:310,358).  Their bit-level arithmetic is therefore unknowable from the tree; the
# fixed-point mode uses the SELF-DEFINED transform below with the same word widths and rounding rule.
# It is "self-consistent, not Xilinx-exact" -- PARITY UNPINNED against the FPGA for this stage.
#
#   data 16-bit two's complement, twiddles Q15: W_N^k = (round(cos(2 pi k / N) * 32767), round(-sin(..) * 32767))
#   log2(N) radix-2 decimation-in-frequency stages; in stage s (half = N >> (s+1)), for every pair (a, b) that is
#   `half` apart inside a block of 2 * half points:
#       a' = a + b                                    (17 bits)
#       b' = cround((a - b) * W_N^(pos << s), 15)     complex product in exact integers, convergent >> 15
#   then ONE shift for the whole frame: sh = the smallest number of bits with ceil(g / 2^sh) <= 32767, g = the
#   largest |component| of the frame after the butterflies; every component becomes cround(v, sh); exponent += sh.
#   After the last stage the frame is in bit-reversed order and is written out in natural order.
#   The transform returns X (int16) and the block exponent e per frame: DFT(x) ~= X * 2^e.
def bfp_twiddles(n: int):
    k = np.arange(n // 2, dtype=np.float64)
    ang = 2.0 * np.pi * k / float(n)
    wr = np.floor(np.cos(ang) * 32767.0 + 0.5).astype(np.int64)
    wi = np.floor(-np.sin(ang) * 32767.0 + 0.5).astype(np.int64)
    return wr, wi


def cround(v: np.ndarray, sh) -> np.ndarray:
    """Convergent (round-half-to-even) arithmetic right shift of integers by `sh` bits (scalar or broadcastable)."""
    v = np.asarray(v, dtype=np.int64)
    sh = np.asarray(sh, dtype=np.int64)
    q = v >> sh
    rem = v - (q << sh)
    half = np.where(sh > 0, np.int64(1) << np.maximum(sh - 1, 0), np.int64(0))
    up = (sh > 0) & ((rem > half) | ((rem == half) & ((q & 1) == 1)))
    return q + up.astype(np.int64)


def _bit_reverse_indices(n: int) -> np.ndarray:
    m = n.bit_length() - 1
    idx = np.arange(n)
    rev = np.zeros(n, dtype=np.int64)
    for b in range(m):
        rev |= ((idx >> b) & 1) << (m - 1 - b)
    return rev


def bfp_fft(re: np.ndarray, im: np.ndarray):
    """Block-floating-point FFT along the LAST axis (length a power of two >= 2) of int16-valued arrays.
    Returns (re int32 holding int16 values, im likewise, exponent int32 per frame)."""
    a_re = np.asarray(re).astype(np.int64)
    a_im = np.asarray(im).astype(np.int64)
    n = a_re.shape[-1]
    m = n.bit_length() - 1
    assert 1 << m == n and n >= 2
    lead = a_re.shape[:-1]
    a_re = a_re.reshape(-1, n)
    a_im = a_im.reshape(-1, n)
    exp = np.zeros(a_re.shape[0], dtype=np.int64)
    wr, wi = bfp_twiddles(n)
    for s in range(m):
        half = n >> (s + 1)
        vr = a_re.reshape(-1, 1 << s, 2, half)
        vi = a_im.reshape(-1, 1 << s, 2, half)
        tr, br, ti, bi = vr[:, :, 0], vr[:, :, 1], vi[:, :, 0], vi[:, :, 1]
        sr, si = tr + br, ti + bi
        dr, di = tr - br, ti - bi
        idx = np.arange(half) << s
        cr, ci = wr[idx], wi[idx]
        pr = cround(dr * cr - di * ci, 15)
        pi_ = cround(dr * ci + di * cr, 15)
        g = np.maximum(np.maximum(np.abs(sr), np.abs(si)), np.maximum(np.abs(pr), np.abs(pi_))).reshape(a_re.shape[0], -1).max(axis=1)
        sh = np.zeros_like(g)
        for k in range(1, 6):
            sh = np.where(((g + (1 << sh) - 1) >> sh) > INT16_MAX, k, sh)
        sh4 = sh[:, None, None]
        out_r = np.stack([cround(sr, sh4), cround(pr, sh4)], axis=2)
        out_i = np.stack([cround(si, sh4), cround(pi_, sh4)], axis=2)
        a_re = out_r.reshape(-1, n)
        a_im = out_i.reshape(-1, n)
        exp += sh
    rev = _bit_reverse_indices(n)
    x_re = np.empty_like(a_re)
    x_im = np.empty_like(a_im)
    x_re[:, rev] = a_re
    x_im[:, rev] = a_im
    return (x_re.reshape(lead + (n,)).astype(np.int32), x_im.reshape(lead + (n,)).astype(np.int32),
            exp.reshape(lead).astype(np.int32))


def process_cpi_fixed(frame: np.ndarray, cfar_ref_r=4, cfar_ref_d=4, cfar_guard_r=2, cfar_guard_d=1, rank_pct=75,
                      scale_min=2, scale_nom=4, scale_max=6, scale_override=0, mti_mode=0, cfar_axes="as_built",
                      cfar_edges="zero"):
    """One CPI through the 16-bit chain of radar_core (rtl/src/radar_core.vhd:267-418) word for word:
    window_multiplier -> range FFT -> corner_turner -> doppler_notch -> window_multiplier -> Doppler FFT ->
    magnitude_calc -> os_cfar_2d -> non-zero filter, every hand-written stage bit-exact as built (x2 window gain,
    +1 LSB bias, saturation), the two FFT IP cores replaced by `bfp_fft` above (self-defined: see its header).
    frame: int16 [channel, chirp, sample, 2] or [chirp, sample, 2].
    Returns (mag uint32 [channel, range, doppler], detections int64 [n, 4] rows (channel, range, doppler, mag) in
    canonical order, saturated bool per channel... ORed into one flag, range exponents [channel, chirp],
    Doppler exponents [channel, range])."""
    frame = np.asarray(frame)
    if frame.ndim == 3:
        frame = frame[None]
    assert frame.dtype == np.int16 and frame.shape[-1] == 2
    xi, xq = frame[..., 0].astype(np.int64), frame[..., 1].astype(np.int64)
    wi, si = window(xi)                                              # range window (radar_core.vhd:267-276)
    wq, sq = window(xq)
    saturated = bool(si.any() or sq.any())
    fr, fi, exp_r = bfp_fft(wi, wq)                                  # range FFT (:303-316)
    cr, ci = corner_turn(fr), corner_turn(fi)                        # corner turn (:318-327)
    mr, mi = mti(cr, mti_mode), mti(ci, mti_mode)                    # MTI (:329-338)
    dwr, s2 = window(mr)                                             # Doppler window (:340-349)
    dwi, s3 = window(mi)
    saturated = saturated or bool(s2.any() or s3.any())
    dr, di, exp_d = bfp_fft(dwr, dwi)                                # Doppler FFT (:351-364)
    mag = magnitude(dr, di)                                          # magnitude (:366-374)
    rows = []
    for ch in range(mag.shape[0]):
        out = os_cfar_2d(mag[ch], cfar_ref_r, cfar_ref_d, cfar_guard_r, cfar_guard_d, rank_pct, scale_min, scale_nom,
                         scale_max, scale_override, edges=cfar_edges, axes=cfar_axes)
        d = detections_from_cfar(out)
        rows.append(np.concatenate([np.full((len(d), 1), ch, dtype=np.int64), d], axis=1))
    dets = np.concatenate(rows) if rows else np.zeros((0, 4), np.int64)
    return mag, dets, saturated, exp_r, exp_d
