"""Float restatement of the radar_core dataflow.  TEST INFRASTRUCTURE ONLY.

Dataflow and parameters follow rtl/src/radar_core.vhd:267-418 (window -> range
FFT -> corner turn -> MTI -> window -> Doppler FFT -> magnitude -> OS-CFAR ->
non-zero filter), with real arithmetic in place of the 16-bit datapath:

  * the FFT stages are the mathematical forward DFT, natural order (config word
    bit0 = 1 = forward, rtl/src/radar_core.vhd:246-248); the fixed-point
    behaviour of the Xilinx IP is not available => "parity unpinned";
  * stages up to the magnitude are evaluated in float64 ("truth" for the
    tolerance check of the float32 CUDA kernels: max |err| <= 1e-4 * map peak);
  * the CFAR decision is evaluated with EXACT float32 semantics on a float32
    map, so that the CUDA kernel can be bit-exact on the same map.
"""
from __future__ import annotations

import numpy as np

from .rtl_fixed import CfarGeometry, cfar_refs

DEFAULTS = dict(
    cfar_ref_r=4, cfar_ref_d=4, cfar_guard_r=2, cfar_guard_d=1,   # radar_core.vhd:12-20
    rank_pct=75, scale_min=2, scale_nom=4, scale_max=6,           # radar_core.vhd:380
    scale_override=0,                                             # cfar_scale_ovr, :49
    mti_mode=0,            # 0 = bypass (tb_radar_core.vhd:66), 2 / 3 = pulse canceller
    mag_mode="amb",        # amb | abs | pow | logpow
    cfar_axes="as_built",  # as_built | as_named   
    cfar_edges="zero",     # zero | raster          
    window_gain=1.0,       # 1.0, or 2.0 to mirror the as-built x2 gain of the RTL window
    # optional stages that are NOT in the reference (the RTL has one channel and no grouping):
    combine="none",        # none | nci: non-coherent integration, magnitudes of all channels added before the CFAR
                           # | beams: coherent beams of a uniform linear array (angle DFT over the channels)
    peak_group=False,      # keep a detection only if no neighbouring detection (8-neighbourhood) is larger
)


def params(**kw) -> dict:
    p = dict(DEFAULTS)
    for k, v in kw.items():
        if k not in p:
            raise TypeError(f"unknown parameter {k!r}")
        p[k] = v
    return p


def hamming(n: int) -> np.ndarray:
    """Symmetric Hamming, same formula as the ROM generator
    (rtl/src/window_multiplier.vhd:41-42), as the float32 table the kernels use."""
    i = np.arange(n, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2.0 * np.pi * i / float(n - 1))).astype(np.float32)


def to_complex(cube: np.ndarray) -> np.ndarray:
    """int16 IQ [..., 2] (I first: I is the low half-word of the stream word,
    rtl/src/window_multiplier.vhd:93-94) or complex -> complex128."""
    cube = np.asarray(cube)
    if np.iscomplexobj(cube):
        return cube.astype(np.complex128)
    assert cube.shape[-1] == 2, "int16 input must be [..., sample, 2] = (I, Q)"
    return cube[..., 0].astype(np.float64) + 1j * cube[..., 1].astype(np.float64)


def range_stage(cube: np.ndarray, window_gain: float = 1.0) -> np.ndarray:
    """[..., chirp, sample] -> [..., range, chirp] complex128.
    window (radar_core.vhd:267-276) -> forward DFT over samples (:303-316) ->
    corner turn (:318-327; out[r][d] = in[d][r], corner_turner.vhd:79-80)."""
    x = to_complex(cube)
    n = x.shape[-1]
    w = hamming(n).astype(np.float64) * float(window_gain)
    X = np.fft.fft(x * w, axis=-1)
    return np.ascontiguousarray(np.swapaxes(X, -1, -2))


def mti_float(x: np.ndarray, mode: int) -> np.ndarray:
    """doppler_notch.vhd:72-93 without saturation; zero initial state per line."""
    if mode == 0:
        return x
    d1 = np.zeros_like(x)
    d1[..., 1:] = x[..., :-1]
    if mode == 2:
        return x - d1
    if mode == 3:
        d2 = np.zeros_like(x)
        d2[..., 2:] = x[..., :-2]
        return x - 2.0 * d1 + d2
    raise ValueError("mti_mode must be 0, 2 or 3")


def magnitude_float(z: np.ndarray, mode: str = "amb") -> np.ndarray:
    """amb = max(|I|,|Q|) + 0.375 min(|I|,|Q|) (magnitude_calc.vhd:55-60,78-81 with
    min>>2 + min>>3 -> 0.375*min); abs / pow / logpow are the alternatives
    north_star names (|X|, |X|^2, 10 log10 |X|^2)."""
    ai, aq = np.abs(z.real), np.abs(z.imag)
    if mode == "amb":
        return np.maximum(ai, aq) + 0.375 * np.minimum(ai, aq)
    if mode == "abs":
        return np.hypot(ai, aq)
    p = ai * ai + aq * aq
    if mode == "pow":
        return p
    if mode == "logpow":
        return 10.0 * np.log10(np.maximum(p, 1e-30))
    raise ValueError(mode)


def doppler_stage(x_rc: np.ndarray, mti_mode: int = 0, mag_mode: str = "amb",
                  window_gain: float = 1.0, return_complex: bool = False) -> np.ndarray:
    """[..., range, chirp] complex -> RD magnitude map [..., range, doppler] float64.
    MTI (radar_core.vhd:329-338) -> window (:340-349) -> forward DFT over chirps
    (:351-364) -> magnitude (:366-374)."""
    x = np.asarray(x_rc).astype(np.complex128)
    n = x.shape[-1]
    w = hamming(n).astype(np.float64) * float(window_gain)
    Y = np.fft.fft(mti_float(x, mti_mode) * w, axis=-1)
    if return_complex:
        return Y
    return magnitude_float(Y, mag_mode)


def rd_map(cube: np.ndarray, **kw) -> np.ndarray:
    """Frame(s) -> float64 RD magnitude map [..., range, doppler]."""
    p = params(**kw)
    return doppler_stage(range_stage(cube, p["window_gain"]), p["mti_mode"], p["mag_mode"],
                         p["window_gain"])


# ----------------------------------------------------------------------------
# CFAR with exact float32 semantics
# ----------------------------------------------------------------------------
def _geometry(p: dict) -> CfarGeometry:
    return CfarGeometry(p["cfar_ref_r"], p["cfar_ref_d"], p["cfar_guard_r"], p["cfar_guard_d"],
                        p["rank_pct"], p["cfar_axes"])


def cfar_stage(mag32: np.ndarray, return_debug: bool = False, **kw):
    """2D OS-CFAR on ONE float32 map [range][doppler]; os_cfar_2d.vhd:152-217 with
    real arithmetic, every operation in float32 except the reference sum:

        S     = sum(refs) accumulated in float64
        mean  = float32(S / N_REF)
        lo    = 0.5f * mean            hi = mean + lo            (float32)
        scale = override or (MAX if x_k > hi else MIN if x_k < lo else NOM)
        thr   = float32(x_k * float32(scale))
        out   = cut if cut > thr else 0
    """
    p = params(**kw)
    mag32 = np.ascontiguousarray(mag32, dtype=np.float32)
    assert mag32.ndim == 2
    g = _geometry(p)
    refs = cfar_refs(mag32, g, p["cfar_edges"])
    S = refs.sum(axis=-1, dtype=np.float64)
    xk = np.partition(refs, g.rank_idx, axis=-1)[..., g.rank_idx].astype(np.float32)
    mean = (S / float(g.n_ref)).astype(np.float32)
    lo = (np.float32(0.5) * mean).astype(np.float32)
    hi = (mean + lo).astype(np.float32)
    if p["scale_override"]:
        scale = np.full(mag32.shape, np.float32(p["scale_override"]), dtype=np.float32)
    else:
        scale = np.where(xk > hi, np.float32(p["scale_max"]),
                         np.where(xk < lo, np.float32(p["scale_min"]),
                                  np.float32(p["scale_nom"]))).astype(np.float32)
    thr = (xk * scale).astype(np.float32)
    out = np.where(mag32 > thr, mag32, np.float32(0)).astype(np.float32)
    if return_debug:
        return out, dict(xk=xk, mean=mean, lo=lo, hi=hi, scale=scale, thr=thr, geometry=g)
    return out


def ambiguous_cells(mag32: np.ndarray, eps: float, **kw) -> np.ndarray:
    """Boolean mask of cells whose CFAR decision may legitimately differ when every
    map value is perturbed by at most `eps` (absolute): the cells "whose statistic
    lies within the stated tolerance of the threshold" (north_star).  Used to
    compare detection lists computed from two maps that agree to `eps`."""
    p = params(**kw)
    _, dbg = cfar_stage(mag32, return_debug=True, **kw)
    xk, lo, hi = dbg["xk"].astype(np.float64), dbg["lo"].astype(np.float64), dbg["hi"].astype(np.float64)
    cut = np.asarray(mag32, dtype=np.float64)
    if p["scale_override"]:
        cands = [float(p["scale_override"])]
        amb = np.zeros(cut.shape, dtype=bool)
        for s in cands:
            amb |= np.abs(cut - s * xk) <= eps * (1.0 + s)
        return amb
    smin, snom, smax = float(p["scale_min"]), float(p["scale_nom"]), float(p["scale_max"])
    amb = np.zeros(cut.shape, dtype=bool)
    # candidate scales: the selected one plus neighbours when x_k is within the
    # perturbation band of a selection boundary (x_k, mean move by <= eps)
    near_hi = np.abs(xk - hi) <= 2.5 * eps
    near_lo = np.abs(xk - lo) <= 1.5 * eps
    sel = dbg["scale"].astype(np.float64)
    for s, extra in ((smin, near_lo), (snom, near_lo | near_hi), (smax, near_hi)):
        use = (sel == s) | extra
        amb |= use & (np.abs(cut - s * xk) <= eps * (1.0 + s))
    # a changed scale flips the decision outright when cut lies between the two thresholds
    both = near_lo & (cut > smin * xk - eps * (1 + smin)) & (cut <= snom * xk + eps * (1 + snom))
    both |= near_hi & (cut > snom * xk - eps * (1 + snom)) & (cut <= smax * xk + eps * (1 + smax))
    return amb | both


DET_DTYPE = np.dtype([("cpi", "<i4"), ("channel", "<i4"), ("range", "<i4"),
                      ("doppler", "<i4"), ("mag", "<f4")])


def detections(cfar_out: np.ndarray, cpi: int = 0, channel: int = 0) -> np.ndarray:
    """Non-zero filter + TRUE coordinates (radar_core.vhd:396-418),
    canonical (range, doppler) order."""
    r, d = np.nonzero(cfar_out)
    out = np.zeros(r.size, dtype=DET_DTYPE)
    out["cpi"], out["channel"], out["range"], out["doppler"] = cpi, channel, r, d
    out["mag"] = cfar_out[r, d]
    return out


def peak_group(cfar_out: np.ndarray) -> np.ndarray:
    """Peak grouping of one map's CFAR output (CUT or 0 per cell): a detection is dropped when one of its 8
    neighbours inside the map is a detection of larger magnitude, or of equal magnitude and earlier in (range,
    doppler) raster order.  Not in the reference; this is the rule librdp's K3b kernel implements."""
    out = np.asarray(cfar_out)
    det = out != 0
    keep = det.copy()
    nr, nd = out.shape
    big = np.pad(out.astype(np.float64), 1, constant_values=0.0)
    dpad = np.pad(det, 1, constant_values=False)
    for dl in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dl == 0 and dc == 0:
                continue
            nb = big[1 + dl:1 + dl + nr, 1 + dc:1 + dc + nd]
            nb_det = dpad[1 + dl:1 + dl + nr, 1 + dc:1 + dc + nd]
            earlier = dl < 0 or (dl == 0 and dc < 0)
            beaten = nb_det & ((nb > out) | ((nb == out) & earlier))
            keep &= ~beaten
    return np.where(keep, out, np.zeros_like(out))


def process_cpi(frame: np.ndarray, **kw):
    """One CPI: frame [chirp, sample(,2)] or [channel, chirp, sample(,2)] ->
    (rd_map float32 [channel, range, doppler], detections DET_DTYPE sorted by
    (channel, range, doppler)).  This is the surface the CUDA path mirrors."""
    frame = np.asarray(frame)
    cplx = np.iscomplexobj(frame)
    if frame.ndim == (2 if cplx else 3):
        frame = frame[None]
    p = params(**kw)
    if p["combine"] == "beams":
        # n_beams = smallest power of two >= channels; beam b = sum_c x_c exp(-j 2 pi b c / n_beams) on the range-major
        # intermediate (linear: the same as an angle DFT across the channels after the Doppler transform)
        x = range_stage(frame, p["window_gain"])                 # [channel, range, chirp]
        n_ch = x.shape[0]
        n_beams = 1
        while n_beams < n_ch:
            n_beams <<= 1
        steer = np.exp(-2j * np.pi * np.outer(np.arange(n_beams), np.arange(n_ch)) / n_beams)
        y = np.einsum("bc,crk->brk", steer, x)
        m = doppler_stage(y, p["mti_mode"], p["mag_mode"], p["window_gain"])
    else:
        m = rd_map(frame, **kw)
    if p["combine"] == "nci":                                   # one combined map per CPI, channel index 0
        m = m.sum(axis=0, keepdims=True)
    m = m.astype(np.float32)
    outs = [cfar_stage(m[c], **kw) for c in range(m.shape[0])]
    if p["peak_group"]:
        outs = [peak_group(o) for o in outs]
    dets = [detections(o, 0, c) for c, o in enumerate(outs)]
    return m, np.concatenate(dets) if dets else np.zeros(0, DET_DTYPE)


def process_batch(cubes: np.ndarray, **kw):
    """[cpi, channel, chirp, sample(,2)] -> (maps [cpi, channel, range, doppler],
    detections for all CPIs in canonical order, per-CPI counts)."""
    maps, dets, counts = [], [], []
    for i in range(cubes.shape[0]):
        m, d = process_cpi(cubes[i], **kw)
        d["cpi"] = i
        maps.append(m)
        dets.append(d)
        counts.append(d.size)
    return np.stack(maps), np.concatenate(dets), np.asarray(counts, dtype=np.int32)
