"""Synthetic FMCW scenarios: the input side of the path (ADC cubes), the counterpart of the stimulus processes
of the reference's testbenches.

* synthetic_cube    -- BASELINE.json config 2 / 3 style multi-target cubes: per CPI a few point
                       targets (range bin, Doppler bin, amplitude), a per-channel phase offset and Gaussian noise
                       per component (as rtl/src/tb_tactical.vhd:59,301-303), quantised like the testbenches do
                       (round to nearest, clamp to int16: rtl/src/tb_radar_core.vhd:109-113).
* TacticalScenario  -- the moving-target scenario of rtl/src/tb_tactical.vhd:28-40,149-171: targets with range /
                       Doppler rates, range-dependent clutter, Gaussian noise, one cube per scan.

Used by bench.py, the examples and the tests for INPUT only; nothing here computes any stage of the path.
"""
from __future__ import annotations

import numpy as np


def quantise_iq(i_acc, q_acc) -> np.ndarray:
    """VHDL integer(real) -- round to nearest, ties away from zero -- then clamp to int16; [..., 2] = (I, Q)."""
    def vhdl_integer(x):
        return (np.sign(x) * np.floor(np.abs(x) + 0.5)).astype(np.int64)
    iq = np.stack([vhdl_integer(i_acc), vhdl_integer(q_acc)], axis=-1)
    return np.clip(iq, -32768, 32767).astype(np.int16)


def synthetic_cube(n_cpi, n_channels, n_chirps, n_samples, seed=1234, dtype="int16",
                   n_targets=(3, 8), amp=(1000.0, 12000.0), sigma=50.0, fixed_targets=None):
    """SURVEY 8(d) config-3 style cubes: per CPI 3-8 targets with uniformly random
    (R, D, A), a per-channel phase offset and Gaussian noise sigma per component
    (as tb_tactical.vhd:59,301-303), default_rng(seed + cpi).
    Returns (cube [cpi, ch, chirp, sample, 2] int16 or [cpi, ch, chirp, sample]
    complex64, list of per-CPI target arrays [(R, D, A)])."""
    s = np.arange(n_samples, dtype=np.float64)[None, :]
    c = np.arange(n_chirps, dtype=np.float64)[:, None]
    out = []
    truth = []
    for i in range(n_cpi):
        rng = np.random.default_rng(seed + i)
        if fixed_targets is not None:
            tg = np.asarray(fixed_targets, dtype=np.float64)
        else:
            nt = int(rng.integers(n_targets[0], n_targets[1] + 1))
            tg = np.stack([rng.integers(max(1, n_samples // 25), n_samples - n_samples // 25, nt).astype(np.float64),
                           rng.integers(0, n_chirps, nt).astype(np.float64),
                           rng.uniform(amp[0], amp[1], nt)], axis=1)
        truth.append(tg)
        base = np.zeros((n_chirps, n_samples), dtype=np.complex128)
        for (r, d, a) in tg:
            base += a * np.exp(2j * np.pi * (r * s / n_samples + d * c / n_chirps))
        ph = np.exp(1j * rng.uniform(0, 2 * np.pi, n_channels))[:, None, None]
        x = base[None] * ph
        x = x + sigma * (rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape))
        if dtype == "int16":
            out.append(quantise_iq(x.real, x.imag))
        else:
            out.append(x.astype(np.complex64))
    return np.stack(out), truth


class TacticalScenario:
    """Scan-by-scan cubes of moving point targets over clutter and noise -- the scenario generator of
    rtl/src/tb_tactical.vhd (target list :28-40, per-scan kinematics :149-171, Gaussian noise by Box-Muller
    :301-303, clutter falling off with range :160-163) with NumPy's generator in place of ieee.math_real.uniform.

    targets: iterable of (range_bin, doppler_bin, amplitude, range_rate, doppler_rate) -- bins and bins per scan.
    Doppler bins are natural-order FFT bins: a target moving to negative Doppler wraps to n_chirps - |d|."""

    def __init__(self, n_chirps=128, n_samples=1024, n_channels=1, targets=((200, 20, 6000.0, 1.5, 0.0),
                 (520, 100, 4000.0, -2.0, 0.25), (760, 64, 2500.0, 0.0, -0.5)), noise_sigma=50.0,
                 clutter_amp=800.0, clutter_extent=0.15, seed=42, dtype="int16"):
        self.n_chirps, self.n_samples, self.n_channels = n_chirps, n_samples, n_channels
        self.targets = [tuple(float(v) for v in t) for t in targets]
        self.noise_sigma, self.clutter_amp, self.clutter_extent = float(noise_sigma), float(clutter_amp), float(clutter_extent)
        self.rng = np.random.default_rng(seed)
        self.dtype = dtype
        self.scan = 0

    def truth(self, scan=None):
        """[(range_bin, doppler_bin, amplitude)] of every target at `scan` (default: the next one)."""
        k = self.scan if scan is None else scan
        return [((r + rr * k) % self.n_samples, (d + dr * k) % self.n_chirps, a) for (r, d, a, rr, dr) in self.targets]

    def next_cube(self):
        """One scan: cube [1, channel, chirp, sample(, 2)] and its truth list; advances the scenario."""
        ns, nc = self.n_samples, self.n_chirps
        s = np.arange(ns, dtype=np.float64)[None, :]
        c = np.arange(nc, dtype=np.float64)[:, None]
        tg = self.truth()
        base = np.zeros((nc, ns), dtype=np.complex128)
        for (r, d, a) in tg:
            base += a * np.exp(2j * np.pi * (r * s / ns + d * c / nc))
        # stationary clutter: zero Doppler, the nearest `clutter_extent` of the range bins, random phase per bin
        n_cl = max(1, int(self.clutter_extent * ns))
        bins = np.arange(1, n_cl + 1)
        amp = self.clutter_amp * (1.0 - bins / (n_cl + 1.0)) / np.sqrt(n_cl)
        phase = self.rng.uniform(0.0, 2.0 * np.pi, n_cl)
        clutter = (amp[:, None] * np.exp(1j * (2.0 * np.pi * bins[:, None] * s / ns + phase[:, None]))).sum(axis=0)
        base += clutter                                              # the same on every chirp: Doppler bin 0
        ph = np.exp(1j * self.rng.uniform(0, 2 * np.pi, self.n_channels))[:, None, None]
        x = base[None] * ph
        x = x + self.noise_sigma * (self.rng.standard_normal(x.shape) + 1j * self.rng.standard_normal(x.shape))
        self.scan += 1
        cube = quantise_iq(x.real, x.imag) if self.dtype == "int16" else x.astype(np.complex64)
        return cube[None], tg
