"""Stimulus generators that reproduce the reference testbenches.  TEST INFRASTRUCTURE ONLY.

* vhdl_uniform      -- ieee.math_real.uniform (IEEE 1076.2; L'Ecuyer 1988 combined
                       LCG), the PRNG behind every reference testbench
                       (e.g. rtl/src/tb_radar_core.vhd:75,88-107).  The algorithm is
                       not in /root/reference; it is restated from the standard.
* tb_radar_core_frames / old_tb_radar_core_cube -- the integration-TB stimuli.
* cfar_tb_map       -- rtl/src/tb_os_cfar_2d.vhd:52-75.
* synthetic_cube    -- re-exported from the product's scenario module (BASELINE config 2 / 3 style cubes).
"""
from __future__ import annotations

import numpy as np


class VhdlUniform:
    """ieee.math_real.uniform(seed1, seed2, x)."""

    def __init__(self, seed1: int, seed2: int):
        self.s1, self.s2 = int(seed1), int(seed2)

    def __call__(self) -> float:
        k = self.s1 // 53668
        self.s1 = 40014 * (self.s1 - k * 53668) - k * 12211
        if self.s1 < 0:
            self.s1 += 2147483563
        k = self.s2 // 52774
        self.s2 = 40692 * (self.s2 - k * 52774) - k * 3791
        if self.s2 < 0:
            self.s2 += 2147483399
        z = self.s1 - self.s2
        if z < 1:
            z += 2147483562
        return float(z) * 4.656613e-10

    def draw(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        s1, s2 = self.s1, self.s2
        for i in range(n):
            k = s1 // 53668
            s1 = 40014 * (s1 - k * 53668) - k * 12211
            if s1 < 0:
                s1 += 2147483563
            k = s2 // 52774
            s2 = 40692 * (s2 - k * 52774) - k * 3791
            if s2 < 0:
                s2 += 2147483399
            z = s1 - s2
            if z < 1:
                z += 2147483562
            out[i] = z * 4.656613e-10
        self.s1, self.s2 = s1, s2
        return out


def _vhdl_integer(x: np.ndarray) -> np.ndarray:
    """VHDL integer(real): round to nearest, ties away from zero."""
    return (np.sign(x) * np.floor(np.abs(x) + 0.5)).astype(np.int64)


def _quantise(i_acc, q_acc) -> np.ndarray:
    iq = np.stack([_vhdl_integer(i_acc), _vhdl_integer(q_acc)], axis=-1)
    return np.clip(iq, -32768, 32767).astype(np.int16)


def tb_radar_core_frames(n_frames=3, n_range=1024, n_doppler=128, seed=(1234, 1234),
                         amp=20000.0, noise_floor=100.0):
    """rtl/src/tb_radar_core.vhd:74-123: per frame two draws (range, Doppler), then
    per sample two draws (I noise, Q noise).  Returns (cube int16
    [frame, chirp, sample, 2], targets [(range, doppler)] as reals)."""
    rng = VhdlUniform(*seed)
    frames, targets = [], []
    s = np.arange(n_range, dtype=np.float64)[None, :]
    c = np.arange(n_doppler, dtype=np.float64)[:, None]
    for _ in range(n_frames):
        tr = rng() * 900.0 + 50.0
        td = rng() * 100.0
        phase = 2.0 * np.pi * (tr * s / n_range + td * c / n_doppler)
        u = rng.draw(2 * n_range * n_doppler).reshape(n_doppler, n_range, 2)
        i_acc = amp * np.cos(phase) + (u[..., 0] - 0.5) * 2.0 * noise_floor
        q_acc = amp * np.sin(phase) + (u[..., 1] - 0.5) * 2.0 * noise_floor
        frames.append(_quantise(i_acc, q_acc))
        targets.append((tr, td))
    return np.stack(frames), targets


def old_tb_radar_core_cube(n_range=1024, n_doppler=128, seed=(1, 1), noise_floor=20.0,
                           targets=((100, 5.0, 8000.0), (500, -10.0, 5000.0)), rng=None):
    """rtl/old/tb_radar_core.vhd:37-44,101-141: two fixed targets + uniform noise.
    One CPI, int16 [chirp, sample, 2].  `rng` (a numpy Generator) replaces the
    VHDL-exact PRNG when given."""
    s = np.arange(n_range, dtype=np.float64)[None, :]
    c = np.arange(n_doppler, dtype=np.float64)[:, None]
    i_acc = np.zeros((n_doppler, n_range))
    q_acc = np.zeros((n_doppler, n_range))
    for (r, d, a) in targets:
        phase = 2.0 * np.pi * (r * s / n_range + d * c / n_doppler)
        i_acc += a * np.cos(phase)
        q_acc += a * np.sin(phase)
    if rng is None:
        u = VhdlUniform(*seed).draw(2 * n_range * n_doppler).reshape(n_doppler, n_range, 2)
    else:
        u = rng.random((n_doppler, n_range, 2))
    i_acc += noise_floor * (u[..., 0] - 0.5) * 2.0
    q_acc += noise_floor * (u[..., 1] - 0.5) * 2.0
    return _quantise(i_acc, q_acc)


def cfar_tb_map(noise_floor=100, tgt_amp=5000, t1=(30, 16), t2=(50, 8), n_range=64, n_doppler=32):
    """rtl/src/tb_os_cfar_2d.vhd:52-75 `make_map`."""
    r = np.arange(n_range)[:, None]
    d = np.arange(n_doppler)[None, :]
    m = noise_floor + ((r * 7 + d * 13) % 30)
    m = m.astype(np.int64)
    for dr in (-1, 0, 1):
        for dd in (-1, 0, 1):
            for (tr, td) in (t1, t2):
                rr, cc = tr + dr, td + dd
                if 0 <= rr < n_range and 0 <= cc < n_doppler:
                    m[rr, cc] = tgt_amp if (dr == 0 and dd == 0) else tgt_amp // 3
    return m.astype(np.uint32)


# BASELINE config 2 / 3 style cubes live with the product's scenario generator (inputs only); the tests reach them
# through this module as before
from fpga_fmcw_radar_processor_b200.scenario import synthetic_cube  # noqa: E402,F401
