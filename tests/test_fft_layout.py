"""The shared-memory layout of the FFT exchanges (csrc/fft_core.cuh, FftPlan::XPAD) restated in Python: for every
plan the kernels use, every 64-bit store and load of every exchange must be bank-conflict free per half-warp
(32 banks of 4 bytes; a 64-bit access by 32 lanes is served as two half-warps).  This is the invariant behind
`profiles/r1_summary.md` "conflict-free second FFT exchange"; a change of the padding that breaks it shows up here
without a GPU."""
import numpy as np
import pytest

# (N, E, radices) as instantiated in csrc/rdp_api.cu
PLANS = [(16, 8, (8, 2)), (32, 8, (8, 4)), (64, 8, (8, 8)), (128, 8, (8, 4, 4)), (256, 8, (8, 8, 4)),
         (512, 8, (8, 8, 8)), (1024, 16, (16, 16, 4)), (2048, 16, (16, 16, 8)), (4096, 16, (16, 16, 16))]


def xpad(i, ns, r):
    """FftPlan::XPAD<PA>: layout of the exchange written by a pass with NS = ns and radix r."""
    if 2 <= ns < 16 and (ns * r) % 16 == 0:
        return i + (i // (ns * r)) * ns
    return i + (i >> 4)


def out_index(n, e, r, ns, tid, b, k):
    j = tid + b * (n // e)
    return (j // ns) * (ns * r) + (j % ns) + k * ns


def in_index(n, e, r, tid, b, k):
    return tid + b * (n // e) + k * (n // r)


def worst_conflict(float2_index_per_lane):
    """Largest number of lanes of one half-warp that hit the same bank with different addresses."""
    worst = 1
    for half in (float2_index_per_lane[:16], float2_index_per_lane[16:]):
        if len(half) == 0:
            continue
        words = {}
        for idx in set(int(v) for v in half):               # identical addresses broadcast
            for w in (2 * idx, 2 * idx + 1):
                words.setdefault(w % 32, set()).add(w)
        worst = max(worst, max(len(v) for v in words.values()))
    return worst


@pytest.mark.parametrize("n,e,radices", PLANS)
def test_exchanges_are_bank_conflict_free(n, e, radices):
    t = n // e
    lanes = np.arange(min(32, t))                            # consecutive threads of one transform share a warp
    ns = 1
    for p, r in enumerate(radices[:-1]):                     # the last pass writes registers / the output tile
        r_next = radices[p + 1]
        extent = 0
        for b in range(e // r):
            for k in range(r):
                st = [xpad(out_index(n, e, r, ns, int(l), b, k), ns, r) for l in lanes]
                assert worst_conflict(st) == 1, f"N={n} exchange {p}: store b={b} r={k} conflicts"
                extent = max(extent, max(st))
        for b in range(e // r_next):
            for k in range(r_next):
                ld = [xpad(in_index(n, e, r_next, int(l), b, k), ns, r) for l in lanes]
                assert worst_conflict(ld) == 1, f"N={n} exchange {p}: load b={b} r={k} conflicts"
        # every element gets its own slot, inside the buffer the plan reserves (XPAD(N - 1) + 2 float2)
        slots = [xpad(i, ns, r) for i in range(n)]
        assert len(set(slots)) == n and max(slots) + 2 >= extent + 1
        ns *= r
