// k3_cfar.cuh -- K3: 2D ordered-statistic CFAR + non-zero filter + detection records.
//
// Replaces os_cfar_2d (rtl/src/os_cfar_2d.vhd:140-217) and the detection filter / coordinate
// counters of rtl/src/radar_core.vhd:396-418.  The RTL bubble-sorts the N_REF reference cells
// of every cell under test (:172-178).  Here nothing is sorted: with k = rank_idx,
//     cut > s * x_(k)   <=>   #{ref : fl(ref * s) < cut} >= k + 1
//     x_(k) > a         <=>   #{ref <= a} <= k
//     x_(k) < a         <=>   #{ref <  a} >= k + 1
// (identities checked in tests/test_oracle_fixed.py).  Phase 1 tests every
// cell against the most permissive scale s1 (scale_override, or the smallest of
// SCALE_MIN/NOM/MAX) from a shared-memory tile of pre-scaled magnitudes; each thread owns a
// strip of S cells along range so a loaded reference value feeds up to S counters.  The few
// survivors (~0.4 % of noise cells) go to phase 2: a warp per survivor evaluates the sum
// (float64), the adaptive scale (:189-199) and the final test exactly.
//   algorithmic HBM bytes / cell: 4 read (+4 if the dense RTL-style output is requested).
#pragma once
#include "rdp_common.cuh"
#include "../../include/rdp.h"

namespace rdp {

struct CfarArgs {
    int n_range, n_doppler, n_channels;
    int half_l, half_c, guard_l, guard_c;   // physical axes: lines = range, cols = Doppler
    int n_ref, rank_idx;
    float s1, smin, snom, smax;
    int adaptive;                           // 1: scale chosen per cell, 0: s1 is the fixed scale
    int raster;                             // cfar_edges
    int max_dets;
};

// magnitude of flat cell (line, col) of one map under the edge rule
__device__ __forceinline__ float cfar_fetch(const float* __restrict__ m, int line, int col,
                                            int n_range, int n_doppler, int raster) {
    if (raster) {
        long long flat = (long long)line * n_doppler + col;
        return (flat >= 0 && flat < (long long)n_range * n_doppler) ? __ldg(m + flat) : 0.0f;
    }
    return (line >= 0 && line < n_range && col >= 0 && col < n_doppler)
               ? __ldg(m + (long long)line * n_doppler + col) : 0.0f;
}

__device__ __forceinline__ void cfar_emit(const CfarArgs& a, long long map, int line, int col, float mag,
                                          rdp_det* __restrict__ dets, int* __restrict__ counts) {
    const int cpi = (int)(map / a.n_channels), ch = (int)(map % a.n_channels);
    const int slot = atomicAdd(counts + cpi, 1);
    if (slot < a.max_dets) {
        rdp_det d;
        d.cpi = cpi; d.channel = ch; d.range = line; d.doppler = col; d.mag = mag;
        dets[(long long)cpi * a.max_dets + slot] = d;
    }
}

// Exact decision for one cell by one warp (all 32 lanes call this with the same cell).
template <int MAXQ>
__device__ __forceinline__ bool cfar_exact_warp(const CfarArgs& a, const float* __restrict__ m,
                                                int line, int col, float cut, int lane) {
    const int wc = 2 * a.half_c + 1;
    const int win = (2 * a.half_l + 1) * wc;
    float vals[MAXQ];
    bool isref[MAXQ];
    double s = 0.0;
    RDP_UNROLL
    for (int q = 0; q < MAXQ; ++q) {
        const int idx = lane + 32 * q;
        const int dl = idx / wc - a.half_l, dc = idx % wc - a.half_c;
        const bool guard = (dl >= -a.guard_l && dl <= a.guard_l && dc >= -a.guard_c && dc <= a.guard_c);
        isref[q] = (idx < win) && !guard;
        vals[q] = isref[q] ? cfar_fetch(m, line + dl, col + dc, a.n_range, a.n_doppler, a.raster) : 0.0f;
        s += (double)vals[q];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float scale = a.s1;
    if (a.adaptive) {
        const float mean = (float)(s / (double)a.n_ref);          // os_cfar_2d.vhd:189
        const float lo = __fmul_rn(0.5f, mean);                    // shift_right(mean, 1)
        const float hi = __fadd_rn(mean, lo);
        int c_le_hi = 0, c_lt_lo = 0;
        RDP_UNROLL
        for (int q = 0; q < MAXQ; ++q) {
            c_le_hi += (isref[q] && vals[q] <= hi) ? 1 : 0;
            c_lt_lo += (isref[q] && vals[q] < lo) ? 1 : 0;
        }
        c_le_hi = __reduce_add_sync(0xffffffffu, c_le_hi);
        c_lt_lo = __reduce_add_sync(0xffffffffu, c_lt_lo);
        // :193-199  x_k > hi -> MAX ; x_k < lo -> MIN ; else NOM
        scale = (c_le_hi <= a.rank_idx) ? a.smax : (c_lt_lo >= a.rank_idx + 1) ? a.smin : a.snom;
    }
    int c = 0;
    RDP_UNROLL
    for (int q = 0; q < MAXQ; ++q) c += (isref[q] && __fmul_rn(vals[q], scale) < cut) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    return c >= a.rank_idx + 1;                                    // cut > x_k * scale  (:213)
}

// ------------------------------------------------------------------------------------------
// Tiled kernel, geometry compile-time.  CTA = 256 threads = TCOLS lanes along Doppler x
// (256 / TCOLS) strips of S range lines.
// ------------------------------------------------------------------------------------------
template <int HL, int HC, int GL, int GC, int S, int TCOLS>
struct K3Shape {
    static constexpr int THREADS = 256;
    static constexpr int TROWS = (THREADS / TCOLS) * S;
    static constexpr int SROWS = TROWS + 2 * HL, SCOLS = TCOLS + 2 * HC;
    static constexpr int WIN = (2 * HL + 1) * (2 * HC + 1);
    static constexpr int MAXQ = (WIN + 31) / 32;
    static constexpr size_t smem_bytes() { return sizeof(float) * SROWS * SCOLS + sizeof(int) * (TROWS * TCOLS + 4); }
};

template <int HL, int HC, int GL, int GC, int S, int TCOLS>
__global__ void __launch_bounds__(256)
k3_cfar_kernel(const float* __restrict__ rd_map, float* __restrict__ cfar_out, rdp_det* __restrict__ dets,
               int* __restrict__ counts, CfarArgs a, int tiles_r, int tiles_c) {
    using Sh = K3Shape<HL, HC, GL, GC, S, TCOLS>;
    RDP_DYN_SMEM(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw);
    int* surv = reinterpret_cast<int*>(tile + Sh::SROWS * Sh::SCOLS);
    int* n_surv = surv + Sh::TROWS * TCOLS;

    const long long tile_id = blockIdx.x;
    const long long map = tile_id / (tiles_r * tiles_c);
    const int tr = (int)(tile_id % (tiles_r * tiles_c)) / tiles_c, tc = (int)(tile_id % tiles_c);
    const int r0 = tr * Sh::TROWS, c0 = tc * TCOLS;
    const float* m = rd_map + map * (long long)a.n_range * a.n_doppler;

    if (threadIdx.x == 0) *n_surv = 0;
    for (int i = threadIdx.x; i < Sh::SROWS * Sh::SCOLS; i += 256) {
        const int lr = i / Sh::SCOLS, lc = i % Sh::SCOLS;
        const float v = cfar_fetch(m, r0 - HL + lr, c0 - HC + lc, a.n_range, a.n_doppler, a.raster);
        tile[i] = __fmul_rn(v, a.s1);
    }
    __syncthreads();

    const int tx = threadIdx.x % TCOLS, ty = threadIdx.x / TCOLS;
    const int col = c0 + tx;
    float cut[S];
    int cnt[S];
    RDP_UNROLL
    for (int j = 0; j < S; ++j) {
        const int line = r0 + ty * S + j;
        cut[j] = (line < a.n_range && col < a.n_doppler) ? __ldg(m + (long long)line * a.n_doppler + col) : 0.0f;
        cnt[j] = 0;
    }
    const float* tp = tile + (ty * S) * Sh::SCOLS + tx + HC;
    for (int dc = -HC; dc <= HC; ++dc) {
        if (dc >= -GC && dc <= GC) {                      // columns crossing the guard box
            RDP_UNROLL
            for (int lr = 0; lr < S + 2 * HL; ++lr) {
                const float v = tp[lr * Sh::SCOLS + dc];
                RDP_UNROLL
                for (int j = 0; j < S; ++j) {
                    const int dl = lr - j - HL;
                    if (dl >= -HL && dl <= HL && !(dl >= -GL && dl <= GL)) cnt[j] += (v < cut[j]) ? 1 : 0;
                }
            }
        } else {
            RDP_UNROLL
            for (int lr = 0; lr < S + 2 * HL; ++lr) {
                const float v = tp[lr * Sh::SCOLS + dc];
                RDP_UNROLL
                for (int j = 0; j < S; ++j) {
                    const int dl = lr - j - HL;
                    if (dl >= -HL && dl <= HL) cnt[j] += (v < cut[j]) ? 1 : 0;
                }
            }
        }
    }
    RDP_UNROLL
    for (int j = 0; j < S; ++j) {
        const int line = r0 + ty * S + j;
        if (line < a.n_range && col < a.n_doppler) {
            const bool pass = cnt[j] >= a.rank_idx + 1;
            if (cfar_out) cfar_out[map * (long long)a.n_range * a.n_doppler + (long long)line * a.n_doppler + col] =
                (pass && !a.adaptive) ? cut[j] : 0.0f;
            if (pass) {
                if (a.adaptive) surv[atomicAdd(n_surv, 1)] = ((ty * S + j) << 16) | tx;
                else cfar_emit(a, map, line, col, cut[j], dets, counts);
            }
        }
    }
    if (!a.adaptive) return;
    __syncthreads();
    const int ns = *n_surv;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int i = warp; i < ns; i += 8) {
        const int code = surv[i];
        const int line = r0 + (code >> 16), c = c0 + (code & 0xffff);
        const float cv = __ldg(m + (long long)line * a.n_doppler + c);
        const bool det = cfar_exact_warp<Sh::MAXQ>(a, m, line, c, cv, lane);
        if (det && lane == 0) {
            if (cfar_out) cfar_out[map * (long long)a.n_range * a.n_doppler + (long long)line * a.n_doppler + c] = cv;
            cfar_emit(a, map, line, c, cv, dets, counts);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Generic kernel, geometry at run time (any window up to 32*MAXQ_GENERIC cells): one warp per
// cell straight from global / L2.  Slow; used only for window shapes without a tiled instance.
// ------------------------------------------------------------------------------------------
constexpr int MAXQ_GENERIC = 40;   // windows up to 1280 cells

__global__ void __launch_bounds__(256)
k3_cfar_generic_kernel(const float* __restrict__ rd_map, float* __restrict__ cfar_out,
                       rdp_det* __restrict__ dets, int* __restrict__ counts, CfarArgs a, long long n_cells) {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const long long cells_per_map = (long long)a.n_range * a.n_doppler;
    for (long long cell = (long long)blockIdx.x * 8 + warp; cell < n_cells; cell += (long long)gridDim.x * 8) {
        const long long map = cell / cells_per_map;
        const int line = (int)((cell % cells_per_map) / a.n_doppler), col = (int)(cell % a.n_doppler);
        const float* m = rd_map + map * cells_per_map;
        const float cv = __ldg(m + (long long)line * a.n_doppler + col);
        const bool det = cfar_exact_warp<MAXQ_GENERIC>(a, m, line, col, cv, lane);
        if (lane == 0) {
            if (cfar_out) cfar_out[cell] = det ? cv : 0.0f;
            if (det) cfar_emit(a, map, line, col, cv, dets, counts);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Canonical order: one CTA per CPI sorts its records by (channel, range, doppler) with a
// shared-memory bitonic network and raises the overflow bit (radar_core.vhd:447-456 analogue).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k4_sort_dets_kernel(rdp_det* __restrict__ dets, const int* __restrict__ counts, uint32_t* __restrict__ status,
                    int max_dets, int pow2_cap) {
    RDP_DYN_SMEM(smem_raw);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    float* mags = reinterpret_cast<float*>(keys + pow2_cap);
    const int cpi = blockIdx.x;
    const int found = counts[cpi];
    const int n = found < max_dets ? found : max_dets;
    if (threadIdx.x == 0) status[cpi] = (found > max_dets) ? (uint32_t)RDP_STATUS_DET_OVERFLOW : 0u;
    rdp_det* row = dets + (long long)cpi * max_dets;
    if (n <= 1) {
        if (n == 1 && threadIdx.x == 0) row[0].cpi = cpi;
        return;
    }
    int p = 2;
    while (p < n) p <<= 1;
    for (int i = threadIdx.x; i < p; i += 256) {
        if (i < n) {
            rdp_det d = row[i];
            keys[i] = ((unsigned long long)(unsigned)d.channel << 48) | ((unsigned long long)(unsigned)d.range << 24) |
                      (unsigned long long)(unsigned)d.doppler;
            mags[i] = d.mag;
        } else {
            keys[i] = ~0ull;
            mags[i] = 0.0f;
        }
    }
    __syncthreads();
    for (int k = 2; k <= p; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < p; i += 256) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long ki = keys[i], kl = keys[l];
                    if ((ki > kl) == up) {
                        keys[i] = kl; keys[l] = ki;
                        const float t = mags[i]; mags[i] = mags[l]; mags[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += 256) {
        const unsigned long long key = keys[i];
        rdp_det d;
        d.cpi = cpi;
        d.channel = (int)(key >> 48);
        d.range = (int)((key >> 24) & 0xffffffull);
        d.doppler = (int)(key & 0xffffffull);
        d.mag = mags[i];
        row[i] = d;
    }
}

}  // namespace rdp
