// k2_doppler.cuh -- K2: (MTI) (x) Doppler window (x) Doppler FFT (x) magnitude.
//
// Replaces doppler_notch -> window_multiplier (N_SAMPLES = N_DOPPLER) -> xfft_doppler ->
// magnitude_calc of rtl/src/radar_core.vhd:329-374.  One FFT group per range line of the
// range-major intermediate; the magnitude is formed in registers and written as float32
// [range][doppler] (Doppler bin 0 = zero Doppler, natural order).  rows_per_map > 0 leaves pad_rows lines
// free around every rows_per_map output lines (the CFAR halo of the range-sharded mode).
//   algorithmic HBM bytes / cell: 8 read + 4 written.
// COMBINE (optional stage, not in the reference: SURVEY 8f row 3): non-coherent integration over the receive
// channels -- the FFT group walks the n_combine channels of one (CPI, range line), adds the magnitudes in registers
// (float32, channel 0 first) and writes ONE line of the combined map [cpi][range][doppler]: 8 read + 4 / n_combine
// written per input cell, and the CFAR stage then sees one map per CPI instead of n_combine.
#pragma once
#include "fft_core.cuh"

namespace rdp {

template <class Plan>
struct K2Shape {
    static constexpr int THREADS = (Plan::T >= 256) ? Plan::T : 256;
    static constexpr int G = THREADS / Plan::T;
    static __host__ __device__ constexpr size_t smem_bytes() {
        return sizeof(float2) * ((size_t)G * Plan::XBUFS) + 16;
    }
};

template <int MAG>
__device__ __forceinline__ float magnitude(float2 z) {
    if (MAG == 0) {                                      // alpha-max-beta-min, magnitude_calc.vhd:55-60,78-81
        float ai = fabsf(z.x), aq = fabsf(z.y);
        float mx = fmaxf(ai, aq), mn = fminf(ai, aq);
        return mx + 0.375f * mn;                         // min>>2 + min>>3 == 0.375 * min
    }
    float p = z.x * z.x + z.y * z.y;
    if (MAG == 1) return sqrtf(p);
    if (MAG == 2) return p;
    return 10.0f * log10f(fmaxf(p, 1e-30f));
}

// COMBINE: 0 = one map per channel; 1 = non-coherent integration (above); 2 = coherent beams of a uniform linear array:
// the n_combine channels of one (CPI, range line) are summed with the steering phases exp(-j 2 pi beam c / n_beams)
// BEFORE the transform (it is linear, so this equals an angle FFT across the channels after it) and ONE Doppler
// transform per beam follows.  Beams b and b + n_beams / 2 differ by the sign of the odd channels only, so one pass over
// the channels (even and odd channels summed separately) yields both: work items are (cpi, range, beam pair) with the
// pair fastest, so that the n_beams / 2 passes over a line's channels hit L1 / L2; maps land in [cpi][beam][range][doppler].
template <class Plan, int MAG, int COMBINE = 0>
__global__ void __launch_bounds__(K2Shape<Plan>::THREADS)
k2_doppler_kernel(const float2* __restrict__ x_rc, float* __restrict__ rd_map,
                  const float* __restrict__ window, const float2* __restrict__ twtab,
                  long long n_rows, int mti_mode, int rows_per_map, int pad_rows, int n_combine, int n_range,
                  int n_beams, const float2* __restrict__ steer) {
    // COMBINE 1: n_rows counts OUTPUT lines (cpi, range); COMBINE 2: work items (cpi, range, beam pair), n_beams >= 2.
    // Input line of channel c = (cpi * n_combine + c) * n_range + range.  steer[m] = exp(-j 2 pi m / n_beams).
    constexpr int N = Plan::N, E = Plan::E, T = Plan::T;
    constexpr int G = K2Shape<Plan>::G;
    using P0 = typename Plan::P0;
    using PL = typename Plan::LAST;
    RDP_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x % T, gid = threadIdx.x / T;
    float2* xbuf = reinterpret_cast<float2*>(smem_raw) + (size_t)gid * Plan::XBUFS;

    float2 tw[Plan::NTW > 0 ? Plan::NTW : 1];
    Plan::load_twiddles(tw, twtab, tid);
    float wv[E];
    RDP_UNROLL
    for (int b = 0; b < P0::NB; ++b) {
        RDP_UNROLL
        for (int r = 0; r < P0::R; ++r) wv[b * P0::R + r] = __ldg(window + P0::in_index(tid, b, r));
    }
    int parity = 0;
    const long long stride = (long long)gridDim.x * G;

    // software pipeline: the next line's samples are in flight during the butterflies
    float2 raw[E];
    const int n_pairs = n_beams > 1 ? n_beams / 2 : 1;
    auto line_of = [&](long long row) -> long long { return COMBINE == 2 ? row / n_pairs : row; };   // (cpi, range) index
    auto in_row = [&](long long row, int ch) -> long long {
        if (COMBINE == 0) return row;
        const long long line = line_of(row), cpi = line / n_range;
        return (cpi * n_combine + ch) * n_range + (line - cpi * n_range);
    };
    auto fetch = [&](long long row, int ch) {
        const bool ok = row < n_rows;
        const float2* src = x_rc + in_row(ok ? row : 0, ch) * N;
        RDP_UNROLL
        for (int b = 0; b < P0::NB; ++b) {
            RDP_UNROLL
            for (int r = 0; r < P0::R; ++r)
                raw[b * P0::R + r] = ok ? __ldg(src + P0::in_index(tid, b, r)) : make_float2(0.f, 0.f);
        }
    };
    fetch((long long)blockIdx.x * G + gid, 0);

    for (long long row0 = (long long)blockIdx.x * G; row0 < n_rows; row0 += stride) {
        const long long row = row0 + gid;
        const bool active = row < n_rows;
        const int beam = COMBINE == 2 ? (int)((active ? row : 0) % n_pairs) : 0;     // the pair's first beam
        float macc[COMBINE == 1 ? E : 1];                            // COMBINE 1: sum of the channels' magnitudes
        float2 bacc[COMBINE == 2 ? E : 1], bodd[COMBINE == 2 ? E : 1];   // COMBINE 2: steered sums of the even / odd channels
        if (COMBINE == 1) {
            RDP_UNROLL
            for (int i = 0; i < (COMBINE == 1 ? E : 1); ++i) macc[i] = 0.0f;
        }
        if (COMBINE == 2) {
            RDP_UNROLL
            for (int i = 0; i < (COMBINE == 2 ? E : 1); ++i) { bacc[i] = make_float2(0.0f, 0.0f); bodd[i] = make_float2(0.0f, 0.0f); }
        }
      for (int ch = 0; ch < (COMBINE ? n_combine : 1); ++ch) {
        float2 v[E];                                     // the line's samples after the notch and (COMBINE != 2) the window
        if (mti_mode == 0) {
            RDP_UNROLL
            for (int i = 0; i < E; ++i) v[i] = COMBINE == 2 ? raw[i] : make_float2(raw[i].x * wv[i], raw[i].y * wv[i]);
        } else {                                         // doppler_notch.vhd:72-93, zero initial state (:99-102)
            const float2* src = x_rc + in_row(active ? row : 0, ch) * N;
            RDP_UNROLL
            for (int b = 0; b < P0::NB; ++b) {
                RDP_UNROLL
                for (int r = 0; r < P0::R; ++r) {
                    const int n = P0::in_index(tid, b, r);
                    float2 s = raw[b * P0::R + r];
                    const float2 d1 = (active && n >= 1) ? __ldg(src + n - 1) : make_float2(0.f, 0.f);
                    if (mti_mode == 2) {
                        s = csub(s, d1);
                    } else {
                        const float2 d2 = (active && n >= 2) ? __ldg(src + n - 2) : make_float2(0.f, 0.f);
                        s = make_float2(s.x - 2.0f * d1.x + d2.x, s.y - 2.0f * d1.y + d2.y);
                    }
                    const float w = COMBINE == 2 ? 1.0f : wv[b * P0::R + r];
                    v[b * P0::R + r] = COMBINE == 2 ? s : make_float2(s.x * w, s.y * w);
                }
            }
        }
        if (COMBINE && ch + 1 < n_combine) fetch(row, ch + 1);
        else fetch(row + stride, 0);
        if (COMBINE == 2) {
            const float2 ph = __ldg(steer + ((beam * ch) & (n_beams - 1)));
            if (ch & 1) {
                RDP_UNROLL
                for (int i = 0; i < (COMBINE == 2 ? E : 1); ++i) bodd[i] = cadd(bodd[i], cmul(v[i], ph));
            } else {
                RDP_UNROLL
                for (int i = 0; i < (COMBINE == 2 ? E : 1); ++i) bacc[i] = cadd(bacc[i], cmul(v[i], ph));
            }
            if (ch + 1 < n_combine) continue;
        }
      // COMBINE 2: two transforms per pass over the channels -- beam (even + odd sums), then beam + n_beams / 2 (even - odd)
      for (int half = 0; half < (COMBINE == 2 && n_beams > 1 ? 2 : 1); ++half) {
        if (COMBINE == 2) {
            RDP_UNROLL
            for (int i = 0; i < (COMBINE == 2 ? E : 1); ++i) {
                const float2 t = half ? csub(bacc[i], bodd[i]) : cadd(bacc[i], bodd[i]);
                v[i] = make_float2(t.x * wv[i], t.y * wv[i]);          // the window, once per beam
            }
        }
        Plan::template run<T>(v, tw, xbuf, parity, tid, gid);
        if (COMBINE == 1) {
            RDP_UNROLL
            for (int i = 0; i < (COMBINE == 1 ? E : 1); ++i) macc[i] += magnitude<MAG>(v[i]);
            if (ch + 1 < n_combine) continue;
        }
        if (active) {
            // sharded mode: every map of rows_per_map lines is written with pad_rows halo lines above and below
            long long orow = rows_per_map > 0 ? row + (2 * (row / rows_per_map) + 1) * pad_rows : row;
            if (COMBINE == 2) {                                      // (cpi, range, beam) -> [cpi][beam][range]
                const long long line = row / n_pairs, cpi = line / n_range;
                orow = (cpi * n_beams + beam + half * n_pairs) * n_range + (line - cpi * n_range);
            }
            float* dst = rd_map + orow * N;
            RDP_UNROLL
            for (int b = 0; b < PL::NB; ++b) {
                RDP_UNROLL
                for (int r = 0; r < PL::R; ++r)
                    dst[PL::out_index(tid, b, r)] = COMBINE == 1 ? macc[b * PL::R + r] : magnitude<MAG>(v[b * PL::R + r]);
            }
        }
      }   // beams of the pair
      }   // channels
    }
}

}  // namespace rdp
