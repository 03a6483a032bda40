// k0_fixed.cuh -- bit-exact 16-bit fixed-point stage taps for the hand-written RTL stages whose
// arithmetic is fully defined in the reference tree (the FFT IP is not: SURVEY.md finding 0.3).
// Elementwise / transpose kernels, HBM-bound; they exist for parity with the reference's unit
// testbenches (rtl/src/tb_window_multiplier.vhd, tb_corner_turner.vhd, tb_doppler_notch.vhd,
// tb_magnitude_calc.vhd), not for the float32 throughput path.
// A sample is one 32-bit stream word {Q[31:16], I[15:0]} (rtl/src/window_multiplier.vhd:93-94).
#pragma once
#include "rdp_common.cuh"

namespace rdp {

__device__ __forceinline__ int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
__device__ __forceinline__ int word_i(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int word_q(uint32_t w) { return (int)w >> 16; }
__device__ __forceinline__ uint32_t make_word(int i, int q) { return ((uint32_t)q << 16) | ((uint32_t)i & 0xffffu); }

// window_multiplier: y = sat16((x * coef[n] + 2^14) >> 14) per component; coef[n] is the mirrored Q15
// Hamming ROM (rtl/src/window_multiplier.vhd:34-49,97-104,126-127,146-158).  *sat_flag |= 1 on clipping.
// sat_flags[idx / words_per_flag] |= flag_bits (one flag word per CPI in the chain, one for the whole call in the tap).
__global__ void __launch_bounds__(256)
kf_window_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const int* __restrict__ coef,
                 int n, long long total, uint32_t* __restrict__ sat_flags, long long words_per_flag, uint32_t flag_bits) {
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const uint32_t w = __ldg(in + idx);
        const int c = __ldg(coef + (int)(idx % n));
        const int vi = (word_i(w) * c + (1 << 14)) >> 14, vq = (word_q(w) * c + (1 << 14)) >> 14;
        const int yi = sat16(vi), yq = sat16(vq);
        if (((yi != vi) || (yq != vq)) && sat_flags) atomicOr(sat_flags + idx / words_per_flag, flag_bits);
        out[idx] = make_word(yi, yq);
    }
}

// doppler_notch: per line of n words, y[k] = sat16(x[k] - x[k-1]) (mode 2) or sat16(x[k] - 2x[k-1] + x[k-2])
// (mode 3) with a zero delay line at the start of each line (rtl/src/doppler_notch.vhd:70-104).
__global__ void __launch_bounds__(256)
kf_mti_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n, long long total, int mode) {
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int k = (int)(idx % n);
        const uint32_t w = __ldg(in + idx);
        const uint32_t d1 = k >= 1 ? __ldg(in + idx - 1) : 0u, d2 = k >= 2 ? __ldg(in + idx - 2) : 0u;
        int yi, yq;
        if (mode == 2) {
            yi = word_i(w) - word_i(d1);
            yq = word_q(w) - word_q(d1);
        } else if (mode == 3) {
            yi = word_i(w) - 2 * word_i(d1) + word_i(d2);
            yq = word_q(w) - 2 * word_q(d1) + word_q(d2);
        } else {                                           // bypass
            yi = word_i(w);
            yq = word_q(w);
        }
        out[idx] = make_word(sat16(yi), sat16(yq));
    }
}

// magnitude_calc: max(|I|,|Q|) + (min >> 2) + (min >> 3), 17-bit unsigned (rtl/src/magnitude_calc.vhd:55-60,78-81)
__global__ void __launch_bounds__(256)
kf_magnitude_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long total) {
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const uint32_t w = __ldg(in + idx);
        int ai = word_i(w), aq = word_q(w);
        ai = ai < 0 ? -ai : ai;
        aq = aq < 0 ? -aq : aq;
        const int mx = ai >= aq ? ai : aq, mn = ai >= aq ? aq : ai;
        out[idx] = (uint32_t)(mx + (mn >> 2) + (mn >> 3));
    }
}

// corner_turner: out[map][r][d] = in[map][d][r] on 32-bit words (rtl/src/corner_turner.vhd:79-80), 32x32 tiles
// through shared memory so that both sides are coalesced.
__global__ void __launch_bounds__(256)
kf_corner_turn_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n_chirps, int n_samples) {
    __shared__ uint32_t t[32][33];
    const long long map = blockIdx.z;
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const uint32_t* src = in + map * (long long)n_chirps * n_samples;
    uint32_t* dst = out + map * (long long)n_chirps * n_samples;
    for (int j = ty; j < 32; j += 8)
        if (c0 + j < n_chirps && s0 + tx < n_samples) t[j][tx] = __ldg(src + (long long)(c0 + j) * n_samples + s0 + tx);
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (s0 + j < n_samples && c0 + tx < n_chirps) dst[(long long)(s0 + j) * n_chirps + c0 + tx] = t[tx][j];
}

// ------------------------------------------------------------------------------------------
// Block-floating-point FFT of the end-to-end fixed-point mode.  The reference's FFTs are Xilinx xfft v9.1 cores
// (16-bit data, 16-bit twiddles, block floating point, convergent rounding: .../ip/xfft_0_1/xfft_0.xci:12-27,59)
// whose RTL is not in the reference tree; this is the SELF-DEFINED transform with the same word widths and rounding
// rule that oracle/rtl_fixed.py:bfp_fft states (and is tested bit for bit against): log2(N) radix-2 decimation-in-
// frequency stages on 16-bit words, exact integer butterflies a' = a + b, b' = cround((a - b) W, 15) with Q15
// twiddles, then ONE convergent shift per stage for the whole frame, chosen from the frame's largest component so
// that everything fits 16 bits again; the shifts add up to the frame's block exponent, which is RETURNED (the RTL
// drops it: radar_core.vhd:310,358).  One warp per frame, the frame in shared memory, warp-synchronous.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long cround_ll(long long v, int sh) {       // convergent (half-to-even) arithmetic >> sh
    if (sh == 0) return v;
    const long long q = v >> sh, rem = v - (q << sh), half = 1ll << (sh - 1);
    return q + ((rem > half || (rem == half && (q & 1))) ? 1 : 0);
}
struct BfpButterfly { int sr, si, pr, pi; };
__device__ __forceinline__ BfpButterfly bfp_butterfly(uint32_t a, uint32_t b, uint32_t w) {
    const int ar = word_i(a), ai = word_q(a), br = word_i(b), bi = word_q(b);
    const long long dr = ar - br, di = ai - bi, cr = word_i(w), ci = word_q(w);
    BfpButterfly o;
    o.sr = ar + br;
    o.si = ai + bi;
    o.pr = (int)cround_ll(dr * cr - di * ci, 15);
    o.pi = (int)cround_ll(dr * ci + di * cr, 15);
    return o;
}
__device__ __forceinline__ int iabs_(int v) { return v < 0 ? -v : v; }

// Fusions of the end-to-end chain (all nullable / 0 = plain transform, the rdp_fixed_bfp_fft tap): on load, doppler_notch
// along the frame (mti_mode 2 / 3, kf_mti_kernel's arithmetic) and then window_multiplier (coef = the mirrored ROM,
// kf_window_kernel's arithmetic, saturation ORed into sat_flags[frame / frames_per_flag]); on store, magnitude_calc
// (mag_out: the frame leaves as 17-bit unsigned magnitudes, kf_magnitude_kernel's arithmetic).
struct BfpFuse {
    int adc_check;             // raise bit 2 (RDP_STATUS_ADC_FULL_SCALE) in the frame's flag word when an input word sits on a rail
    const int* coef;
    uint32_t* sat_flags;
    long long frames_per_flag;
    uint32_t flag_bits;
    int mti_mode;
    int mag_out;
};

// NB > 0: n == 64 * NB (NB butterflies per lane and stage, kept in registers); NB == 0: any n, butterflies recomputed
template <int NB>
__global__ void __launch_bounds__(256)
kf_bfp_fft_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int* __restrict__ exps,
                  const uint32_t* __restrict__ tw, int n, int log2n, long long n_frames, int warps_per_cta, BfpFuse fu) {
    RDP_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint32_t* fb = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * n;
    if (warp >= warps_per_cta) return;
    for (long long f = (long long)blockIdx.x * warps_per_cta + warp; f < n_frames; f += (long long)gridDim.x * warps_per_cta) {
        const uint32_t* src = in + f * n;
        bool sat = false, rail = false;
        for (int i = lane; i < n; i += 32) {
            uint32_t w = __ldg(src + i);
            rail |= word_i(w) >= 32767 || word_i(w) <= -32768 || word_q(w) >= 32767 || word_q(w) <= -32768;
            if (fu.mti_mode == 2 || fu.mti_mode == 3) {
                const uint32_t d1 = i >= 1 ? __ldg(src + i - 1) : 0u, d2 = i >= 2 ? __ldg(src + i - 2) : 0u;
                const int yi = fu.mti_mode == 2 ? word_i(w) - word_i(d1) : word_i(w) - 2 * word_i(d1) + word_i(d2);
                const int yq = fu.mti_mode == 2 ? word_q(w) - word_q(d1) : word_q(w) - 2 * word_q(d1) + word_q(d2);
                w = make_word(sat16(yi), sat16(yq));
            }
            if (fu.coef) {
                const int c = __ldg(fu.coef + i);
                const int vi = (word_i(w) * c + (1 << 14)) >> 14, vq = (word_q(w) * c + (1 << 14)) >> 14;
                const int yi = sat16(vi), yq = sat16(vq);
                sat |= (yi != vi) || (yq != vq);
                w = make_word(yi, yq);
            }
            fb[i] = w;
        }
        if (fu.sat_flags) {
            const uint32_t bits = (__any_sync(0xffffffffu, sat) ? fu.flag_bits : 0u) | ((fu.adc_check && __any_sync(0xffffffffu, rail)) ? 4u : 0u);
            if (bits && lane == 0) atomicOr(fu.sat_flags + f / fu.frames_per_flag, bits);
        }
        __syncwarp();
        int e = 0;
        for (int s = 0; s < log2n; ++s) {
            const int half = n >> (s + 1);
            int sh = 0;
            if (NB > 0) {
                // n = 64 NB points: the lane's NB butterflies of the stage stay in registers between the pass that finds
                // the frame's largest component and the one that shifts -- each butterfly is computed once
                BfpButterfly o[NB > 0 ? NB : 1];
                int i0s[NB > 0 ? NB : 1];
                int g = 0;
                RDP_UNROLL
                for (int q = 0; q < NB; ++q) {
                    const int j = lane + 32 * q;
                    const int pos = j & (half - 1);
                    i0s[q] = ((j >> (log2n - 1 - s)) << (log2n - s)) + pos;
                    o[q] = bfp_butterfly(fb[i0s[q]], fb[i0s[q] + half], __ldg(tw + (pos << s)));
                    const int m1 = iabs_(o[q].sr) > iabs_(o[q].si) ? iabs_(o[q].sr) : iabs_(o[q].si);
                    const int m2 = iabs_(o[q].pr) > iabs_(o[q].pi) ? iabs_(o[q].pr) : iabs_(o[q].pi);
                    g = g > m1 ? g : m1;
                    g = g > m2 ? g : m2;
                }
                g = __reduce_max_sync(0xffffffffu, g);
                while (((g + (1 << sh) - 1) >> sh) > 32767) ++sh;
                RDP_UNROLL
                for (int q = 0; q < NB; ++q) {
                    fb[i0s[q]] = make_word((int)cround_ll(o[q].sr, sh), (int)cround_ll(o[q].si, sh));
                    fb[i0s[q] + half] = make_word((int)cround_ll(o[q].pr, sh), (int)cround_ll(o[q].pi, sh));
                }
            } else {
                int g = 0;
                for (int j = lane; j < n / 2; j += 32) {
                    const int pos = j & (half - 1), i0 = ((j >> (log2n - 1 - s)) << (log2n - s)) + pos;
                    const BfpButterfly o = bfp_butterfly(fb[i0], fb[i0 + half], __ldg(tw + (pos << s)));
                    const int m1 = iabs_(o.sr) > iabs_(o.si) ? iabs_(o.sr) : iabs_(o.si);
                    const int m2 = iabs_(o.pr) > iabs_(o.pi) ? iabs_(o.pr) : iabs_(o.pi);
                    g = g > m1 ? g : m1;
                    g = g > m2 ? g : m2;
                }
                g = __reduce_max_sync(0xffffffffu, g);
                while (((g + (1 << sh) - 1) >> sh) > 32767) ++sh;
                for (int j = lane; j < n / 2; j += 32) {                 // every butterfly rewrites its own two words
                    const int pos = j & (half - 1), i0 = ((j >> (log2n - 1 - s)) << (log2n - s)) + pos;
                    const BfpButterfly o = bfp_butterfly(fb[i0], fb[i0 + half], __ldg(tw + (pos << s)));
                    fb[i0] = make_word((int)cround_ll(o.sr, sh), (int)cround_ll(o.si, sh));
                    fb[i0 + half] = make_word((int)cround_ll(o.pr, sh), (int)cround_ll(o.pi, sh));
                }
            }
            __syncwarp();
            e += sh;
        }
        for (int k = lane; k < n; k += 32) {                         // bit-reversed -> natural order, coalesced stores
            const uint32_t w = fb[(int)(__brev((unsigned)k) >> (32 - log2n))];
            uint32_t o = w;
            if (fu.mag_out) {
                int ai = word_i(w), aq = word_q(w);
                ai = ai < 0 ? -ai : ai;
                aq = aq < 0 ? -aq : aq;
                const int mx = ai >= aq ? ai : aq, mn = ai >= aq ? aq : ai;
                o = (uint32_t)(mx + (mn >> 2) + (mn >> 3));
            }
            out[f * n + k] = o;
        }
        if (lane == 0 && exps) exps[f] = e;
        __syncwarp();
    }
}

}  // namespace rdp
