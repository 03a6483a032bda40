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
__global__ void __launch_bounds__(256)
kf_window_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const int* __restrict__ coef,
                 int n, long long total, uint32_t* __restrict__ sat_flag) {
    bool sat = false;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const uint32_t w = __ldg(in + idx);
        const int c = __ldg(coef + (int)(idx % n));
        const int vi = (word_i(w) * c + (1 << 14)) >> 14, vq = (word_q(w) * c + (1 << 14)) >> 14;
        const int yi = sat16(vi), yq = sat16(vq);
        sat |= (yi != vi) || (yq != vq);
        out[idx] = make_word(yi, yq);
    }
    if (sat && sat_flag) atomicOr(sat_flag, 1u);
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

}  // namespace rdp
