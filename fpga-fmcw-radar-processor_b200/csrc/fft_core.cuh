// fft_core.cuh -- register-resident Stockham FFT building blocks (forward DFT, natural order).
//
// The reference's two FFTs are Xilinx xfft v9.1 cores configured forward / natural order
// (rtl/src/radar_core.vhd:246-248,303-316,351-364).  Here an N-point transform is done by a
// group of T = N/E threads, each holding E complex points in registers.  The transform is a
// sequence of Stockham passes of radix R_p (prod R_p = N, R_p | E).  In the pass with
// Ns = prod of the previous radices, butterfly j (0 <= j < N/R) does
//     v[r]  = in[j + r*N/R] * W_{Ns*R}^{(j mod Ns) r}        r = 0..R-1
//     v     = DFT_R(v)
//     out[(j div Ns)*Ns*R + (j mod Ns) + r*Ns] = v[r]
// and after the last pass `out` is the DFT in natural order.  A thread owns the E/R
// butterflies j = tid + b*T.  Its twiddles depend only on (tid, b, r), so a persistent
// thread keeps them in registers across all the transforms it performs.  Passes exchange
// data through a per-group shared-memory buffer (double-buffered: one barrier per exchange).
#pragma once
#include "rdp_common.cuh"
#include <type_traits>

namespace rdp {

#define RDP_SQRT1_2 0.70710678118654752440f
#define RDP_COS_PI_8 0.92387953251128675613f
#define RDP_SIN_PI_8 0.38268343236508977173f

// v * W16^M, W16 = exp(-j 2 pi / 16), M compile-time
template <int M>
__device__ __forceinline__ float2 mul_w16(float2 a) {
    constexpr int m = ((M % 16) + 16) % 16;
    if (m == 0) return a;
    if (m == 4) return make_float2(a.y, -a.x);
    if (m == 8) return make_float2(-a.x, -a.y);
    if (m == 12) return make_float2(-a.y, a.x);
    if (m == 2) return make_float2(RDP_SQRT1_2 * (a.x + a.y), RDP_SQRT1_2 * (a.y - a.x));
    if (m == 6) return make_float2(RDP_SQRT1_2 * (a.y - a.x), -RDP_SQRT1_2 * (a.x + a.y));
    if (m == 10) return make_float2(-RDP_SQRT1_2 * (a.x + a.y), RDP_SQRT1_2 * (a.x - a.y));
    if (m == 14) return make_float2(RDP_SQRT1_2 * (a.x - a.y), RDP_SQRT1_2 * (a.x + a.y));
    // odd powers: (c, -s) with c = cos(2 pi m/16), s = sin(2 pi m/16)
    constexpr float c = (m == 1 || m == 15) ? RDP_COS_PI_8 : (m == 3 || m == 13) ? RDP_SIN_PI_8
                      : (m == 5 || m == 11) ? -RDP_SIN_PI_8 : -RDP_COS_PI_8;
    constexpr float s = (m == 1 || m == 7) ? RDP_SIN_PI_8 : (m == 3 || m == 5) ? RDP_COS_PI_8
                      : (m == 9 || m == 15) ? -RDP_SIN_PI_8 : -RDP_COS_PI_8;
    return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}

__device__ __forceinline__ void dft2(float2& a, float2& b) {
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// natural-order 4-point forward DFT of (a0, a1, a2, a3)
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    float2 t2 = cadd(a1, a3), t3 = cmul_mj(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

template <int R>
struct Dft;

template <>
struct Dft<2> {
    static __device__ __forceinline__ void run(float2* v) { dft2(v[0], v[1]); }
};
template <>
struct Dft<4> {
    static __device__ __forceinline__ void run(float2* v) { dft4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Dft<8> {
    // n = 4*n1 + n2, k = k1 + 2*k2:  DFT2 over n1, twiddle W8^{n2 k1}, DFT4 over n2
    static __device__ __forceinline__ void run(float2* v) {
        float2 a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
        float2 b0 = v[4], b1 = v[5], b2 = v[6], b3 = v[7];
        dft2(a0, b0);
        dft2(a1, b1);
        dft2(a2, b2);
        dft2(a3, b3);
        b1 = mul_w16<2>(b1);
        b2 = mul_w16<4>(b2);
        b3 = mul_w16<6>(b3);
        dft4(a0, a1, a2, a3);   // X[0], X[2], X[4], X[6]
        dft4(b0, b1, b2, b3);   // X[1], X[3], X[5], X[7]
        v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
        v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
    }
};
template <>
struct Dft<16> {
    // n = 4*n1 + n2, k = k1 + 4*k2:  DFT4 over n1, twiddle W16^{n2 k1}, DFT4 over n2
    static __device__ __forceinline__ void run(float2* v) {
        float2 a[4][4];   // a[n2][k1]
        RDP_UNROLL
        for (int n2 = 0; n2 < 4; ++n2) {
            a[n2][0] = v[n2]; a[n2][1] = v[4 + n2]; a[n2][2] = v[8 + n2]; a[n2][3] = v[12 + n2];
            dft4(a[n2][0], a[n2][1], a[n2][2], a[n2][3]);
        }
        a[1][1] = mul_w16<1>(a[1][1]); a[1][2] = mul_w16<2>(a[1][2]); a[1][3] = mul_w16<3>(a[1][3]);
        a[2][1] = mul_w16<2>(a[2][1]); a[2][2] = mul_w16<4>(a[2][2]); a[2][3] = mul_w16<6>(a[2][3]);
        a[3][1] = mul_w16<3>(a[3][1]); a[3][2] = mul_w16<6>(a[3][2]); a[3][3] = mul_w16<9>(a[3][3]);
        RDP_UNROLL
        for (int k1 = 0; k1 < 4; ++k1) {
            dft4(a[0][k1], a[1][k1], a[2][k1], a[3][k1]);   // -> k2 = 0..3
            v[k1] = a[0][k1]; v[k1 + 4] = a[1][k1]; v[k1 + 8] = a[2][k1]; v[k1 + 12] = a[3][k1];
        }
    }
};

// One Stockham pass.  N points, E per thread, radix R, NS = product of earlier radices.
template <int N, int E, int R_, int NS>
struct FftPass {
    static constexpr int R = R_;
    static constexpr int NSV = NS;        // product of the earlier radices
    static constexpr int T = N / E;       // threads per transform
    static constexpr int NB = E / R;      // butterflies per thread
    static constexpr int NTW = (NS > 1) ? NB * (R - 1) : 0;
    static_assert(N % E == 0 && E % R == 0 && N % (NS * R) == 0, "bad FFT pass");

    static __device__ __forceinline__ int in_index(int tid, int b, int r) { return tid + b * T + r * (N / R); }
    static __device__ __forceinline__ int out_index(int tid, int b, int r) {
        int j = tid + b * T;
        return (j / NS) * (NS * R) + (j % NS) + r * NS;
    }
    // table = exp(-j 2 pi m / N), m = 0..N-1
    static __device__ __forceinline__ void load_twiddles(float2* tw, const float2* __restrict__ table, int tid) {
        if constexpr (NS > 1) {
            RDP_UNROLL
            for (int b = 0; b < NB; ++b) {
                int k = (tid + b * T) % NS;
                RDP_UNROLL
                for (int r = 1; r < R; ++r) tw[b * (R - 1) + r - 1] = table[k * r * (N / (NS * R))];
            }
        }
    }
    static __device__ __forceinline__ void butterflies(float2* v, const float2* tw) {
        RDP_UNROLL
        for (int b = 0; b < NB; ++b) {
            if constexpr (NS > 1) {
                RDP_UNROLL
                for (int r = 1; r < R; ++r) v[b * R + r] = cmul(v[b * R + r], tw[b * (R - 1) + r - 1]);
            }
            Dft<R>::run(v + b * R);
        }
    }
};

// Compile-time plan: up to four passes (radix 1 = unused).
template <int N_, int E_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1, int DB_ = 1>
struct FftPlan {
    static constexpr int DB = DB_;       // 1: two exchange buffers, one barrier per exchange; 0: one buffer, two barriers
    static constexpr int N = N_, E = E_, R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int T = N / E;
    static constexpr int NPASS = 1 + (R1 > 1) + (R2 > 1) + (R3 > 1);
    static_assert(R0 * R1 * R2 * R3 == N, "radices must multiply to N");
    using P0 = FftPass<N, E, R0, 1>;
    using P1 = FftPass<N, E, (R1 > 1 ? R1 : R0), (R1 > 1 ? R0 : 1)>;             // unused passes alias P0
    using P2 = FftPass<N, E, (R2 > 1 ? R2 : R0), (R2 > 1 ? R0 * R1 : 1)>;
    using P3 = FftPass<N, E, (R3 > 1 ? R3 : R0), (R3 > 1 ? R0 * R1 * R2 : 1)>;
    static constexpr int NTW = (R1 > 1 ? P1::NTW : 0) + (R2 > 1 ? P2::NTW : 0) + (R3 > 1 ? P3::NTW : 0);
    static constexpr int TW1 = 0;
    static constexpr int TW2 = (R1 > 1 ? P1::NTW : 0);
    static constexpr int TW3 = TW2 + (R2 > 1 ? P2::NTW : 0);
    // Exchange buffer (per transform, per parity) in float2.  The layout is chosen per exchange from the pass that
    // WRITES it, so that its stores are bank-conflict free (64-bit accesses go out per half-warp):
    //  * first pass (NS = 1): lane j writes j R + r -> one pad slot per 16 elements;
    //  * a pass with 2 <= NS < 16: lanes write runs of NS elements that start NS R apart; NS pad slots per NS R
    //    elements put consecutive runs on consecutive banks (without them the runs of a half-warp collide:
    //    measured 2-way conflicts on every store of the second exchange of the 512- and 256-point plans);
    //  * NS >= 16: a run fills the half-warp, any layout works.
    // The reading pass takes consecutive elements per lane, which stay consecutive within every padded block.
    template <class PA>
    static __host__ __device__ constexpr bool run_padded() {
        return PA::NSV >= 2 && PA::NSV < 16 && (PA::NSV * PA::R) % 16 == 0;
    }
    template <class PA>
    static __host__ __device__ constexpr int XPAD(int i) {
        if constexpr (run_padded<PA>()) return i + (i / (PA::NSV * PA::R)) * PA::NSV;
        else return i + (i >> 4);
    }
    template <class PA>
    static __host__ __device__ constexpr int xextent() { return XPAD<PA>(N - 1) + 2; }
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int XBUF = (NPASS > 1) ? cmax(xextent<P0>(), cmax(R2 > 1 ? xextent<P1>() : 0, R3 > 1 ? xextent<P2>() : 0)) : 0;
    static constexpr int XBUFS = (1 + DB) * XBUF;        // float2 per transform group
    using LAST = typename std::conditional<(R3 > 1), P3, typename std::conditional<(R2 > 1), P2,
                 typename std::conditional<(R1 > 1), P1, P0>::type>::type>::type;

    static __device__ __forceinline__ void load_twiddles(float2* tw, const float2* __restrict__ table, int tid) {
        if constexpr (R1 > 1) P1::load_twiddles(tw + TW1, table, tid);
        if constexpr (R2 > 1) P2::load_twiddles(tw + TW2, table, tid);
        if constexpr (R3 > 1) P3::load_twiddles(tw + TW3, table, tid);
    }

    template <class PA, class PB, int GT>
    static __device__ __forceinline__ void exchange(float2* v, float2* xb, int tid, int gid) {
        if constexpr (!DB) group_sync<GT>(gid);            // previous readers of the single buffer are done
        RDP_UNROLL
        for (int b = 0; b < PA::NB; ++b) {
            RDP_UNROLL
            for (int r = 0; r < PA::R; ++r) xb[XPAD<PA>(PA::out_index(tid, b, r))] = v[b * PA::R + r];
        }
        group_sync<GT>(gid);
        RDP_UNROLL
        for (int b = 0; b < PB::NB; ++b) {
            RDP_UNROLL
            for (int r = 0; r < PB::R; ++r) v[b * PB::R + r] = xb[XPAD<PA>(PB::in_index(tid, b, r))];
        }
    }

    // v holds in[P0::in_index(tid, b, r)] at v[b*R0 + r] on entry; on exit v[b*RL + r] is
    // X[LAST::out_index(tid, b, r)].  xbuf = this transform's two exchange buffers
    // (2 * XBUF float2); `parity` alternates per exchange so one barrier per exchange suffices.
    template <int GT>
    static __device__ __forceinline__ void run(float2* v, const float2* tw, float2* xbuf, int& parity,
                                               int tid, int gid) {
        P0::butterflies(v, tw);
        if constexpr (R1 > 1) {
            exchange<P0, P1, GT>(v, xbuf + parity * XBUF, tid, gid);
            parity ^= DB;
            P1::butterflies(v, tw + TW1);
        }
        if constexpr (R2 > 1) {
            exchange<P1, P2, GT>(v, xbuf + parity * XBUF, tid, gid);
            parity ^= DB;
            P2::butterflies(v, tw + TW2);
        }
        if constexpr (R3 > 1) {
            exchange<P2, P3, GT>(v, xbuf + parity * XBUF, tid, gid);
            parity ^= DB;
            P3::butterflies(v, tw + TW3);
        }
    }
};

}  // namespace rdp
