// k3_cfar.cuh -- K3: 2D ordered-statistic CFAR, and K4: ordered compaction to detection records.
//
// Replaces os_cfar_2d (rtl/src/os_cfar_2d.vhd:140-217) and the non-zero filter / coordinate
// counters of rtl/src/radar_core.vhd:396-418.  The RTL bubble-sorts the N_REF reference cells
// of every cell under test (:172-178).  Here nothing is sorted: with k = rank_idx,
//     cut > s * x_(k)   <=>   #{ref : fl(ref * s) < cut} >= k + 1
//     x_(k) > a         <=>   #{ref <= a} <= k
//     x_(k) < a         <=>   #{ref <  a} >= k + 1
// (identities checked in tests/test_oracle_fixed.py).
//
// A tile of the map (+ halo) is staged in shared memory twice: raw 32-bit words (TMA), and as
// bfloat16 TRUNCATIONS packed two VERTICALLY adjacent cells per word (rows 2i, 2i+1 of one column),
// column-major in two planes (even / odd columns) so that a thread reads the rows of a column with
// 128-bit loads, conflict-free.  One packed compare (HSET2.BF16_V2 -> 0xFFFF masks, summed two at a
// time by IADD3) then tests two reference cells against one threshold; references are rounded
// DOWN and thresholds UP, so a packed count never undercounts.  Three phases, each a strict
// (conservative) pre-filter of the next; only the last one decides:
//
// Phase 0 (every cell) -- SHARED threshold per strip, exploiting the window overlap.  A thread owns
//   two adjacent strips of S = 8 cells along range.  A strip gets ONE threshold T_strip >= the
//   pre-filter threshold of each of its cells (below), so `ref < T_strip` is one binary value per
//   support cell, evaluated ONCE for the strip: (S + 2 HL) x (2 HC + 1) compares for 8 cells instead
//   of 8 x N_REF, and the two strips of a thread share their loads.  The masks are accumulated per
//   row pair; the window count of each cell (guard cells included: an upper bound) is a sliding sum
//   of row counts.  count(T_strip) >= count(T_cell), so a strip none of whose cells reaches rank_idx + 1
//   holds no detection.  In noise T_strip sits near 0.6 sigma: ~20 % of the references instead of the
//   ~76 % needed -- whole strips are rejected with 15 packed compares per cell.
// Phase 1 (surviving strips, compacted into a list so that warps stay dense) -- per-cell threshold:
//   the conservative test against the most permissive scale s1.  In adaptive mode the threshold is
//   tightened per cell before it is rounded up: a cell can only be detected with the scale the RTL
//   would pick, so
//     detection  =>  x_k < T = max(cut / SCALE_MAX, cut / SCALE_NOM, min(cut / SCALE_MIN, lo'))
//   where lo' >= lo = mean / 2 is an upper bound of the reference mean (the sum over the whole window,
//   columns widened to aligned groups of four, from a separable box filter over the staged tile).
//   SCALE_MIN is only chosen when x_k < lo (:197), i.e. next to strong cells.  All bounds carry a 1e-3
//   relative margin, far above float32 rounding and far below the bfloat16 step.
// Phase 2 (surviving cells) -- the EXACT decision in float32 (or the RTL's integer arithmetic) from
//   the raw tile: float64 sum -> mean -> scale 2/4/6 (:189-199) -> final count.
// The packed pre-filters assume magnitudes >= 0.  A tile that holds a negative word (a signed map
// handed to rdp_cfar_stage, or mag_mode = logpow below 0 dB) is decided cell by cell with the exact,
// sign-safe arithmetic instead -- slow, but identical to the oracle.
// K4 walks the bitmask of each CPI in (channel, range, doppler) order and writes the records, so
// the list is canonical without a sort and truncation on overflow is deterministic.
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
    int words_per_row;                      // bitmask words per range line = ceil(n_doppler / 32)
    int int_mode;                           // 1: the map holds unsigned integers and the decision is the bit-exact
                                            //    integer arithmetic of os_cfar_2d.vhd (DATA_WIDTH = 17)
    int is1, ismin, isnom, ismax;           // the scales as integers (int_mode)
    // pre-filter threshold: T = max(cut * pf_a, min(cut * pf_b, window_sum * pf_c))
    // (non-adaptive: pf_a = 1.001 / s1, pf_b = pf_c = 0)
    float pf_a, pf_b, pf_c;
    int neg1;                               // -1 (an opaque multiplier: see k3_count)
    int skew_div;                           // CTAs per residency slot (= SM count)
    int skew_ns;                            // RDP_K3_SKEW_NS: start-up skew between the CTAs that share an SM (see the kernel)
    int debug_skip;                         // RDP_K3_DEBUG_SKIP (timing experiments only; results are WRONG when set): 1 = no phases 1 / 2, 2 = no phase 0, 4 = no conversion
};

// magnitude word -> float for the pre-filter; integers below 2^24 convert exactly
template <bool INT_MODE>
__device__ __forceinline__ float cfar_word_as_float(float raw) {
    return INT_MODE ? (float)(uint32_t)__float_as_int(raw) : raw;
}
__device__ __forceinline__ float cfar_as_float(float raw, int int_mode) {
    return int_mode ? (float)(uint32_t)__float_as_int(raw) : raw;
}

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

// 0xFFFF in each half where a.half < b.half (bfloat16 compare)
__device__ __forceinline__ uint32_t bf16x2_lt_mask(uint32_t a, uint32_t b) {
#ifdef RDP_EMU
    uint32_t r = 0;
    for (int h = 0; h < 2; ++h) {
        float fa = __int_as_float((int)(((a >> (16 * h)) & 0xffffu) << 16));
        float fb = __int_as_float((int)(((b >> (16 * h)) & 0xffffu) << 16));
        if (fa < fb) r |= 0xffffu << (16 * h);
    }
    return r;
#else
    uint32_t r;
    asm("set.lt.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#endif
}
// acc -= mask on the FMA pipe (IMAD: mask * -1 + acc).  The packed compares and most of the bookkeeping around them
// run on the ALU pipe, which is the busier one in phases 0 and 1; RDP_K3_IMAD=0 at build time keeps plain subtractions.
#ifndef RDP_K3_IMAD
#define RDP_K3_IMAD 1
#endif
// (`neg1` is -1 handed in as a kernel argument: with a literal the assembler turns the multiply-add back into IADD3)
__device__ __forceinline__ void k3_count(uint32_t& acc, uint32_t mask, uint32_t neg1) {
#if defined(RDP_EMU) || !RDP_K3_IMAD
    (void)neg1;
    acc -= mask;
#else
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(mask), "r"(neg1));
#endif
}
// two floats (>= 0) -> their bfloat16 truncations, `a` in the low half and `b` in the high half
__device__ __forceinline__ uint32_t bf16_pack_down(float a, float b) {
#ifdef RDP_EMU
    return ((uint32_t)__float_as_int(a) >> 16) | ((uint32_t)__float_as_int(b) & 0xffff0000u);
#else
    return __byte_perm((uint32_t)__float_as_int(a), (uint32_t)__float_as_int(b), 0x7632);
#endif
}
// float (>= 0) -> bfloat16 bits rounded toward +inf
__device__ __forceinline__ uint32_t bf16_up(float v) { return ((uint32_t)__float_as_int(v) + 0xffffu) >> 16; }
// positive cuts so small that a threshold may fall into the bfloat16 subnormal range (pf_a >= 1e-6): one unsigned
// compare on the bit pattern covers 0 < cut < 1e-30; such cells / strips skip the packed tests
__device__ __forceinline__ bool cfar_tiny_cut(float cut) {
    return (uint32_t)__float_as_int(cut) - 1u < 0x0da24260u - 1u;
}

// Exact decision for one cell by one warp.  FETCH(dl, dc) returns the magnitude at that offset from the cell.
// Sign-safe: nothing here assumes magnitudes >= 0 (a CUT equal to zero is never a detection: the RTL's output
// word is the CUT itself and radar_core.vhd:415 keeps non-zero words only).
template <int MAXQ, class FETCH>
__device__ __forceinline__ bool cfar_exact_warp(const CfarArgs& a, FETCH fetch, float cut, int lane) {
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
        vals[q] = isref[q] ? fetch(dl, dc) : 0.0f;
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
    return c >= a.rank_idx + 1 && cut != 0.0f;                     // cut > x_k * scale  (:213), non-zero word
}

// ------------------------------------------------------------------------------------------
// Tiled kernel, geometry compile-time.  CTA = TR / 8 warps; tile = TR range lines x 64 Doppler
// columns; warp = one band of 8 lines, lane = two adjacent columns.
// ------------------------------------------------------------------------------------------
template <int HL, int HC, int GL, int GC, int TR_, int NSTAGE_, int P0B_ = 1>
struct K3Shape {
    static constexpr int S = 8, TC = 64, TR = TR_, NSTAGE = NSTAGE_;
    static constexpr int NBAND = TR / S, NWARP = NBAND, THREADS = 32 * NWARP;
    // phase 0: a thread owns P0B vertically adjacent bands (x 2 columns); with P0B = 2 only half of the warps run it,
    // but every 128-bit load of a column serves two strips more -- phase 0 is bound by shared-memory bandwidth
    static constexpr int P0B = P0B_;
    static_assert(NBAND % P0B_ == 0, "bands per phase-0 thread");
    // TMA needs the box to START on a 16-byte boundary in the innermost dimension (an unaligned start
    // coordinate traps with "illegal instruction", tools/microbench/mb_tma_probe.cu), so the column halo is
    // padded to a multiple of 4 floats on both sides: tile column c sits at staged column c + HCP
    static constexpr int HCP = (HC + 3) / 4 * 4;
    static constexpr int SC = TC + 2 * HCP;                       // staged columns
    static constexpr int SR = TR + 2 * HL;                        // staged rows (the TMA box)
    static constexpr int SRP = (SR + 7) / 8 * 8;                  // ... padded with rows that stay zero
    static constexpr int WIN = (2 * HL + 1) * (2 * HC + 1);
    static constexpr int MAXQ = (WIN + 31) / 32;
    static constexpr int TWORDS = TC / 32;                        // bitmask words per tile row
    static constexpr int NREF = WIN - (2 * GL + 1) * (2 * GC + 1);
    // packed planes: word (col, i) = bf16 of staged rows (2i, 2i+1) of column col; plane = col & 1
    static constexpr int NP = S / 2 + HL;                         // row pairs under one strip (its support)
    static constexpr int NPL = (NP + 3) / 4;                      // 128-bit loads per column
    static constexpr int PAIRS = SRP / 2;
    static constexpr int NP0 = NP + (S / 2) * (P0B - 1), NPL0 = (NP0 + 3) / 4;   // row pairs / 128-bit loads per column in phase 0
    static constexpr int LAST_READ = ((NBAND - 1) * (S / 2) + 4 * NPL > (NBAND - P0B) * (S / 2) + 4 * NPL0) ? (NBAND - 1) * (S / 2) + 4 * NPL : (NBAND - P0B) * (S / 2) + 4 * NPL0;
    static constexpr int PITCH_MIN = PAIRS > LAST_READ ? PAIRS : LAST_READ;
    // consecutive lanes read consecutive columns of one plane with 128-bit loads: the 8 lanes of a quarter-warp
    // cover all 32 banks iff the pitch is 4 mod 8 words
    static constexpr int PITCH = PITCH_MIN + ((12 - PITCH_MIN % 8) % 8);
    static constexpr int PLANE = (SC / 2) * PITCH;
    // conversion pass: one item = 8 staged rows x 4 staged columns
    static constexpr int NG = SC / 4, NQ8 = SRP / 8, NITEMS = NG * NQ8;
    // window-sum bound of the adaptive pre-filter: aligned column groups of 4, NCOV groups cover any window
    // (and the union of the windows of an even-aligned column pair)
    static constexpr int NCOV = (2 * HC + 7) / 4;
    // phase 2: G lanes per survivor (32 / G survivors per warp pass), NQ window slots per lane
    static constexpr int G = WIN <= 160 ? 8 : (WIN <= 320 ? 16 : 32);
    // (the slot table lists the reference cells only, padded to a multiple of G with slots that read a zero cell)
    static constexpr int NQ = (NREF + G - 1) / G, NSLOT = NQ * G, NZ = NSLOT - NREF;
    static constexpr int ZERO_IDX = SR * SC;
    static constexpr int TF_WORDS = (SR * SC + 4 + 31) / 32 * 32;     // one raw stage (+ a cell that reads 0.0f), 128 B multiple
    // phase 1 runs in rounds of STRIPS_PER_ROUND strips so that their survivors (<= 8 each) always fit the list
    static constexpr int STRIPS_PER_ROUND = THREADS, SURV_CAP = S * STRIPS_PER_ROUND;   // every thread of the CTA takes a strip
    static constexpr int TILE_BYTES = SR * SC * 4;
    static constexpr int N_STRIPS = NBAND * TC;
    static_assert(TR % S == 0 && TR * TC <= 65536 && (SC % 4) == 0 && (PITCH % 8) == 4 && PITCH % 4 == 0, "tile geometry");
    static_assert(4 * NCOV >= 2 * HC + 1 + 3 && 4 * NCOV >= 2 * HC + 2 + (HC % 2 ? 3 : 2), "NCOV groups cover the window at any alignment");
    static constexpr size_t smem_bytes() {
        return sizeof(float) * ((size_t)NSTAGE * TF_WORDS + SR * NG) + sizeof(uint32_t) * (2 * PLANE + TR * TWORDS) +
               sizeof(unsigned short) * (N_STRIPS + SURV_CAP) + sizeof(int) * (12 + 2 * NSLOT) + 2 * sizeof(unsigned long long) + 128;
    }
    static constexpr int occupancy_hint() {                        // CTAs per SM the shared memory allows (<= 2048 threads)
        int n = (int)((227 * 1024) / (smem_bytes() + 1024));
        if (n > 2048 / THREADS) n = 2048 / THREADS;
        if (P0B > 1 && n > 3) n = 3;                                  // the wider phase 0 wants ~80 registers
#ifdef RDP_K3_MAX_CTAS
        if (n > RDP_K3_MAX_CTAS) n = RDP_K3_MAX_CTAS;                 // build-time experiment: fewer CTAs per SM, more registers each
#endif
        return n < 1 ? 1 : n;
    }
};


// pf_c x the sums of the S vertical windows of a strip: wsum[j] = pf_c * sum_{k = j .. j + 2 HL} x[k].  Additions only
// (a sliding subtraction could cancel a huge line against the rest and fall BELOW the true sum).
template <int HL, int S>
__device__ __forceinline__ void k3_window_sums(const float (&x)[S + 2 * HL], float pf_c, float (&wsum)[S]) {
    if constexpr (2 * HL >= S - 1) {
        float common = x[S - 1];                                   // lines S-1 .. 2 HL belong to every window
        RDP_UNROLL
        for (int k = S; k <= 2 * HL; ++k) common += x[k];
        common *= pf_c;
        float part = 0.0f;
        wsum[S - 1] = common;
        RDP_UNROLL
        for (int j = S - 2; j >= 0; --j) { part += x[j]; wsum[j] = fmaf(part, pf_c, common); }
        part = 0.0f;
        RDP_UNROLL
        for (int j = 1; j < S; ++j) { part += x[2 * HL + j]; wsum[j] = fmaf(part, pf_c, wsum[j]); }
    } else {
        RDP_UNROLL
        for (int j = 0; j < S; ++j) {
            wsum[j] = 0.0f;
            RDP_UNROLL
            for (int k = 0; k <= 2 * HL; ++k) wsum[j] += x[j + k];
            wsum[j] *= pf_c;
        }
    }
}

// Packed counters: every packed compare yields 0xFFFF per true half and is SUBTRACTED from a 32-bit accumulator, so
// the low half counts its own hits and the high half holds (high hits - low hits) mod 2^16 (the borrow of the low
// half).  Sums of accumulators keep that form.
__device__ __forceinline__ uint32_t k3_lo_count(uint32_t p) { return p & 0xffffu; }
__device__ __forceinline__ uint32_t k3_hi_count(uint32_t p) { return ((p >> 16) + p) & 0xffffu; }

// Phase 0 verdict of one strip from its per-row-pair packed counters P[i] (low half: row 2i, high half: row 2i+1):
// does any of the 8 cells reach `need` references below the strip threshold?  Cells 2m and 2m+1 share the rows
// 2m .. 2m + 2 HL + 1 = pairs m .. m + HL; cell 2m lacks the last of them, cell 2m+1 the first.
template <int HL, int NP>
__device__ __forceinline__ bool k3_strip_reaches(const uint32_t (&P)[NP], int need) {
    uint32_t R = 0u;
    RDP_UNROLL
    for (int i = 0; i <= HL; ++i) R += P[i];
    bool f = false;
    RDP_UNROLL
    for (int m = 0; m < 4; ++m) {
        const uint32_t t = k3_lo_count(R) + k3_hi_count(R);
        const uint32_t hi_last = k3_hi_count(P[m + HL]), lo_first = k3_lo_count(P[m]);
        f = f || (int)(t - (hi_last < lo_first ? hi_last : lo_first)) >= need;
        if (m < 3) R = R - P[m] + P[m + HL + 1];
    }
    return f;
}

template <int HL, int HC, int GL, int GC, int TR, int NSTAGE, int P0B, bool INT_MODE>
__global__ void __launch_bounds__((K3Shape<HL, HC, GL, GC, TR, NSTAGE, P0B>::THREADS), (K3Shape<HL, HC, GL, GC, TR, NSTAGE, P0B>::occupancy_hint()))
k3_cfar_kernel(const float* __restrict__ rd_map, float* __restrict__ cfar_out, uint32_t* __restrict__ det_bits,
               CfarArgs a, int tiles_r, int tiles_c, long long n_tiles, int use_tma,
               const RDP_GRID_CONSTANT CUtensorMap tmap, uint32_t* __restrict__ cpi_flags, int* __restrict__ tile_queue) {
    // tile_queue (nullable): a counter zeroed before the launch.  The persistent CTAs take their first tile by index and
    // every later one from the queue (gridDim.x + atomicAdd), one tile ahead so that its TMA load can be issued early:
    // tiles differ several-fold in cost (crowded CPIs), and a fixed stride leaves the slowest CTA ~10 % behind.
    // cpi_flags (nullable, INT_MODE): word [map / n_channels] |= RDP_STATUS_MAP_RANGE when a tile holds a magnitude word
    // >= 2^24 (the packed pre-filter converts words to float and is only conservative while that is exact)
    using Sh = K3Shape<HL, HC, GL, GC, TR, NSTAGE, P0B>;
    constexpr int S = Sh::S, TC = Sh::TC, SC = Sh::SC, SR = Sh::SR, SRP = Sh::SRP, HCP = Sh::HCP, NG = Sh::NG;
    constexpr int NP = Sh::NP, NPL = Sh::NPL, PITCH = Sh::PITCH, PLANE = Sh::PLANE, THREADS = Sh::THREADS, NWARP = Sh::NWARP;
    RDP_DYN_SMEM(smem_raw);
    // 128-byte aligned carve-up: [raw stages][packed planes][cover sums][bitmask][counters][window table][strip list][survivors][mbarriers]
    unsigned char* smem_al = smem_raw + ((128 - ((uintptr_t)smem_raw & 127)) & 127);
    float* tf0 = reinterpret_cast<float*>(smem_al);                             // raw magnitudes
    uint32_t* planes = reinterpret_cast<uint32_t*>(tf0 + NSTAGE * Sh::TF_WORDS);  // [2][SC / 2][PITCH] bf16 row pairs
    float* cov = reinterpret_cast<float*>(planes + 2 * PLANE);                  // [SR][NG] sums of the 4 NCOV columns from group g on
    uint32_t* bits = reinterpret_cast<uint32_t*>(cov + SR * NG);                // [TR][TWORDS]
    int* ctr = reinterpret_cast<int*>(bits + TR * Sh::TWORDS);                  // 0: strips listed, 1: survivors listed, 2 / 3: tile holds a negative word (by iteration parity)
    int* wtab = ctr + 12;                                                       // [NSLOT][2]: tf index = ctr * mul + add  (ctr[8], ctr[9]: the next tile, by parity)
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(wtab + 2 * Sh::NSLOT);
    unsigned short* strips = reinterpret_cast<unsigned short*>(mbar + 2);       // band * 64 + column
    unsigned short* surv = strips + Sh::N_STRIPS;                               // tile line * 64 + column
    const int tiles_per_map = tiles_r * tiles_c;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int need = a.rank_idx + 1;
    const uint32_t neg1 = (uint32_t)a.neg1;

    // the zero cell behind every raw stage is never written again
    for (int st = 0; st < NSTAGE; ++st)
        for (int i = SR * SC + threadIdx.x; i < Sh::TF_WORDS; i += THREADS) tf0[st * Sh::TF_WORDS + i] = 0.0f;
    for (int i = threadIdx.x; i < TR * Sh::TWORDS; i += THREADS) bits[i] = 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) ctr[i] = 0;
#ifndef RDP_EMU
        if (use_tma) {
            mbar_init(mbar, 1);
            mbar_init(mbar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
#endif
    }
    for (int w = threadIdx.x; w < Sh::WIN + Sh::NZ; w += THREADS) {   // slot table: the reference cells in raster order, then padding
        if (w < Sh::WIN) {
            const int dl = w / (2 * HC + 1) - HL, dc = w % (2 * HC + 1) - HC;
            if (!(dl >= -GL && dl <= GL && dc >= -GC && dc <= GC)) {
                // guard cells in front of this one: whole guard rows above, part of its own row
                int gb = (dl > GL ? 2 * GL + 1 : (dl > -GL ? dl + GL : 0)) * (2 * GC + 1);
                if (dl >= -GL && dl <= GL) gb += dc > GC ? 2 * GC + 1 : 0;
                const int sl = w - gb;
                wtab[2 * sl] = 1;
                wtab[2 * sl + 1] = dl * SC + dc;
            }
        } else {
            const int sl = Sh::NREF + (w - Sh::WIN);
            wtab[2 * sl] = 0;
            wtab[2 * sl + 1] = Sh::ZERO_IDX;
        }
    }
    __syncthreads();
#ifndef RDP_EMU
    // TMA producer: the (tile + halo) box of tile `t` into raw stage `st`; coordinates may be negative /
    // past the edge -- the hardware zero-fills, which is exactly cfar_edges = ZERO
    const CUtensorMap* tmap_ptr = &tmap;                   // address of the __grid_constant__ parameter itself
    auto issue = [=](int t, int st) {                     // tile indices fit 32 bits (checked by the host)
        const int mp = t / tiles_per_map;
        const int rem = t - mp * tiles_per_map;
        mbar_expect_tx(mbar + st, (uint32_t)Sh::TILE_BYTES);
        tma_load_3d(tf0 + st * Sh::TF_WORDS, tmap_ptr, mbar + st, (rem % tiles_c) * TC - HCP, (rem / tiles_c) * TR - HL, (int)mp);
    };
    auto prefetch = [=](int t) {
        const int mp = t / tiles_per_map;
        const int rem = t - mp * tiles_per_map;
        tma_prefetch_3d(tmap_ptr, (rem % tiles_c) * TC - HCP, (rem / tiles_c) * TR - HL, (int)mp);
    };
    if (use_tma && threadIdx.x == 0 && (long long)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
#endif

#ifndef RDP_EMU
    // CTAs that share an SM start in lock-step and would run their issue-bound phase together and then their
    // latency-bound phases together; a one-off skew by residency slot lets one CTA's phase 0 fill another's gaps
    if (a.skew_ns > 0) {
        const unsigned slot = blockIdx.x / (unsigned)a.skew_div;
        for (unsigned i = 0; i < slot; ++i) __nanosleep((unsigned)a.skew_ns);
    }
#endif
    int iter = 0;
    const int n_tiles32 = (int)n_tiles;
    int next_tile = (int)blockIdx.x;
    for (int tile_id = blockIdx.x; tile_id < n_tiles32; tile_id = next_tile, ++iter) {
    // the tile after this one: thread 0 draws it now (before the first barrier of the iteration) and publishes it in
    // ctr[8 + parity]; everybody reads it after the last barrier
    if (tile_queue != nullptr) {
        if (threadIdx.x == 0) { next_tile = (int)gridDim.x + atomicAdd(tile_queue, 1); ctr[8 + (iter & 1)] = next_tile; }
    } else {
        next_tile = tile_id + (int)gridDim.x;
    }
    const int stage = (use_tma && NSTAGE == 2) ? (iter & 1) : 0;
    float* tf = tf0 + stage * Sh::TF_WORDS;
    const long long map = tile_id / tiles_per_map;
    const int in_map = tile_id - (int)map * tiles_per_map;
    const int tr = in_map / tiles_c, tc = in_map - tr * tiles_c;
    const int r0 = tr * TR, c0 = tc * TC;
    const float* m = rd_map + map * (long long)a.n_range * a.n_doppler;
    const long long out_base = map * (long long)a.n_range * a.n_doppler;

#ifndef RDP_EMU
    if (use_tma) {
        // two stages: the next tile is fetched into the other stage (its readers finished before the barrier that
        // ended the previous iteration) while this one is processed; one stage: it was issued at the end of the
        // previous iteration and other CTAs of the SM cover the latency
        if (NSTAGE == 2 && threadIdx.x == 0 && next_tile < n_tiles32) issue(next_tile, stage ^ 1);
        // one stage: the next tile cannot land in shared memory yet, but it can be on its way into L2
        if (NSTAGE == 1 && (a.debug_skip & 8) == 0 && tile_queue == nullptr && threadIdx.x == 32 && next_tile < n_tiles32) prefetch(next_tile);
        mbar_wait(mbar + stage, (uint32_t)((NSTAGE == 2 ? (iter >> 1) : iter) & 1));
    } else
#endif
    {   // tile + halo, one warp per staged row (coalesced, no per-element index arithmetic)
        for (int lr = warp; lr < SR; lr += NWARP) {
            const int line = r0 - HL + lr;
            const bool row_ok = a.raster || (line >= 0 && line < a.n_range);
            const long long row_base = (long long)line * a.n_doppler;
            RDP_UNROLL
            for (int k = 0; k < (SC + 31) / 32; ++k) {
                const int lc = lane + 32 * k;
                if (lc < SC) {
                    const int c = c0 - HCP + lc;
                    float v = 0.0f;
                    if (a.raster) {
                        const long long flat = row_base + c;
                        if (flat >= 0 && flat < (long long)a.n_range * a.n_doppler) v = __ldg(m + flat);
                    } else if (row_ok && c >= 0 && c < a.n_doppler) {
                        v = __ldg(m + row_base + c);
                    }
                    tf[lr * SC + lc] = v;
                }
            }
        }
        __syncthreads();
    }

    // ---- conversion: raw words -> packed bf16 row pairs (two planes), column-group cover sums, sign check --------
    if (!(a.debug_skip & 4)) {
        constexpr int OV = Sh::NCOV - 1, ADV = 32 - OV;
        static_assert(ADV >= 8, "cover shuffles");
        const float4* tf4 = reinterpret_cast<const float4*>(tf);
        uint32_t neg = 0u;
        for (int base = warp * ADV; base < Sh::NITEMS; base += NWARP * ADV) {
            const int it = base + lane;
            const bool ok = it < Sh::NITEMS;
            const int q = ok ? it / NG : 0, g = ok ? it - q * NG : 0;
            float gs[8];
            uint32_t wd[4][4];                                      // [column of the group][row pair]
            RDP_UNROLL
            for (int i = 0; i < 4; ++i) {
                // (SR is even: a row pair is inside the staged tile or past its end, where the planes read as zero)
                const bool in_tile = SR % 8 == 0 || 8 * q + 2 * i < SR;
                const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                const float4 ra = in_tile ? tf4[((8 * q + 2 * i) * SC + 4 * g) >> 2] : zero4;
                const float4 rb = in_tile ? tf4[((8 * q + 2 * i + 1) * SC + 4 * g) >> 2] : zero4;
                {   // float maps: the sign bits; integer maps: the bits above 2^24
                    neg |= (uint32_t)__float_as_int(ra.x) | (uint32_t)__float_as_int(ra.y) | (uint32_t)__float_as_int(ra.z) | (uint32_t)__float_as_int(ra.w);
                    neg |= (uint32_t)__float_as_int(rb.x) | (uint32_t)__float_as_int(rb.y) | (uint32_t)__float_as_int(rb.z) | (uint32_t)__float_as_int(rb.w);
                }
                const float ax = cfar_word_as_float<INT_MODE>(ra.x), ay = cfar_word_as_float<INT_MODE>(ra.y);
                const float az = cfar_word_as_float<INT_MODE>(ra.z), aw = cfar_word_as_float<INT_MODE>(ra.w);
                const float bx = cfar_word_as_float<INT_MODE>(rb.x), by = cfar_word_as_float<INT_MODE>(rb.y);
                const float bz = cfar_word_as_float<INT_MODE>(rb.z), bw = cfar_word_as_float<INT_MODE>(rb.w);
                wd[0][i] = bf16_pack_down(ax, bx);
                wd[1][i] = bf16_pack_down(ay, by);
                wd[2][i] = bf16_pack_down(az, bz);
                wd[3][i] = bf16_pack_down(aw, bw);
                gs[2 * i] = (ax + ay) + (az + aw);
                gs[2 * i + 1] = (bx + by) + (bz + bw);
            }
            const bool mine = ok && lane < ADV;
            if (mine) {
                RDP_UNROLL
                for (int c = 0; c < 4; ++c) {
                    uint4 o;
                    o.x = wd[c][0]; o.y = wd[c][1]; o.z = wd[c][2]; o.w = wd[c][3];
                    *reinterpret_cast<uint4*>(planes + (c & 1) * PLANE + (2 * g + (c >> 1)) * PITCH + 4 * q) = o;
                }
            }
            if (a.adaptive) {
                // cov[row][g] = sum of the 4 NCOV columns from group g on (groups past the end of the row count as 0);
                // each warp steps by 32 - (NCOV - 1) items so that the neighbours it needs sit in its own lanes
                RDP_UNROLL
                for (int k = 0; k < 8; ++k) {
                    float cover = gs[k];
                    RDP_UNROLL
                    for (int d = 1; d <= OV; ++d) {
                        const float nb = __shfl_down_sync(0xffffffffu, gs[k], d);
                        cover += g + d < NG ? nb : 0.0f;
                    }
                    if (mine && (SR % 8 == 0 || 8 * q + k < SR)) cov[(8 * q + k) * NG + g] = cover;
                }
            }
        }
        if (!INT_MODE) {
            if (__any_sync(0xffffffffu, (int)(neg >> 31)) && lane == 0) ctr[4 + (iter & 1)] = 1;
        } else if (cpi_flags != nullptr) {
            if (__any_sync(0xffffffffu, (neg >> 24) != 0u) && lane == 0) atomicOr(cpi_flags + map / a.n_channels, (uint32_t)RDP_STATUS_MAP_RANGE);
        }
    }
    __syncthreads();
    // counters alternate by tile parity: this tile's are reset by thread 0 while the NEXT tile is in flight (after
    // that tile's first barrier, when nobody reads them any more and nobody has written the ones after it yet)
    const int par = iter & 1;
    int* const n_strips_ctr = ctr + par;                           // strips listed by phase 0
    int* const n_surv_ctr = ctr + 2 + par;                         // survivors listed by phase 1, running over the rounds
    if (threadIdx.x == 0) { ctr[par ^ 1] = 0; ctr[2 + (par ^ 1)] = 0; ctr[4 + (par ^ 1)] = 0; }
    const bool signed_tile = !INT_MODE && ctr[4 + par] != 0;

    if (!signed_tile) {
    // ---- phase 0: one shared threshold per strip ------------------------------------------------------------
    if (!(a.debug_skip & 2) && warp < NWARP / P0B) {
        constexpr int NP0 = Sh::NP0, NPL0 = Sh::NPL0;
        const int b0 = P0B * warp;                                 // bands b0 .. b0 + P0B - 1: tile lines 8 b0 .., staged rows 8 b0 .. 8 (b0 + P0B) - 1 + 2 HL
        const int s0 = 2 * lane + HCP;                             // staged column of the first of this lane's two columns
        uint32_t td[P0B][2];                                       // strip thresholds (band, column), duplicated into both halves
        bool tiny[P0B][2];
        RDP_UNROLL
        for (int bb = 0; bb < P0B; ++bb) {
            float cm0 = 0.0f, cm1 = 0.0f;                          // largest cut of either strip
            RDP_UNROLL
            for (int j = 0; j < S; ++j) {
                const float2 c2 = *reinterpret_cast<const float2*>(tf + (S * (b0 + bb) + HL + j) * SC + s0);
                cm0 = fmaxf(cm0, cfar_word_as_float<INT_MODE>(c2.x));
                cm1 = fmaxf(cm1, cfar_word_as_float<INT_MODE>(c2.y));
            }
            float lo_b = 0.0f;                                     // pf_c x the largest window-sum bound of the strip's cells
            if (a.adaptive) {
                const float* gp = cov + (S * (b0 + bb)) * NG + ((s0 - HC) >> 2);
                float x[S + 2 * HL];
                RDP_UNROLL
                for (int k = 0; k < S + 2 * HL; ++k) x[k] = gp[k * NG];
                float wsum[S];
                k3_window_sums<HL, S>(x, a.pf_c, wsum);
                lo_b = wsum[0];
                RDP_UNROLL
                for (int j = 1; j < S; ++j) lo_b = fmaxf(lo_b, wsum[j]);
            }
            const uint32_t h0 = bf16_up(fmaxf(cm0 * a.pf_a, fminf(cm0 * a.pf_b, lo_b)));
            const uint32_t h1 = bf16_up(fmaxf(cm1 * a.pf_a, fminf(cm1 * a.pf_b, lo_b)));
            td[bb][0] = h0 | (h0 << 16);
            td[bb][1] = h1 | (h1 << 16);
            tiny[bb][0] = cfar_tiny_cut(cm0);
            tiny[bb][1] = cfar_tiny_cut(cm1);
        }
        uint32_t acc[P0B][2][NP];
        RDP_UNROLL
        for (int bb = 0; bb < P0B; ++bb)
            RDP_UNROLL
            for (int i = 0; i < NP; ++i) { acc[bb][0][i] = 0u; acc[bb][1][i] = 0u; }
        const uint32_t* pl = planes + (S / 2) * b0 + lane * PITCH;
        // one window column: staged column s0 - HC + k; LEFT / RIGHT = which of the thread's two strips it belongs to
        auto p0_column = [&](const uint32_t* colp, bool left, bool right) {
            const uint4* src = reinterpret_cast<const uint4*>(colp);
            uint32_t w[4 * NPL0];
            RDP_UNROLL
            for (int v = 0; v < NPL0; ++v) {
                const uint4 u = src[v];
                w[4 * v] = u.x; w[4 * v + 1] = u.y; w[4 * v + 2] = u.z; w[4 * v + 3] = u.w;
            }
            RDP_UNROLL
            for (int bb = 0; bb < P0B; ++bb) {                     // band bb sees the pairs (S / 2) bb .. (S / 2) bb + NP - 1 of this load
                if (left) {
                    RDP_UNROLL
                    for (int i = 0; i < NP; ++i) k3_count(acc[bb][0][i], bf16x2_lt_mask(w[(S / 2) * bb + i], td[bb][0]), neg1);
                }
                if (right) {
                    RDP_UNROLL
                    for (int i = 0; i < NP; ++i) k3_count(acc[bb][1][i], bf16x2_lt_mask(w[(S / 2) * bb + i], td[bb][1]), neg1);
                }
            }
        };
        // column k sits in plane (HC + k) & 1 at position (HCP - HC + k) >> 1.  The first column belongs to the left strip
        // only, the last to the right one; the 2 HC columns between them go in HC steps of two (one per plane).
        // RDP_ROLL_P0 / _P1 keep these loops rolled or unrolled (rdp_common.cuh: measured, unrolled is the default)
        p0_column(pl + (HC & 1) * PLANE + ((HCP - HC) >> 1) * PITCH, true, false);
        {
            const uint32_t* ca = pl + ((HC + 1) & 1) * PLANE + ((HCP - HC + 1) >> 1) * PITCH;     // k = 1, 3, ...
            const uint32_t* cb = pl + (HC & 1) * PLANE + ((HCP - HC + 2) >> 1) * PITCH;           // k = 2, 4, ...
            RDP_ROLL_P0
            for (int kp = 0; kp < HC; ++kp) {
                p0_column(ca, true, true);
                p0_column(cb, true, true);
                ca += PITCH;
                cb += PITCH;
            }
        }
        p0_column(pl + ((HC + 2 * HC + 1) & 1) * PLANE + ((HCP - HC + 2 * HC + 1) >> 1) * PITCH, false, true);
        static_assert(NP0 <= 4 * NPL0, "phase-0 loads");
        RDP_UNROLL
        for (int bb = 0; bb < P0B; ++bb) {
            const bool f0 = k3_strip_reaches<HL, NP>(acc[bb][0], need) || tiny[bb][0];
            const bool f1 = k3_strip_reaches<HL, NP>(acc[bb][1], need) || tiny[bb][1];
            const uint32_t m0 = __ballot_sync(0xffffffffu, f0), m1 = __ballot_sync(0xffffffffu, f1);
            const int n0 = __popc(m0), n1 = __popc(m1);
            if (n0 + n1) {
                int base = 0;
                if (lane == 0) base = atomicAdd(n_strips_ctr, n0 + n1);
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t below = (1u << lane) - 1u;
                if (f0) strips[base + __popc(m0 & below)] = (unsigned short)((b0 + bb) * TC + 2 * lane);
                if (f1) strips[base + n0 + __popc(m1 & below)] = (unsigned short)((b0 + bb) * TC + 2 * lane + 1);
            }
        }
        if (cfar_out && c0 + 2 * lane < a.n_doppler) {             // dense output: zero everywhere, phase 2 writes the detections
            RDP_UNROLL
            for (int j = 0; j < P0B * S; ++j)
                if (r0 + S * b0 + j < a.n_range)
                    *reinterpret_cast<float2*>(cfar_out + out_base + (long long)(r0 + S * b0 + j) * a.n_doppler + c0 + 2 * lane) = make_float2(0.0f, 0.0f);
        }
    }
    __syncthreads();

    // ---- phases 1 and 2 in rounds of STRIPS_PER_ROUND strips (one round unless the tile is crowded) ---------------
    const int n_strips = (a.debug_skip & 1) ? 0 : *n_strips_ctr;
    int surv_begin = 0;
    for (int sbase = 0; sbase < n_strips; sbase += Sh::STRIPS_PER_ROUND) {
    // ---- phase 1: surviving strips, per-cell thresholds -----------------------------------------------------
    if (n_strips <= THREADS / S) {
        // few strips (the usual case away from targets): S lanes per strip, one cell each, so that the whole test is
        // ~300 instructions deep instead of ~1100.  The window membership of a row pair depends on the lane's cell:
        // it is folded into per-pair threshold operands (0 in a half that is outside the window / inside the guard box:
        // no magnitude is below 0), and all NP pairs of every column are compared.
        const int sidx = (int)threadIdx.x / S, j = (int)threadIdx.x % S;
        if (sidx < n_strips) {
            const int code = strips[sidx];
            const int b = code / TC, tx = code % TC;
            const int s = tx + HCP;
            const float cut = cfar_word_as_float<INT_MODE>(tf[(S * b + HL + j) * SC + s]);
            float t = cut * a.pf_a;
            if (a.adaptive) {
                const float* gp = cov + (S * b + j) * NG + ((s - HC) >> 2);
                float w = 0.0f;
                RDP_UNROLL
                for (int k = 0; k <= 2 * HL; ++k) w += gp[k * NG];
                t = fmaxf(t, fminf(cut * a.pf_b, w * a.pf_c));
            }
            const uint32_t h = bf16_up(t);
            uint32_t opn[NP], opg[NP];
            RDP_UNROLL
            for (int i = 0; i < NP; ++i) {
                const int ra = 2 * i, rb = 2 * i + 1;
                const bool in_a = ra >= j && ra <= j + 2 * HL, in_b = rb >= j && rb <= j + 2 * HL;
                const bool g_a = ra >= j + HL - GL && ra <= j + HL + GL, g_b = rb >= j + HL - GL && rb <= j + HL + GL;
                opn[i] = (in_a ? h : 0u) | (in_b ? h << 16 : 0u);
                opg[i] = (in_a && !g_a ? h : 0u) | (in_b && !g_b ? h << 16 : 0u);
            }
            const int par0 = (s - HC) & 1, pos0 = (s - HC) >> 1;
            const uint32_t* pa = planes + par0 * PLANE + pos0 * PITCH + (S / 2) * b;
            const uint32_t* pb = planes + (par0 ^ 1) * PLANE + (pos0 + par0) * PITCH + (S / 2) * b;
            uint32_t acc = 0u;
            auto p1s_column = [&](const uint32_t* colp, bool guard_col) {
                const uint4* src = reinterpret_cast<const uint4*>(colp);
                uint32_t w[4 * NPL];
                RDP_UNROLL
                for (int v = 0; v < NPL; ++v) {
                    const uint4 u = src[v];
                    w[4 * v] = u.x; w[4 * v + 1] = u.y; w[4 * v + 2] = u.z; w[4 * v + 3] = u.w;
                }
                RDP_UNROLL
                for (int i = 0; i < NP; ++i) k3_count(acc, bf16x2_lt_mask(w[i], guard_col ? opg[i] : opn[i]), neg1);
            };
            // window column kk: plane pa / pb by parity, position kk / 2; rolled, one body per column kind
            RDP_ROLL_P1
            for (int kk = 0; kk <= 2 * HC; ++kk) {
                const uint32_t* colp = ((kk & 1) ? pb : pa) + (kk >> 1) * PITCH;
                if (kk >= HC - GC && kk <= HC + GC) p1s_column(colp, true);
                else p1s_column(colp, false);
            }
            const bool keep = ((int)(k3_lo_count(acc) + k3_hi_count(acc)) >= need || cfar_tiny_cut(cut)) &&
                              c0 + tx < a.n_doppler && r0 + S * b + j < a.n_range;
            if (keep) surv[atomicAdd(n_surv_ctr, 1) - surv_begin] = (unsigned short)((S * b + j) * TC + tx);
        }
    } else {
        const int idx = sbase + (int)threadIdx.x;
        if (threadIdx.x < Sh::STRIPS_PER_ROUND && idx < n_strips) {
            const int code = strips[idx];
            const int b = code / TC, tx = code % TC;
            const int s = tx + HCP;                                 // staged column of the strip
            float wsum[S];
            if (a.adaptive) {
                // vertical pass of the box filter in registers: wsum[j] = pf_c x sum_{k = j .. j + 2 HL} x[k], additions only
                const float* gp = cov + (S * b) * NG + ((s - HC) >> 2);
                float x[S + 2 * HL];
                RDP_UNROLL
                for (int k = 0; k < S + 2 * HL; ++k) x[k] = gp[k * NG];
                k3_window_sums<HL, S>(x, a.pf_c, wsum);
            }
            uint32_t td[S], tl[S], th[S], acc[S];
            uint32_t fmask = 0u;                                   // bit j: cell j goes to the exact phase regardless
            RDP_UNROLL
            for (int j = 0; j < S; ++j) {
                const float cut = cfar_word_as_float<INT_MODE>(tf[(S * b + HL + j) * SC + s]);
                float t = cut * a.pf_a;
                if (a.adaptive) t = fmaxf(t, fminf(cut * a.pf_b, wsum[j]));
                const uint32_t h = bf16_up(t);
                tl[j] = h; th[j] = h << 16; td[j] = h | (h << 16);
                acc[j] = 0u;
                fmask |= cfar_tiny_cut(cut) ? (1u << j) : 0u;
            }
            const int par0 = (s - HC) & 1, pos0 = (s - HC) >> 1;
            const uint32_t* pa = planes + par0 * PLANE + pos0 * PITCH + (S / 2) * b;            // columns s - HC + even
            const uint32_t* pb = planes + (par0 ^ 1) * PLANE + (pos0 + par0) * PITCH + (S / 2) * b;   // columns s - HC + odd
            auto p1d_column = [&](const uint32_t* colp, bool guard_col) {
                const uint4* src = reinterpret_cast<const uint4*>(colp);
                uint32_t w[4 * NPL];
                RDP_UNROLL
                for (int v = 0; v < NPL; ++v) {
                    const uint4 u = src[v];
                    w[4 * v] = u.x; w[4 * v + 1] = u.y; w[4 * v + 2] = u.z; w[4 * v + 3] = u.w;
                }
                RDP_UNROLL
                for (int j = 0; j < S; ++j) {
                    RDP_UNROLL
                    for (int i = 0; i < NP; ++i) {
                        // rows 2i, 2i+1 of the support against cell j: window rows j .. j + 2 HL, guard rows j + HL - GL .. j + HL + GL
                        const int ra = 2 * i, rb = 2 * i + 1;
                        const bool in_a = ra >= j && ra <= j + 2 * HL && !(guard_col && ra >= j + HL - GL && ra <= j + HL + GL);
                        const bool in_b = rb >= j && rb <= j + 2 * HL && !(guard_col && rb >= j + HL - GL && rb <= j + HL + GL);
                        if (in_a && in_b) k3_count(acc[j], bf16x2_lt_mask(w[i], td[j]), neg1);
                        else if (in_a) k3_count(acc[j], bf16x2_lt_mask(w[i], tl[j]), neg1);
                        else if (in_b) k3_count(acc[j], bf16x2_lt_mask(w[i], th[j]), neg1);
                    }
                }
            };
            RDP_ROLL_P1
            for (int kk = 0; kk <= 2 * HC; ++kk) {                    // rolled: one body for guard columns, one for the others
                const uint32_t* colp = ((kk & 1) ? pb : pa) + (kk >> 1) * PITCH;
                if (kk >= HC - GC && kk <= HC + GC) p1d_column(colp, true);
                else p1d_column(colp, false);
            }
            uint32_t smask = fmask;
            RDP_UNROLL
            for (int j = 0; j < S; ++j) smask |= ((int)(k3_lo_count(acc[j]) + k3_hi_count(acc[j])) >= need) ? (1u << j) : 0u;
            const int lines_left = a.n_range - (r0 + S * b);            // strips / columns past the edge of the map
            if (c0 + tx >= a.n_doppler || lines_left <= 0) smask = 0u;
            else if (lines_left < S) smask &= (1u << lines_left) - 1u;
            if (smask) {
                int slot = atomicAdd(n_surv_ctr, __popc(smask)) - surv_begin;
                while (smask) {
                    const int j = __ffs(smask) - 1;
                    smask &= smask - 1u;
                    surv[slot++] = (unsigned short)((S * b + j) * TC + tx);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: exact decision from the raw tile: G lanes per survivor, 32/G survivors per warp pass ---
    // Window slots are dealt to the G lanes (slot = sub + G q); guard / padding slots read the zero cell and
    // their known contribution to each count is subtracted.  Counting = sign bit of an exact float subtraction.
    {
    const int surv_end = *n_surv_ctr;
    const int ns = surv_end - surv_begin;
    surv_begin = surv_end;
    constexpr int G = Sh::G, NQ = Sh::NQ, NZ = Sh::NZ;
    const int sub = lane % G, grp = lane / G;
    const bool n_ref_pow2 = (a.n_ref & (a.n_ref - 1)) == 0;
    const double inv_n = 1.0 / (double)a.n_ref;
    // this lane's NQ window slots as word offsets from the cell (padding slots: the zero cell, addressed absolutely):
    // fixed for the lane, so they are read once per round instead of once per survivor
    int soff[NQ];
    if (warp * (32 / G) < ns) {
        RDP_UNROLL
        for (int q = 0; q < NQ; ++q) soff[q] = wtab[2 * (q * G + sub) + 1];
    }
    constexpr int NQ_REAL = Sh::NREF / G;                          // slots q < NQ_REAL are reference cells in every lane
    for (int i0 = warp * (32 / G); i0 < ns; i0 += NWARP * (32 / G)) {
        const int i = i0 + grp;
        const bool valid = i < ns;
        const int code = valid ? (int)surv[i] : 0;                 // idle groups chew on tile cell (0, 0)
        const int tl = code / TC, tcx = code % TC;
        const int cidx = (tl + HL) * SC + tcx + HCP;
        bool det;
        if (!INT_MODE) {
        const float cut = tf[cidx];
        float v[NQ];
        const float* tfc = tf + cidx;
        RDP_UNROLL
        for (int q = 0; q < NQ; ++q) {
            if (q < NQ_REAL) v[q] = tfc[soff[q]];
            else v[q] = (q * G + sub < Sh::NREF) ? tfc[soff[q]] : tf[soff[q]];
        }
        float scale = a.s1;
        if (a.adaptive) {
            double s = 0.0, s_b = 0.0;                             // two chains: half the add latency (the sums are exact, the order is free, for any realistic spread of exponents)
            RDP_UNROLL
            for (int q = 0; q + 1 < NQ; q += 2) { s += (double)v[q]; s_b += (double)v[q + 1]; }
            if (NQ & 1) s += (double)v[NQ - 1];
            s += s_b;
            RDP_UNROLL
            for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            // mean = float32(S / N_REF)  (os_cfar_2d.vhd:189); a power-of-two N_REF divides exactly by multiplication
            const float mean = n_ref_pow2 ? (float)(s * inv_n) : (float)(s / (double)a.n_ref);
            const float lo = __fmul_rn(0.5f, mean);
            const float hi = __fadd_rn(mean, lo);
            uint32_t c_gt_hi = 0u, c_lt_lo = 0u;
            RDP_UNROLL
            for (int q = 0; q < NQ; ++q) {
                c_gt_hi += (uint32_t)__float_as_int(__fadd_rn(hi, -v[q])) >> 31;      // v > hi
                c_lt_lo += (uint32_t)__float_as_int(__fadd_rn(v[q], -lo)) >> 31;      // v < lo (zero slots count iff lo > 0)
            }
            int packed = (int)(c_gt_hi | (c_lt_lo << 16));
            RDP_UNROLL
            for (int o = G / 2; o > 0; o >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, o);
            const int n_le_hi = a.n_ref - (packed & 0xffff);
            const int n_lt_lo = (packed >> 16) - (lo > 0.0f ? NZ : 0);
            // :193-199  x_k > hi -> MAX ; x_k < lo -> MIN ; else NOM
            scale = (n_le_hi <= a.rank_idx) ? a.smax : (n_lt_lo >= a.rank_idx + 1) ? a.smin : a.snom;
        }
        int c = 0;
        RDP_UNROLL
        for (int q = 0; q < NQ; ++q) c += (int)((uint32_t)__float_as_int(__fadd_rn(__fmul_rn(v[q], scale), -cut)) >> 31);
        RDP_UNROLL
        for (int o = G / 2; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        const int n_lt = c - NZ;                                   // zero slots: 0 - cut < 0 (cut > 0 here)
        det = valid && cut > 0.0f && n_lt >= a.rank_idx + 1;       // cut > x_k * scale  (:213)
        } else {
        // bit-exact integer arithmetic of os_cfar_2d.vhd (17-bit magnitudes; values < 2^24 supported)
        const uint32_t* ti = reinterpret_cast<const uint32_t*>(tf);
        const int cut = (int)ti[cidx];
        int v[NQ];
        const uint32_t* tic = ti + cidx;
        RDP_UNROLL
        for (int q = 0; q < NQ; ++q) {
            if (q < NQ_REAL) v[q] = (int)tic[soff[q]];
            else v[q] = (int)((q * G + sub < Sh::NREF) ? tic[soff[q]] : ti[soff[q]]);
        }
        int scale = a.is1;
        if (a.adaptive) {
            unsigned long long s = 0ull;                          // sum_refs (:163), 25 bits in the RTL
            RDP_UNROLL
            for (int q = 0; q < NQ; ++q) s += (unsigned long long)(uint32_t)v[q];
            RDP_UNROLL
            for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            // mean_val = resize(sum_refs / N_REF, 17) (:189): floor division (exact through float64: the quotient
            // of integers below 2^53 is either an integer or at least 1/N_REF away from one), truncated to 17 bits
            const int mean = (int)((unsigned long long)((double)s / (double)a.n_ref) & 0x1ffffull);
            const int lo = mean >> 1;                             // shift_right(mean_val, 1)
            const int hi = (mean + lo) & 0x1ffff;                 // 17-bit unsigned add (:193)
            uint32_t c_gt_hi = 0u, c_lt_lo = 0u;
            RDP_UNROLL
            for (int q = 0; q < NQ; ++q) {
                c_gt_hi += (uint32_t)(hi - v[q]) >> 31;           // v > hi
                c_lt_lo += (uint32_t)(v[q] - lo) >> 31;           // v < lo (zero slots count iff lo > 0)
            }
            int packed = (int)(c_gt_hi | (c_lt_lo << 16));
            RDP_UNROLL
            for (int o = G / 2; o > 0; o >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, o);
            const int n_le_hi = a.n_ref - (packed & 0xffff);
            const int n_lt_lo = (packed >> 16) - (lo > 0 ? NZ : 0);
            scale = (n_le_hi <= a.rank_idx) ? a.ismax : (n_lt_lo >= a.rank_idx + 1) ? a.ismin : a.isnom;   // :191-199
        }
        int c = 0;
        RDP_UNROLL
        for (int q = 0; q < NQ; ++q) c += (int)((uint32_t)(v[q] * scale - cut) >> 31);     // ref * scale < cut (:204,213)
        RDP_UNROLL
        for (int o = G / 2; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        det = valid && cut > 0 && (c - NZ) >= a.rank_idx + 1;
        }
        if (det && sub == 0) {
            atomicOr(bits + tl * Sh::TWORDS + (tcx >> 5), 1u << (tcx & 31));
            if (cfar_out) cfar_out[out_base + (long long)(r0 + tl) * a.n_doppler + c0 + tcx] = tf[cidx];   // raw word, unchanged in either mode
        }
    }
    }
    __syncthreads();                                               // the survivor list is free again; bits of this round are in
    }   // rounds
    } else {
    // ---- a tile with negative words: every cell decided exactly by one warp (sign-safe, slow) ---------------
    for (int cell = warp; cell < TR * TC; cell += NWARP) {
        const int tl = cell / TC, tcx = cell % TC;
        if (r0 + tl >= a.n_range || c0 + tcx >= a.n_doppler) continue;
        const float* ctr_cell = tf + (tl + HL) * SC + tcx + HCP;
        const float cut = *ctr_cell;
        const bool det = cfar_exact_warp<Sh::MAXQ>(a, [&](int dl, int dc) { return ctr_cell[dl * SC + dc]; }, cut, lane);
        if (lane == 0) {
            if (det) atomicOr(bits + tl * Sh::TWORDS + (tcx >> 5), 1u << (tcx & 31));
            if (cfar_out) cfar_out[out_base + (long long)(r0 + tl) * a.n_doppler + c0 + tcx] = det ? cut : 0.0f;
        }
    }
    __syncthreads();
    }
    // every read of this tile's raw stage, planes and lists lies before a barrier every thread has passed
    if (tile_queue != nullptr) next_tile = ctr[8 + (iter & 1)];   // published before this iteration's first barrier
#ifndef RDP_EMU
    if (use_tma && NSTAGE == 1 && threadIdx.x == 0 && next_tile < n_tiles32) issue(next_tile, 0);
#endif
    // bitmask of this tile: rows r0.., words (c0 / 32)..  (TC is a multiple of 32 so tiles own whole words)
    uint32_t* bm = det_bits + map * (long long)a.n_range * a.words_per_row;
    for (int i = threadIdx.x; i < TR * Sh::TWORDS; i += THREADS) {
        const int rl = i / Sh::TWORDS, w = i % Sh::TWORDS;
        const int line = r0 + rl, word = c0 / 32 + w;
        if (line < a.n_range && word < a.words_per_row) bm[(long long)line * a.words_per_row + word] = bits[i];
        bits[i] = 0u;
    }
    }   // tile loop
}

// ------------------------------------------------------------------------------------------
// Generic kernel, geometry at run time (any window up to 32*MAXQ_GENERIC cells): one warp per
// 32 consecutive cells (one bitmask word) straight from global / L2.  Slow; used only for window
// shapes without a tiled instance.
// ------------------------------------------------------------------------------------------
constexpr int MAXQ_GENERIC = 40;   // windows up to 1280 cells

__global__ void __launch_bounds__(256)
k3_cfar_generic_kernel(const float* __restrict__ rd_map, float* __restrict__ cfar_out,
                       uint32_t* __restrict__ det_bits, CfarArgs a, long long n_words) {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const long long words_per_map = (long long)a.n_range * a.words_per_row;
    const long long cells_per_map = (long long)a.n_range * a.n_doppler;
    for (long long w = (long long)blockIdx.x * 8 + warp; w < n_words; w += (long long)gridDim.x * 8) {
        const long long map = w / words_per_map;
        const int line = (int)((w % words_per_map) / a.words_per_row), word = (int)(w % a.words_per_row);
        const float* m = rd_map + map * cells_per_map;
        uint32_t out = 0u;
        for (int b = 0; b < 32; ++b) {
            const int col = word * 32 + b;
            if (col >= a.n_doppler) break;
            const float cv = __ldg(m + (long long)line * a.n_doppler + col);
            const bool det = cfar_exact_warp<MAXQ_GENERIC>(a, [&](int dl, int dc) { return cfar_fetch(m, line + dl, col + dc, a.n_range, a.n_doppler, a.raster); }, cv, lane);
            if (det) out |= 1u << b;
            if (lane == 0 && cfar_out) cfar_out[map * cells_per_map + (long long)line * a.n_doppler + col] = det ? cv : 0.0f;
        }
        if (lane == 0) det_bits[w] = out;
    }
}

// ------------------------------------------------------------------------------------------
// K3b (optional, rdp_config.peak_group; not in the reference -- SURVEY 8f row 3 "detection clustering / peak
// grouping"): a detection survives only if no detection among its 8 neighbours (same map, inside the map) has a
// larger magnitude -- ties go to the cell that comes first in (range, doppler) raster order -- so a target that
// spreads over adjacent cells is reported once, at its peak.  One thread per bitmask word; the grouped mask goes to
// a second buffer (neighbours are read from the ungrouped one).  oracle/float_path.py:peak_group states the rule.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k3b_peak_group_kernel(const uint32_t* __restrict__ bits_in, uint32_t* __restrict__ bits_out, const float* __restrict__ rd_map,
                      CfarArgs a, long long n_words) {
    const unsigned w = blockIdx.x * 256u + threadIdx.x;            // n_words < 2^31 (the host caps the range lines of a launch)
    if (w >= (unsigned)n_words) return;
    const uint32_t in = __ldg(bits_in + w);
    if (in == 0u) { bits_out[w] = 0u; return; }                   // nearly every word: no index arithmetic at all
    const unsigned words_per_map = (unsigned)a.n_range * (unsigned)a.words_per_row;
    const unsigned map = w / words_per_map, in_map = w - map * words_per_map;
    const int line = (int)(in_map / (unsigned)a.words_per_row), word = (int)(in_map - (unsigned)line * (unsigned)a.words_per_row);
    const uint32_t* bm = bits_in + (size_t)map * words_per_map;
    const float* m = rd_map + (size_t)map * a.n_range * a.n_doppler;
    uint32_t out = in, todo = in;
    while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1u;
        const int col = word * 32 + b;
        const float v = cfar_as_float(__ldg(m + (long long)line * a.n_doppler + col), a.int_mode);
        bool keep = true;
        for (int dl = -1; dl <= 1; ++dl) {
            const int l2 = line + dl;
            if (l2 < 0 || l2 >= a.n_range) continue;
            for (int dc = -1; dc <= 1; ++dc) {
                const int c2 = col + dc;
                if ((dl == 0 && dc == 0) || c2 < 0 || c2 >= a.n_doppler) continue;
                if (!((__ldg(bm + (long long)l2 * a.words_per_row + (c2 >> 5)) >> (c2 & 31)) & 1u)) continue;
                const float v2 = cfar_as_float(__ldg(m + (long long)l2 * a.n_doppler + c2), a.int_mode);
                if (v2 > v || (v2 == v && (dl < 0 || (dl == 0 && dc < 0)))) keep = false;
            }
        }
        if (!keep) out &= ~(1u << b);
    }
    bits_out[w] = out;
}

// ------------------------------------------------------------------------------------------
// K4: ordered compaction, two small kernels.  k4_row_offsets: one 1024-thread CTA per CPI counts the set bits
// of every range line of every channel and exclusive-scans them in (channel, range) order ->
// per-line record offsets, the CPI's detection count and its status word.  k4_emit: one warp per
// line writes the `r d mag` records (radar_core.vhd:396-418, TRUE coordinates) at its offset.
// The list is canonical without a sort; records beyond max_dets are dropped (the LAST ones in
// canonical order) and RDP_STATUS_DET_OVERFLOW is raised (status_overflow analogue, :447-456).
// ------------------------------------------------------------------------------------------
constexpr int K4_THREADS = 1024;

__device__ __forceinline__ int k4_row_count(const uint32_t* __restrict__ bm, int r, int words_per_row) {
    const uint32_t* p = bm + (long long)r * words_per_row;
    int n = 0;
    if ((words_per_row & 3) == 0) {
        for (int w = 0; w < words_per_row; w += 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + w));
            n += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        }
    } else {
        for (int w = 0; w < words_per_row; ++w) n += __popc(__ldg(p + w));
    }
    return n;
}

__global__ void __launch_bounds__(K4_THREADS)
k4_row_offsets_kernel(const uint32_t* __restrict__ det_bits, int* __restrict__ row_off, int* __restrict__ counts,
                      uint32_t* __restrict__ status, CfarArgs a, const uint32_t* __restrict__ cpi_flags) {
    // cpi_flags (nullable): status bits the earlier stages raised for this CPI (RDP_STATUS_WINDOW_SAT / _ADC_FULL_SCALE)
    // every thread owns `per` consecutive lines; one block-wide exclusive scan of the per-thread totals
    __shared__ int warp_sum[K4_THREADS / 32];
    const int cpi = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int rows = a.n_channels * a.n_range;
    const uint32_t* bm = det_bits + (long long)cpi * rows * a.words_per_row;
    int* off = row_off + (long long)cpi * rows;
    const int per = (rows + K4_THREADS - 1) / K4_THREADS;
    const int r_begin = threadIdx.x * per;
    constexpr int KEEP = 8;                                       // line counts kept in registers (loads all in flight)
    int n[KEEP];
    int tot = 0;
    RDP_UNROLL
    for (int i = 0; i < KEEP; ++i) {
        const int r = r_begin + i;
        n[i] = (i < per && r < rows) ? k4_row_count(bm, r, a.words_per_row) : 0;
        tot += n[i];
    }
    for (int i = KEEP; i < per; ++i) {
        const int r = r_begin + i;
        if (r < rows) tot += k4_row_count(bm, r, a.words_per_row);
    }
    int incl = tot;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = warp_sum[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        warp_sum[lane] = v;
    }
    __syncthreads();
    int base = (warp > 0 ? warp_sum[warp - 1] : 0) + incl - tot;
    RDP_UNROLL
    for (int i = 0; i < KEEP; ++i) {
        const int r = r_begin + i;
        if (i < per && r < rows) off[r] = base;
        base += n[i];
    }
    for (int i = KEEP; i < per; ++i) {
        const int r = r_begin + i;
        if (r < rows) {
            off[r] = base;
            base += k4_row_count(bm, r, a.words_per_row);
        }
    }
    if (threadIdx.x == K4_THREADS - 1) {
        counts[cpi] = base;
        status[cpi] = (base > a.max_dets ? (uint32_t)RDP_STATUS_DET_OVERFLOW : 0u) | (cpi_flags ? cpi_flags[cpi] : 0u);
    }
}

__global__ void __launch_bounds__(256)
k4_emit_kernel(const uint32_t* __restrict__ det_bits, const int* __restrict__ row_off, const float* __restrict__ rd_map,
               rdp_det* __restrict__ dets, CfarArgs a, long long total_rows, int cpi_label_base, int lanes_per_row) {
    // lanes_per_row = power of two >= min(words_per_row, 32): a warp serves 32 / lanes_per_row range lines at once.
    // total_rows < 2^31 (the host caps a chunk): 32-bit index arithmetic, the record's CPI / channel / range once per line
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int rows_per_warp = 32 / lanes_per_row;
    const int sub = lane % lanes_per_row;
    const unsigned row = ((unsigned)blockIdx.x * 8u + (unsigned)warp) * (unsigned)rows_per_warp + (unsigned)(lane / lanes_per_row);
    const bool row_ok = row < (unsigned)total_rows;
    const unsigned rows_per_cpi = (unsigned)(a.n_channels * a.n_range);
    const unsigned cpi = row_ok ? row / rows_per_cpi : 0u, line_all = row_ok ? row - cpi * rows_per_cpi : 0u;
    const int channel = (int)(line_all / (unsigned)a.n_range), range = (int)(line_all - (unsigned)channel * (unsigned)a.n_range);
    const uint32_t* bm = det_bits + (size_t)(row_ok ? row : 0u) * a.words_per_row;
    const float* mrow = rd_map + (size_t)(row_ok ? row : 0u) * a.n_doppler;
    int base = row_ok ? __ldg(row_off + row) : 0;
    rdp_det* out = dets + (size_t)cpi * a.max_dets;
    for (int w0 = 0; w0 < a.words_per_row; w0 += lanes_per_row) {
        const int w = w0 + sub;
        uint32_t word = (row_ok && w < a.words_per_row) ? __ldg(bm + w) : 0u;
        const int n = __popc(word);
        int incl = n;                                              // inclusive scan within the line's lanes
        for (int o = 1; o < lanes_per_row; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (sub >= o) incl += t;
        }
        int slot = base + incl - n;
        base += __shfl_sync(0xffffffffu, incl, (lane / lanes_per_row) * lanes_per_row + lanes_per_row - 1);
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            if (slot < a.max_dets) {
                const int col = w * 32 + b;
                rdp_det d;
                d.cpi = cpi_label_base + (int)cpi;
                d.channel = channel;
                d.range = range;
                d.doppler = col;
                d.mag = cfar_as_float(__ldg(mrow + col), a.int_mode);
                out[slot] = d;
            }
            ++slot;
        }
    }
}

}  // namespace rdp
