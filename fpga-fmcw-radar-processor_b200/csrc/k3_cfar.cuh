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
// Phase 1 (every cell, the ALU-bound bulk): a CONSERVATIVE pre-filter against the most
// permissive scale s1 (scale_override, or the smallest of SCALE_MIN/NOM/MAX).  The tile is
// staged in shared memory twice: raw float32, and fl(v*s1) truncated to bfloat16 and duplicated
// into both halves of a word.  Each thread owns S cells along range (one Doppler column), holds
// their cut values rounded UP to bfloat16 packed two per register, and counts with one packed
// compare per two (cell, reference) pairs (HSET2.BF16_V2 -> 0xFFFF masks, summed two at a time by
// IADD3): measured 3.5x the rate of FSETP + predicated add on B200
// (profiles/r1_mb_cmpacc_b200.txt).  ref_lo <= fl(v*s1) < cut <= cut_hi, so the packed count never
// undercounts: no true detection is rejected.
// In adaptive mode the pre-filter threshold is tightened per cell before it is rounded up: a cell can only
// be detected with the scale the RTL would pick, so
//     detection  =>  x_k < T = max(cut / SCALE_MAX, cut / SCALE_NOM, min(cut / SCALE_MIN, lo'))
// where lo' >= lo = mean / 2 is an upper bound of the reference mean (the sum over the whole window, columns
// widened to aligned groups of four, from a separable box filter over the staged tile).  SCALE_MIN is only
// chosen when x_k < lo (:197), i.e. next to strong cells, so ordinary noise cells are tested against
// cut / SCALE_NOM instead of cut / SCALE_MIN and almost never reach phase 2 (bench maps: survivors 5.0 % ->
// 2.1 % of cells with 1.75 % true detections).  All bounds carry a 1e-3 relative margin, far above float32
// rounding and far below the bfloat16 step; the decision itself stays with the exact phase 2.
// Phase 2 (survivors only, ~0.5 % of noise cells + target cells): one warp per survivor decides
// EXACTLY in float32 from the raw tile -- the s1 test, then (adaptive mode) float64 sum -> mean
// -> scale 2/4/6 (:189-199) -> final test.  Decisions set bits in a per-tile bitmask.
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
    // pre-filter threshold in the s1-scaled domain: Ts = max(cut * pf_a, min(cut * pf_b, window_sum * pf_c))
    // (non-adaptive: pf_a = 1, pf_b = 0, i.e. Ts = cut)
    float pf_a, pf_b, pf_c;
};

// magnitude word -> float for the pre-filter; integers below 2^24 convert exactly
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
// pre-filter reference value as float (int_mode: the integer word; below 2^24 the conversion is exact).  The
// scale s1 is NOT applied to the references: 1 / s1 (with a 1e-3 margin) is folded into the per-cell threshold.
__device__ __forceinline__ float cfar_scaled(float raw, float s1, bool int_mode) {
    (void)s1;
    return int_mode ? (float)(uint32_t)__float_as_int(raw) : raw;
}
// bfloat16 truncation of v (>= 0), duplicated into both halves of a word
__device__ __forceinline__ uint32_t bf16_dup(float v) {
#ifdef RDP_EMU
    const uint32_t lo = (uint32_t)__float_as_int(v) >> 16;
    return lo | (lo << 16);
#else
    return __byte_perm((uint32_t)__float_as_int(v), 0u, 0x3232);
#endif
}
// float (>= 0) -> bfloat16 bits rounded toward zero / toward +inf
__device__ __forceinline__ uint32_t bf16_down(float v) { return (uint32_t)__float_as_int(v) >> 16; }
__device__ __forceinline__ uint32_t bf16_up(float v) { return ((uint32_t)__float_as_int(v) + 0xffffu) >> 16; }

// Exact decision for one cell by one warp from GLOBAL memory (generic kernel).
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
    static constexpr int SROWS = TROWS + 2 * HL;
    // TMA needs the box to START on a 16-byte boundary in the innermost dimension (an unaligned start
    // coordinate traps with "illegal instruction", tools/microbench/mb_tma_probe.cu), so the column halo is
    // padded to a multiple of 4 floats on both sides: tile column c sits at shared column c + HCP
    static constexpr int HCP = (HC + 3) / 4 * 4;
    static constexpr int SCOLS = TCOLS + 2 * HCP;
    static constexpr int WIN = (2 * HL + 1) * (2 * HC + 1);
    static constexpr int MAXQ = (WIN + 31) / 32;
    static constexpr int TWORDS = (TCOLS + 31) / 32;              // bitmask words per tile row
    static constexpr int NREF = WIN - (2 * GL + 1) * (2 * GC + 1);
    // phase 2: G lanes per survivor (32 / G survivors per warp pass), NQ window slots per lane
    static constexpr int G = WIN <= 160 ? 8 : (WIN <= 320 ? 16 : 32);
    // (the slot table lists the reference cells only, padded to a multiple of G with slots that read a zero cell)
    static constexpr int NQ = (NREF + G - 1) / G, NSLOT = NQ * G, NZ = NSLOT - NREF;
    static constexpr int TF_WORDS = (SROWS * SCOLS + 4 + 31) / 32 * 32;   // one raw stage (+ a cell that reads 0.0f), 128 B multiple
    static constexpr int ZERO_IDX = SROWS * SCOLS;
    static constexpr int TILE_BYTES = SROWS * SCOLS * 4;
    // window-sum bound of the adaptive pre-filter: column groups of 4, NCOV groups cover any window
    static constexpr int NCG = SCOLS / 4, NCOV = (2 * HC + 7) / 4;
    static constexpr int SURV_WORDS = TROWS * TCOLS / 2;
    static_assert(TROWS * TCOLS <= 65536, "survivor codes are 16-bit");
    static_assert(4 * NCOV >= 2 * HC + 1 + 3, "NCOV groups cover the window at any alignment");
    static constexpr size_t smem_bytes() {
        return sizeof(float) * (2 * TF_WORDS + SROWS * SCOLS) +
               sizeof(int) * (SURV_WORDS + SROWS * NCG + TROWS * TWORDS + 4 + 2 * NSLOT) + 2 * sizeof(unsigned long long) + 128;
    }
};


#ifdef RDP_EMU
struct CUtensorMap_st { alignas(64) unsigned char opaque[128]; };
typedef CUtensorMap_st CUtensorMap;
#define RDP_GRID_CONSTANT
#else
#include <cuda.h>          // CUtensorMap (type only; nothing from libcuda is linked)
#define RDP_GRID_CONSTANT __grid_constant__
// mbarrier + TMA (cp.async.bulk.tensor) primitives, sm_90+ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 3-D tiled TMA load global -> shared; out-of-bounds elements are written as 0 (== cfar_edges ZERO)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, unsigned long long* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
#endif

// one packed compare for cells (2p, 2p+1) of the strip against tile row lr, if either is a reference
template <int HL, int GL, bool GUARD_COL, int P, int LR>
__device__ __forceinline__ void k3_pair(uint32_t v, const uint32_t (&cb)[4], const uint32_t (&cl)[4],
                                        const uint32_t (&ch)[4], uint32_t (&acc)[4]) {
    constexpr int dl0 = LR - 2 * P - HL, dl1 = dl0 - 1;            // line offset of row LR from cell 2p, 2p+1
    constexpr bool in0 = dl0 >= -HL && dl0 <= HL && !(GUARD_COL && dl0 >= -GL && dl0 <= GL);
    constexpr bool in1 = dl1 >= -HL && dl1 <= HL && !(GUARD_COL && dl1 >= -GL && dl1 <= GL);
    if (in0 && in1) acc[P] -= bf16x2_lt_mask(v, cb[P]);
    else if (in0) acc[P] -= bf16x2_lt_mask(v, cl[P]);
    else if (in1) acc[P] -= bf16x2_lt_mask(v, ch[P]);
}

template <int HL, int GL, bool GUARD_COL, int LR, int NROWS>
struct K3Rows {
    static __device__ __forceinline__ void run(const uint32_t* __restrict__ tp, int pitch, const uint32_t (&cb)[4],
                                               const uint32_t (&cl)[4], const uint32_t (&ch)[4], uint32_t (&acc)[4]) {
        const uint32_t v = tp[LR * pitch];
        k3_pair<HL, GL, GUARD_COL, 0, LR>(v, cb, cl, ch, acc);
        k3_pair<HL, GL, GUARD_COL, 1, LR>(v, cb, cl, ch, acc);
        k3_pair<HL, GL, GUARD_COL, 2, LR>(v, cb, cl, ch, acc);
        k3_pair<HL, GL, GUARD_COL, 3, LR>(v, cb, cl, ch, acc);
        K3Rows<HL, GL, GUARD_COL, LR + 1, NROWS>::run(tp, pitch, cb, cl, ch, acc);
    }
};
template <int HL, int GL, bool GUARD_COL, int NROWS>
struct K3Rows<HL, GL, GUARD_COL, NROWS, NROWS> {
    static __device__ __forceinline__ void run(const uint32_t* __restrict__, int, const uint32_t (&)[4],
                                               const uint32_t (&)[4], const uint32_t (&)[4], uint32_t (&)[4]) {}
};

template <int HL, int HC, int GL, int GC, int S, int TCOLS, bool INT_MODE>
__global__ void __launch_bounds__(256)
k3_cfar_kernel(const float* __restrict__ rd_map, float* __restrict__ cfar_out, uint32_t* __restrict__ det_bits,
               CfarArgs a, int tiles_r, int tiles_c, long long n_tiles, int use_tma,
               const RDP_GRID_CONSTANT CUtensorMap tmap) {
    static_assert(S == 8, "the strip is four packed pairs");
    using Sh = K3Shape<HL, HC, GL, GC, S, TCOLS>;
    RDP_DYN_SMEM(smem_raw);
    // 128-byte aligned carve-up: [2 raw stages][packed tile][survivors][group sums][bitmask][counter][window table][mbarriers]
    unsigned char* smem_al = smem_raw + ((128 - ((uintptr_t)smem_raw & 127)) & 127);
    float* tf0 = reinterpret_cast<float*>(smem_al);                            // raw magnitudes, two stages
    uint32_t* th = reinterpret_cast<uint32_t*>(tf0 + 2 * Sh::TF_WORDS);        // bf16x2 (lo, lo) of fl(v*s1)
    unsigned short* surv = reinterpret_cast<unsigned short*>(th + Sh::SROWS * Sh::SCOLS);   // 16-bit survivor codes
    float* grp = reinterpret_cast<float*>(surv) + Sh::SURV_WORDS;              // [SROWS][NCG] sums of 4 NCOV columns
    uint32_t* bits = reinterpret_cast<uint32_t*>(grp + Sh::SROWS * Sh::NCG);
    int* n_surv = reinterpret_cast<int*>(bits + Sh::TROWS * Sh::TWORDS);
    int* wtab = n_surv + 4;                                                    // [NSLOT][2]: tf index = ctr * mul + add
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(wtab + 2 * Sh::NSLOT);
    const int tiles_per_map = tiles_r * tiles_c;

    if (threadIdx.x == 0) {
        tf0[Sh::ZERO_IDX] = 0.0f;
        tf0[Sh::TF_WORDS + Sh::ZERO_IDX] = 0.0f;
#ifndef RDP_EMU
        if (use_tma) {
            mbar_init(mbar, 1);
            mbar_init(mbar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
#endif
    }
    for (int w = threadIdx.x; w < Sh::WIN + Sh::NZ; w += 256) {   // slot table: the reference cells in raster order, then padding
        if (w < Sh::WIN) {
            const int dl = w / (2 * HC + 1) - HL, dc = w % (2 * HC + 1) - HC;
            if (!(dl >= -GL && dl <= GL && dc >= -GC && dc <= GC)) {
                // guard cells in front of this one: whole guard rows above, part of its own row
                int gb = (dl > GL ? 2 * GL + 1 : (dl > -GL ? dl + GL : 0)) * (2 * GC + 1);
                if (dl >= -GL && dl <= GL) gb += dc > GC ? 2 * GC + 1 : 0;
                const int sl = w - gb;
                wtab[2 * sl] = 1;
                wtab[2 * sl + 1] = dl * Sh::SCOLS + dc;
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
        tma_load_3d(tf0 + st * Sh::TF_WORDS, tmap_ptr, mbar + st, (rem % tiles_c) * TCOLS - Sh::HCP, (rem / tiles_c) * Sh::TROWS - HL, (int)mp);
    };
    if (use_tma && threadIdx.x == 0 && (long long)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
#endif

    int iter = 0;
    const int n_tiles32 = (int)n_tiles;
    for (int tile_id = blockIdx.x; tile_id < n_tiles32; tile_id += gridDim.x, ++iter) {
    const int stage = use_tma ? (iter & 1) : 0;
    float* tf = tf0 + stage * Sh::TF_WORDS;
    const long long map = tile_id / tiles_per_map;
    const int in_map = tile_id - (int)map * tiles_per_map;
    const int tr = in_map / tiles_c, tc = in_map - tr * tiles_c;
    const int r0 = tr * Sh::TROWS, c0 = tc * TCOLS;
    const float* m = rd_map + map * (long long)a.n_range * a.n_doppler;

    if (threadIdx.x == 0) *n_surv = 0;
    for (int i = threadIdx.x; i < Sh::TROWS * Sh::TWORDS; i += 256) bits[i] = 0u;
#ifndef RDP_EMU
    if (use_tma) {
        // prefetch the next tile into the other stage (its previous readers finished before the barrier that
        // ended the previous iteration), then wait for this tile's bytes
        if (threadIdx.x == 0 && tile_id + gridDim.x < n_tiles) issue(tile_id + gridDim.x, stage ^ 1);
        mbar_wait(mbar + stage, (uint32_t)((iter >> 1) & 1));
        // raw words -> packed bf16x2 (lo, lo), four cells per thread per step (128-bit shared-memory accesses)
        // int_mode: the product is rounded toward zero so that ref_lo never exceeds the exact integer product
        const float4* tf4 = reinterpret_cast<const float4*>(tf);
        uint4* th4 = reinterpret_cast<uint4*>(th);
        if (a.adaptive) {
            // ... and, for the adaptive pre-filter bound, grp[row][g] = sum of the 4 NCOV columns from group g on
            // (each warp steps by 32 - (NCOV - 1) chunks so that the neighbours it needs sit in its own lanes)
            constexpr int OV = Sh::NCOV - 1, ADV = 32 - OV, NCH = Sh::SROWS * Sh::SCOLS / 4;
            const int lane_s = threadIdx.x % 32;
            for (int base = (threadIdx.x / 32) * ADV; base < NCH; base += 8 * ADV) {
                const int i = base + lane_s;
                float g = 0.0f;
                if (i < NCH) {
                    const float4 r = tf4[i];
                    if (lane_s < ADV) {
                        uint4 o;
                        o.x = bf16_dup(cfar_scaled(r.x, a.s1, INT_MODE));
                        o.y = bf16_dup(cfar_scaled(r.y, a.s1, INT_MODE));
                        o.z = bf16_dup(cfar_scaled(r.z, a.s1, INT_MODE));
                        o.w = bf16_dup(cfar_scaled(r.w, a.s1, INT_MODE));
                        th4[i] = o;
                    }
                    g = (cfar_as_float(r.x, INT_MODE) + cfar_as_float(r.y, INT_MODE)) +
                        (cfar_as_float(r.z, INT_MODE) + cfar_as_float(r.w, INT_MODE));
                }
                float cover = g;
                const int cg = i % Sh::NCG;                       // groups past the end of the row count as 0
                RDP_UNROLL
                for (int k = 1; k <= OV; ++k) {
                    const float nb = __shfl_down_sync(0xffffffffu, g, k);
                    cover += cg + k < Sh::NCG ? nb : 0.0f;
                }
                if (lane_s < ADV && i < NCH) grp[i] = cover;
            }
        } else {
            for (int i = threadIdx.x; i < Sh::SROWS * Sh::SCOLS / 4; i += 256) {
                const float4 r = tf4[i];
                uint4 o;
                o.x = bf16_dup(cfar_scaled(r.x, a.s1, INT_MODE));
                o.y = bf16_dup(cfar_scaled(r.y, a.s1, INT_MODE));
                o.z = bf16_dup(cfar_scaled(r.z, a.s1, INT_MODE));
                o.w = bf16_dup(cfar_scaled(r.w, a.s1, INT_MODE));
                th4[i] = o;
            }
        }
    } else
#endif
    {   // tile + halo, one warp per tile row (coalesced, no per-element index arithmetic)
        const int lw = threadIdx.x / 32, ll = threadIdx.x % 32;
        for (int lr = lw; lr < Sh::SROWS; lr += 8) {
            const int line = r0 - HL + lr;
            const bool row_ok = a.raster || (line >= 0 && line < a.n_range);
            const long long row_base = (long long)line * a.n_doppler;
            RDP_UNROLL
            for (int k = 0; k < (Sh::SCOLS + 31) / 32; ++k) {
                const int lc = ll + 32 * k;
                if (lc < Sh::SCOLS) {
                    const int c = c0 - Sh::HCP + lc;
                    float v = 0.0f;
                    if (a.raster) {
                        const long long flat = row_base + c;
                        if (flat >= 0 && flat < (long long)a.n_range * a.n_doppler) v = __ldg(m + flat);
                    } else if (row_ok && c >= 0 && c < a.n_doppler) {
                        v = __ldg(m + row_base + c);
                    }
                    tf[lr * Sh::SCOLS + lc] = v;
                    th[lr * Sh::SCOLS + lc] = bf16_dup(cfar_scaled(v, a.s1, INT_MODE));
                }
            }
        }
        if (a.adaptive) {
            __syncthreads();
            for (int i = threadIdx.x; i < Sh::SROWS * Sh::NCG; i += 256) {
                float cover = 0.0f;
                RDP_UNROLL
                for (int k = 0; k < 4 * Sh::NCOV; ++k)
                    if (4 * (i % Sh::NCG) + k < Sh::SCOLS) cover += cfar_as_float(tf[4 * i + k], INT_MODE);
                grp[i] = cover;
            }
        }
    }
    __syncthreads();

    // ---- phase 1: conservative packed pre-filter ---------------------------------------------------
    const int tx = threadIdx.x % TCOLS, ty = threadIdx.x / TCOLS;
    const int col = c0 + tx;
    uint32_t cb[4], cl[4], ch[4], acc[4];
    uint32_t fmask = 0u;                                          // bit j: cell j of the strip goes to the exact phase regardless
    float wsum[S];                                                // adaptive: pf_c x bound of the window sum of each cell
    if (a.adaptive) {
        // vertical pass of the box filter in registers: wsum[j] = sum_{k = j .. j + 2 HL} x[k], additions only
        const float* gp = grp + (ty * S) * Sh::NCG + ((tx + Sh::HCP - HC) >> 2);
        float x[S + 2 * HL];
        RDP_UNROLL
        for (int k = 0; k < S + 2 * HL; ++k) x[k] = gp[k * Sh::NCG];
        if constexpr (2 * HL >= S - 1) {
            float common = x[S - 1];                              // lines S-1 .. 2 HL belong to every window
            RDP_UNROLL
            for (int k = S; k <= 2 * HL; ++k) common += x[k];
            common *= a.pf_c;                                     // wsum comes out already times pf_c
            float part = 0.0f;
            wsum[S - 1] = common;
            RDP_UNROLL
            for (int j = S - 2; j >= 0; --j) { part += x[j]; wsum[j] = fmaf(part, a.pf_c, common); }
            part = 0.0f;
            RDP_UNROLL
            for (int j = 1; j < S; ++j) { part += x[2 * HL + j]; wsum[j] = fmaf(part, a.pf_c, wsum[j]); }
        } else {
            RDP_UNROLL
            for (int j = 0; j < S; ++j) {
                wsum[j] = 0.0f;
                RDP_UNROLL
                for (int k = 0; k <= 2 * HL; ++k) wsum[j] += x[j + k];
                wsum[j] *= a.pf_c;
            }
        }
    }
    RDP_UNROLL
    for (int p = 0; p < 4; ++p) {
        const float cut_lo = cfar_as_float(tf[(ty * S + 2 * p + HL) * Sh::SCOLS + tx + Sh::HCP], INT_MODE);
        const float cut_hi = cfar_as_float(tf[(ty * S + 2 * p + 1 + HL) * Sh::SCOLS + tx + Sh::HCP], INT_MODE);
        float c_lo = cut_lo, c_hi = cut_hi;
        if (a.adaptive) {
            c_lo = fmaxf(c_lo * a.pf_a, fminf(c_lo * a.pf_b, wsum[2 * p]));
            c_hi = fmaxf(c_hi * a.pf_a, fminf(c_hi * a.pf_b, wsum[2 * p + 1]));
        } else {
            c_lo *= a.pf_a;
            c_hi *= a.pf_a;
        }
        const uint32_t h0 = bf16_up(c_lo), h1 = bf16_up(c_hi);
        cb[p] = h0 | (h1 << 16);
        cl[p] = h0;
        ch[p] = h1 << 16;
        acc[p] = 0u;
        // bfloat16 subnormal range: leave the decision to the exact phase
        // positive cuts so small that the threshold may fall into the bfloat16 subnormal range (pf_a >= 1e-6) go
        // to the exact phase; one unsigned compare on the bit pattern covers 0 < cut < 1e-30
        fmask |= ((uint32_t)__float_as_int(cut_lo) - 1u < 0x0da24260u - 1u) ? (1u << (2 * p)) : 0u;
        fmask |= ((uint32_t)__float_as_int(cut_hi) - 1u < 0x0da24260u - 1u) ? (2u << (2 * p)) : 0u;
    }
    const uint32_t* tp = th + (ty * S) * Sh::SCOLS + tx + Sh::HCP;
    for (int dc = -HC; dc <= 0; ++dc) {
        if (dc >= -GC) K3Rows<HL, GL, true, 0, S + 2 * HL>::run(tp + dc, Sh::SCOLS, cb, cl, ch, acc);
        else K3Rows<HL, GL, false, 0, S + 2 * HL>::run(tp + dc, Sh::SCOLS, cb, cl, ch, acc);
    }
    // Half-way exit: a cell needs rank_idx + 1 of its n_ref references below the threshold, so of the HALF_REFS
    // references in columns -HC .. 0 at least HALF_REFS - (n_ref - rank_idx - 1) must be.  If no cell of the warp's
    // 32 x S block gets there (noise-only blocks: almost never), the remaining columns cannot change any decision.
    bool go_on = true;
    {
        constexpr int HALF_REFS = (HC - GC) * (2 * HL + 1) + (GC + 1) * (2 * HL - 2 * GL);
        const int need = HALF_REFS - (a.n_ref - a.rank_idx - 1);
        if (need > 0) {
            bool mine = fmask != 0u;
            RDP_UNROLL
            for (int p = 0; p < 4; ++p) {
                const uint32_t lo16 = acc[p] & 0xffffu, hi16 = ((acc[p] >> 16) + lo16) & 0xffffu;
                mine = mine || (int)lo16 >= need || (int)hi16 >= need;
            }
            go_on = __any_sync(0xffffffffu, mine);
        }
    }
    if (go_on) {
        for (int dc = 1; dc <= HC; ++dc) {
            if (dc <= GC) K3Rows<HL, GL, true, 0, S + 2 * HL>::run(tp + dc, Sh::SCOLS, cb, cl, ch, acc);
            else K3Rows<HL, GL, false, 0, S + 2 * HL>::run(tp + dc, Sh::SCOLS, cb, cl, ch, acc);
        }
    }
    {
        // survivors of this strip as a bit mask; most strips have none and skip the list append altogether
        uint32_t smask = fmask;
        RDP_UNROLL
        for (int p = 0; p < 4; ++p) {
            const uint32_t lo16 = acc[p] & 0xffffu, hi16 = ((acc[p] >> 16) + lo16) & 0xffffu;
            smask |= ((int)lo16 >= a.rank_idx + 1) ? (1u << (2 * p)) : 0u;
            smask |= ((int)hi16 >= a.rank_idx + 1) ? (2u << (2 * p)) : 0u;
        }
        const int lines_left = a.n_range - (r0 + ty * S);            // strips / columns past the edge of the map
        if (col >= a.n_doppler || lines_left <= 0) smask = 0u;
        else if (lines_left < S) smask &= (1u << lines_left) - 1u;
        if (cfar_out && col < a.n_doppler) {
            RDP_UNROLL
            for (int j = 0; j < S; ++j)
                if (j < lines_left) cfar_out[map * (long long)a.n_range * a.n_doppler + (long long)(r0 + ty * S + j) * a.n_doppler + col] = 0.0f;
        }
        if (smask) {
            int slot = atomicAdd(n_surv, __popc(smask));
            while (smask) {
                const int j = __ffs(smask) - 1;
                smask &= smask - 1u;
                surv[slot++] = (unsigned short)((ty * S + j) * TCOLS + tx);
            }
        }
    }
    __syncthreads();

    // ---- phase 2: exact decision from the raw tile: G lanes per survivor, 32/G survivors per warp pass ---
    // Window slots are dealt to the G lanes (slot = sub + G q); guard / padding slots read the zero cell and
    // their known contribution to each count is subtracted.  Counting = sign bit of an exact float subtraction.
    const int ns = *n_surv;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    constexpr int G = Sh::G, NQ = Sh::NQ, NZ = Sh::NZ;
    const int sub = lane % G, grp = lane / G;
    const bool n_ref_pow2 = (a.n_ref & (a.n_ref - 1)) == 0;
    const double inv_n = 1.0 / (double)a.n_ref;
    for (int i0 = warp * (32 / G); i0 < ns; i0 += 8 * (32 / G)) {
        const int i = i0 + grp;
        const bool valid = i < ns;
        const int code = valid ? (int)surv[i] : 0;                 // idle groups chew on tile cell (0, 0)
        const int tl = code / TCOLS, tcx = code % TCOLS;
        const int cidx = (tl + HL) * Sh::SCOLS + tcx + Sh::HCP;
        bool det;
        if (!INT_MODE) {
        const float cut = tf[cidx];
        float v[NQ];
        RDP_UNROLL
        for (int q = 0; q < NQ; ++q) {
            const int2 w = *reinterpret_cast<const int2*>(wtab + 2 * (q * G + sub));
            v[q] = tf[cidx * w.x + w.y];
        }
        float scale = a.s1;
        if (a.adaptive) {
            double s = 0.0;
            RDP_UNROLL
            for (int q = 0; q < NQ; ++q) s += (double)v[q];
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
        RDP_UNROLL
        for (int q = 0; q < NQ; ++q) {
            const int2 w = *reinterpret_cast<const int2*>(wtab + 2 * (q * G + sub));
            v[q] = (int)ti[cidx * w.x + w.y];
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
        const float cut = tf[cidx];                                // raw word: written back unchanged in either mode
        if (det && sub == 0) {
            atomicOr(bits + tl * Sh::TWORDS + (tcx >> 5), 1u << (tcx & 31));
            if (cfar_out)
                cfar_out[map * (long long)a.n_range * a.n_doppler + (long long)(r0 + tl) * a.n_doppler + c0 + tcx] = cut;
        }
    }
    __syncthreads();
    // bitmask of this tile: rows r0.., words (c0 / 32)..  (TCOLS is a multiple of 32 so tiles own whole words)
    uint32_t* bm = det_bits + map * (long long)a.n_range * a.words_per_row;
    for (int i = threadIdx.x; i < Sh::TROWS * Sh::TWORDS; i += 256) {
        const int rl = i / Sh::TWORDS, w = i % Sh::TWORDS;
        const int line = r0 + rl, word = c0 / 32 + w;
        if (line < a.n_range && word < a.words_per_row) bm[(long long)line * a.words_per_row + word] = bits[i];
    }
    __syncthreads();                                      // the tile's shared state is reused by the next tile
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
            const bool det = cfar_exact_warp<MAXQ_GENERIC>(a, m, line, col, cv, lane);
            if (det) out |= 1u << b;
            if (lane == 0 && cfar_out) cfar_out[map * cells_per_map + (long long)line * a.n_doppler + col] = det ? cv : 0.0f;
        }
        if (lane == 0) det_bits[w] = out;
    }
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
                      uint32_t* __restrict__ status, CfarArgs a) {
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
        status[cpi] = base > a.max_dets ? (uint32_t)RDP_STATUS_DET_OVERFLOW : 0u;
    }
}

__global__ void __launch_bounds__(256)
k4_emit_kernel(const uint32_t* __restrict__ det_bits, const int* __restrict__ row_off, const float* __restrict__ rd_map,
               rdp_det* __restrict__ dets, CfarArgs a, long long total_rows, int cpi_label_base, int lanes_per_row) {
    // lanes_per_row = power of two >= min(words_per_row, 32): a warp serves 32 / lanes_per_row range lines at once
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int rows_per_warp = 32 / lanes_per_row;
    const int sub = lane % lanes_per_row;
    const long long row = ((long long)blockIdx.x * 8 + warp) * rows_per_warp + lane / lanes_per_row;
    const bool row_ok = row < total_rows;
    const int rows_per_cpi = a.n_channels * a.n_range;
    const int cpi = row_ok ? (int)(row / rows_per_cpi) : 0, line_all = row_ok ? (int)(row % rows_per_cpi) : 0;
    const uint32_t* bm = det_bits + (row_ok ? row : 0) * a.words_per_row;
    int base = row_ok ? __ldg(row_off + row) : 0;
    rdp_det* out = dets + (long long)cpi * a.max_dets;
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
                d.cpi = cpi_label_base + cpi;
                d.channel = line_all / a.n_range;
                d.range = line_all % a.n_range;
                d.doppler = col;
                d.mag = cfar_as_float(__ldg(rd_map + row * a.n_doppler + col), a.int_mode);
                out[slot] = d;
            }
            ++slot;
        }
    }
}

}  // namespace rdp
