// k1_range.cuh -- K1: range window (x) range FFT (x) corner turn, one pass over the cube.
//
// Replaces window_multiplier (N_SAMPLES = N_RANGE) -> xfft_range -> corner_turner of
// rtl/src/radar_core.vhd:267-327.  A CTA owns a tile of TC consecutive chirps of one
// (CPI, channel): each FFT group transforms one chirp at a time in registers and drops the
// spectrum into a padded shared-memory tile [range][chirp]; the tile is then written out
// range-major ([range][chirp], corner_turner.vhd:79-80) as TC*8-byte contiguous segments,
// so the transpose costs no extra pass over HBM.
//   algorithmic HBM bytes / cell: B_in (4 int16-IQ | 8 complex64) read + 8 written.
// With 16-chirp tiles the tile goes out as TMA stores (cp.async.bulk.tensor, SASS UTMASTG): the spectra are written
// into a dense [range][16 chirps] box in the tensor map's 128-byte swizzle (the XOR of the 16-byte chunk index with
// the row keeps the column writes of the FFT epilogue spread over the banks), one elected thread issues the
// 256-row boxes, and the CTA is already transforming the next tile while the TMA unit drains this one -- no store loop,
// no second pass of the tile through the LSU.
//
// Long chirps (2048 / 4096 samples) leave room for only 8 / 4 chirps of a tile in one SM's shared memory, i.e.
// 64- / 32-byte output segments.  For those the kernel runs as thread-block CLUSTERS of CS = 2 / 4 CTAs: each
// CTA transforms its own TC_MAX chirps but owns N / CS range bins of the cluster's CS * TC_MAX chirps -- the
// spectra are written straight into the owning CTA's shared memory (distributed shared memory,
// st.shared::cluster), so every CTA stores 128-byte segments.  Measured on B200 this only pays when the segments
// are peer stores over NVLink (sharded mode, +15 %); into local HBM the 8-byte remote shared-memory stores cost
// more than the short segments did (2.4x slower), so the host uses the plain kernel there.
#pragma once
#include "fft_core.cuh"

namespace rdp {

template <class Plan>
struct K1Shape {
    static constexpr int THREADS = (Plan::T >= 256) ? Plan::T : 256;
    static constexpr int G = THREADS / Plan::T;           // transforms in flight per CTA
    // chirps per tile: 128-byte output segments (16 x complex64) unless shared memory forbids
    static constexpr int TC_MAX = (Plan::N <= 1024) ? (G > 16 ? G : 16) : (Plan::N == 2048 ? 8 : 4);
    static __host__ __device__ constexpr size_t smem_bytes(int tile_chirps) {
        return sizeof(float2) * ((size_t)Plan::N * (tile_chirps + 1) + (size_t)G * Plan::XBUFS) + 1024;   // + slack: the tile starts on a 1 KiB boundary (swizzle atom)
    }
    static constexpr int BOX_ROWS = Plan::N < 256 ? Plan::N : 256;    // range lines per TMA store box
    // cluster variant: CS CTAs, each holding N / CS range bins of CS * TC_MAX chirps
    static constexpr int CS = (Plan::N == 4096) ? 4 : (Plan::N == 2048 ? 2 : 1);
    static constexpr int CL_ROWS = Plan::N / CS, CL_COLS = CS * TC_MAX, CL_PITCH = CL_COLS + 1;
    static __host__ __device__ constexpr size_t smem_bytes_cluster() {
        return sizeof(float2) * ((size_t)CL_ROWS * CL_PITCH + (size_t)G * Plan::XBUFS);
    }
};

#ifndef RDP_EMU
__device__ __forceinline__ uint32_t k1_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t k1_mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void k1_st_cluster(uint32_t addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void k1_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
#endif

// raw sample word(s) of one thread for one chirp: int16 IQ = one 32-bit word {Q[31:16], I[15:0]}
// (window_multiplier.vhd:93-94), complex64 = one float2
template <int IN_FMT>
struct K1Raw;
template <>
struct K1Raw<0> {
    typedef uint32_t T;
    static __device__ __forceinline__ T zero() { return 0u; }
    static __device__ __forceinline__ T load(const void* __restrict__ cube, long long idx) {
        return __ldg(reinterpret_cast<const uint32_t*>(cube) + idx);
    }
    static __device__ __forceinline__ float2 cvt(T w) {
        return make_float2((float)(short)(w & 0xffffu), (float)((int)w >> 16));
    }
};
template <>
struct K1Raw<1> {
    typedef float2 T;
    static __device__ __forceinline__ T zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ T load(const void* __restrict__ cube, long long idx) {
        return __ldg(reinterpret_cast<const float2*>(cube) + idx);
    }
    static __device__ __forceinline__ float2 cvt(T w) { return w; }
};

// Distributed corner turn (BASELINE config 4): the chirps of a CPI are sharded over `world` GPUs, the range
// lines of the range-major intermediate over the same GPUs.  K1 then stores each range line straight into the
// memory of the GPU that owns it (peer pointers opened through CUDA IPC, stores travel over NVLink): the
// all-to-all exchange is fused into the transposed store instead of being a separate collective + repack.
struct K1Scatter {
    float2* peer[16];          // x_rc of every rank: [map][rows_local][n_chirps_total]
    int world, rank;
    int rows_shift;            // rows_local = 1 << rows_shift range lines per rank
    int n_chirps_total;        // world * local chirps
};

template <class Plan, int IN_FMT>
__global__ void __launch_bounds__(K1Shape<Plan>::THREADS)
k1_range_kernel(const void* __restrict__ cube, float2* __restrict__ x_rc,
                const float* __restrict__ window, const float2* __restrict__ twtab,
                int n_chirps, int tile_chirps, int tiles_per_map, long long total_tiles, K1Scatter sc,
                int use_tma_store, const RDP_GRID_CONSTANT CUtensorMap tmap_out,
                uint32_t* __restrict__ cpi_flags, int n_channels) {
    // cpi_flags (nullable, int16 input only): word [map / n_channels] |= RDP_STATUS_ADC_FULL_SCALE when a sample of the
    // tile sits on a rail -- a running max of |I|, |Q| per thread (two FMNMX per word), one vote per warp and tile
    constexpr int N = Plan::N, E = Plan::E, T = Plan::T;
    constexpr int G = K1Shape<Plan>::G, THREADS = K1Shape<Plan>::THREADS, TCM = K1Shape<Plan>::TC_MAX;
    using P0 = typename Plan::P0;
    using PL = typename Plan::LAST;
    using Raw = K1Raw<IN_FMT>;
    RDP_DYN_SMEM(smem_raw);
    float2* tile = reinterpret_cast<float2*>(smem_raw + ((1024 - ((uintptr_t)smem_raw & 1023)) & 1023));
    unsigned char* const tile_bytes = reinterpret_cast<unsigned char*>(tile);
    bool store_pending = false;                            // a TMA store of the previous tile may still be reading the box
    const int tcp = tile_chirps + 1;                       // +1 complex: conflict-free column writes
    float2* xbuf = tile + (size_t)N * tcp + (size_t)(threadIdx.x / T) * Plan::XBUFS;
    const int tid = threadIdx.x % T, gid = threadIdx.x / T;

    float2 tw[Plan::NTW > 0 ? Plan::NTW : 1];
    Plan::load_twiddles(tw, twtab, tid);
    float wv[E];
    RDP_UNROLL
    for (int b = 0; b < P0::NB; ++b) {
        RDP_UNROLL
        for (int r = 0; r < P0::R; ++r) wv[b * P0::R + r] = __ldg(window + P0::in_index(tid, b, r));
    }
    int parity = 0;
    const int iters = (tile_chirps + G - 1) / G;

    // software pipeline: the raw samples of the NEXT chirp this group transforms are in flight while the
    // current one is in the butterflies
    typename Raw::T raw[E];
    const int n_tiles = (int)total_tiles;                  // < 2^31 (checked by the host): 32-bit index arithmetic
    auto fetch = [&](int t, int it) {
        const int c = it * G + gid;
        const bool ok = t < n_tiles && c < tile_chirps;
        const int mp = t / tiles_per_map;
        const int chirp = mp * n_chirps + (t - mp * tiles_per_map) * tile_chirps + c;     // global chirp index
        const long long first = (long long)chirp * N + tid;
        RDP_UNROLL
        for (int b = 0; b < P0::NB; ++b) {
            RDP_UNROLL
            for (int r = 0; r < P0::R; ++r)
                raw[b * P0::R + r] = ok ? Raw::load(cube, first + (b * T + r * (N / P0::R))) : Raw::zero();
        }
    };
    fetch((int)blockIdx.x, 0);
    // the raw samples of the NEXT tile are pulled into L2 while this one is transformed, so that the register
    // prefetch above meets L2 latency instead of loaded-HBM latency (tiles are contiguous: tile_chirps * N samples)
    auto l2_prefetch = [&](int t) {
#ifndef RDP_EMU
        const int bytes = tile_chirps * N * (int)sizeof(typename Raw::T);
        if (t < n_tiles && bytes <= 64 * 1024) {               // measured: helps up to 64 KB tiles, hurts 128 KB ones
            const char* p = reinterpret_cast<const char*>(cube) + (long long)t * tile_chirps * N * (long long)sizeof(typename Raw::T);
            for (int off = threadIdx.x * 128; off < bytes; off += THREADS * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
        }
#else
        (void)t;
#endif
    };

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mp = t / tiles_per_map;
        const long long map = mp;
        const int c0 = (t - mp * tiles_per_map) * tile_chirps;
        l2_prefetch(t + (int)gridDim.x);
        float peak = 0.0f;
        for (int it = 0; it < iters; ++it) {
            const int c = it * G + gid;
            const bool active = c < tile_chirps;
            float2 v[E];
            RDP_UNROLL
            for (int i = 0; i < E; ++i) {
                const float2 s = Raw::cvt(raw[i]);
                if (IN_FMT == 0) peak = fmaxf(peak, fmaxf(fabsf(s.x), fabsf(s.y)));
                v[i] = make_float2(s.x * wv[i], s.y * wv[i]);
            }
            if (it + 1 < iters) fetch(t, it + 1);
            else fetch(t + (int)gridDim.x, 0);
            Plan::template run<T>(v, tw, xbuf, parity, tid, gid);
#ifndef RDP_EMU
            if (use_tma_store) {
                if (it == 0 && store_pending) {            // the box is free once the previous tile's stores have READ it
                    if (threadIdx.x == 0) tma_store_wait_read();
                    __syncthreads();
                }
                if (active) {
                    // dense 128-byte rows, 16-byte chunk index XOR (row & 7): CU_TENSOR_MAP_SWIZZLE_128B
                    RDP_UNROLL
                    for (int b = 0; b < PL::NB; ++b) {
                        RDP_UNROLL
                        for (int r = 0; r < PL::R; ++r) {
                            const int k = PL::out_index(tid, b, r);
                            *reinterpret_cast<float2*>(tile_bytes + k * 128 + ((((c >> 1) ^ k) & 7) << 4) + ((c & 1) << 3)) = v[b * PL::R + r];
                        }
                    }
                }
            } else
#endif
            if (active) {
                RDP_UNROLL
                for (int b = 0; b < PL::NB; ++b) {
                    RDP_UNROLL
                    for (int r = 0; r < PL::R; ++r) tile[PL::out_index(tid, b, r) * tcp + c] = v[b * PL::R + r];
                }
            }
        }
        if (IN_FMT == 0 && cpi_flags != nullptr) {
            if (__any_sync(0xffffffffu, peak >= 32767.0f) && (threadIdx.x & 31) == 0) atomicOr(cpi_flags + mp / n_channels, 4u);
        }
#ifndef RDP_EMU
        if (use_tma_store) {
            tma_store_fence();                             // generic-proxy writes of the box -> visible to the TMA unit
            __syncthreads();
            if (threadIdx.x == 0) {
                constexpr int BR = K1Shape<Plan>::BOX_ROWS;
                RDP_UNROLL
                for (int y = 0; y < N; y += BR) tma_store_3d(&tmap_out, tile_bytes + y * 128, 2 * c0, y, mp);
                tma_store_commit();
            }
            store_pending = true;
            continue;                                      // no store loop, no barrier: on to the next tile's transforms
        }
#endif
        __syncthreads();
        // range-major store: row k of the tile -> x_rc[map][k][c0 .. c0+tile_chirps)
        float2* dst = x_rc + map * (long long)N * n_chirps + c0;
        if (sc.world > 1) {
            // range line k belongs to rank k >> rows_shift; this rank's chirps sit at column rank * n_chirps
            const int rows_local = 1 << sc.rows_shift;
            const long long col0 = (long long)sc.rank * n_chirps + c0;
            if (tile_chirps == TCM && (TCM % 2) == 0 && (n_chirps % 2) == 0) {
                // two adjacent chirps per 128-bit store, as in the local path
                constexpr int HALF = TCM / 2, KSTEP = THREADS / HALF;
                const int c = (threadIdx.x % HALF) * 2;
                for (int k = threadIdx.x / HALF; k < N; k += KSTEP) {
                    float2* row = sc.peer[k >> sc.rows_shift] +
                                  (map * rows_local + (k & (rows_local - 1))) * (long long)sc.n_chirps_total + col0 + c;
                    const float2 a = tile[k * (TCM + 1) + c], b2 = tile[k * (TCM + 1) + c + 1];
                    *reinterpret_cast<float4*>(row) = make_float4(a.x, a.y, b2.x, b2.y);
                }
            } else {
                const int total = N * tile_chirps;
                for (int i = threadIdx.x; i < total; i += THREADS) {
                    const int k = i / tile_chirps, c = i % tile_chirps;
                    float2* base = sc.peer[k >> sc.rows_shift];
                    base[(map * rows_local + (k & (rows_local - 1))) * (long long)sc.n_chirps_total + col0 + c] = tile[k * tcp + c];
                }
            }
        } else if (tile_chirps == TCM && (TCM % 2) == 0 && (n_chirps % 2) == 0) {
            // compile-time tile width: each thread moves two adjacent chirps of one range bin per step with one
            // 128-bit store (rows of the output are 16-byte aligned: c0, n_chirps even); pointers just advance
            constexpr int HALF = TCM / 2, KSTEP = THREADS / HALF;
            const int c = (threadIdx.x % HALF) * 2;
            int k = threadIdx.x / HALF;
            const float2* src = tile + k * (TCM + 1) + c;
            float4* out = reinterpret_cast<float4*>(dst + (long long)k * n_chirps + c);
            const long long ostep = (long long)KSTEP * n_chirps / 2;
            for (; k < N; k += KSTEP) {
                const float2 a = src[0], b2 = src[1];
                *out = make_float4(a.x, a.y, b2.x, b2.y);
                src += KSTEP * (TCM + 1);
                out += ostep;
            }
        } else {
            const int total = N * tile_chirps;
            for (int i = threadIdx.x; i < total; i += THREADS) {
                const int k = i / tile_chirps, c = i % tile_chirps;
                dst[(long long)k * n_chirps + c] = tile[k * tcp + c];
            }
        }
        __syncthreads();
    }
#ifndef RDP_EMU
    if (store_pending && threadIdx.x == 0) tma_store_wait_read();   // shared memory must outlive the last store's reads
#endif
}

#ifndef RDP_EMU
// Cluster variant (CS = K1Shape::CS > 1).  Launch with cluster dimension (CS, 1, 1); tiles_per_map (tiles of TC_MAX
// chirps) must be a multiple of CS.  total_tiles counts TC_MAX-chirp tiles as in the plain kernel.
template <class Plan, int IN_FMT>
__global__ void __launch_bounds__(K1Shape<Plan>::THREADS)
k1_range_cluster_kernel(const void* __restrict__ cube, float2* __restrict__ x_rc,
                        const float* __restrict__ window, const float2* __restrict__ twtab,
                        int n_chirps, int tiles_per_map, long long total_tiles, K1Scatter sc) {
    using Sh = K1Shape<Plan>;
    constexpr int N = Plan::N, E = Plan::E, T = Plan::T;
    constexpr int G = Sh::G, THREADS = Sh::THREADS, TCM = Sh::TC_MAX, CS = Sh::CS;
    constexpr int ROWS = Sh::CL_ROWS, COLS = Sh::CL_COLS, PITCH = Sh::CL_PITCH;
    static_assert(CS > 1 && (COLS % 2) == 0, "cluster variant");
    using P0 = typename Plan::P0;
    using PL = typename Plan::LAST;
    using Raw = K1Raw<IN_FMT>;
    RDP_DYN_SMEM(smem_raw);
    float2* tile = reinterpret_cast<float2*>(smem_raw);                      // [ROWS][PITCH]: my range bins, the cluster's chirps
    float2* xbuf = tile + (size_t)ROWS * PITCH + (size_t)(threadIdx.x / T) * Plan::XBUFS;
    const int tid = threadIdx.x % T, gid = threadIdx.x / T;
    const int rank = (int)k1_cluster_rank();
    const int cid = blockIdx.x / CS, n_clusters = gridDim.x / CS;
    // the same tile in every CTA of the cluster (shared::cluster window addresses)
    uint32_t peer_tile[CS];
    RDP_UNROLL
    for (int r = 0; r < CS; ++r) peer_tile[r] = k1_mapa((uint32_t)__cvta_generic_to_shared(tile), (uint32_t)r);

    float2 tw[Plan::NTW > 0 ? Plan::NTW : 1];
    Plan::load_twiddles(tw, twtab, tid);
    float wv[E];
    RDP_UNROLL
    for (int b = 0; b < P0::NB; ++b) {
        RDP_UNROLL
        for (int r = 0; r < P0::R; ++r) wv[b * P0::R + r] = __ldg(window + P0::in_index(tid, b, r));
    }
    int parity = 0;
    constexpr int ITERS = (TCM + G - 1) / G;
    const int n_ctiles = (int)(total_tiles / CS);                          // cluster tiles of CS * TCM chirps
    const int ctiles_per_map = tiles_per_map / CS;

    typename Raw::T raw[E];
    auto fetch = [&](int ct, int it) {
        const int c = it * G + gid;
        const bool ok = ct < n_ctiles && c < TCM;
        const int mp = ct / ctiles_per_map;
        const int chirp = mp * n_chirps + (ct - mp * ctiles_per_map) * COLS + rank * TCM + c;
        const long long first = (long long)chirp * N + tid;
        RDP_UNROLL
        for (int b = 0; b < P0::NB; ++b) {
            RDP_UNROLL
            for (int r = 0; r < P0::R; ++r)
                raw[b * P0::R + r] = ok ? Raw::load(cube, first + (b * T + r * (N / P0::R))) : Raw::zero();
        }
    };
    fetch(cid, 0);
    k1_cluster_sync();                                                     // every CTA of the cluster is running

    for (int ct = cid; ct < n_ctiles; ct += n_clusters) {
        const int mp = ct / ctiles_per_map;
        const long long map = mp;
        const int c0 = (ct - mp * ctiles_per_map) * COLS;                   // first chirp of the cluster tile
        for (int it = 0; it < ITERS; ++it) {
            const int c = it * G + gid;
            const bool active = c < TCM;
            float2 v[E];
            RDP_UNROLL
            for (int i = 0; i < E; ++i) {
                const float2 s = Raw::cvt(raw[i]);
                v[i] = make_float2(s.x * wv[i], s.y * wv[i]);
            }
            if (it + 1 < ITERS) fetch(ct, it + 1);
            else fetch(ct + n_clusters, 0);
            Plan::template run<T>(v, tw, xbuf, parity, tid, gid);
            if (active) {
                const uint32_t col_off = (uint32_t)(rank * TCM + c) * 8u;
                RDP_UNROLL
                for (int b = 0; b < PL::NB; ++b) {
                    RDP_UNROLL
                    for (int r = 0; r < PL::R; ++r) {
                        const int k = PL::out_index(tid, b, r);            // range bin: owner CTA k / ROWS, its row k % ROWS
                        const int owner = k / ROWS;
                        uint32_t base = peer_tile[0];
                        RDP_UNROLL
                        for (int q = 1; q < CS; ++q) base = owner == q ? peer_tile[q] : base;
                        k1_st_cluster(base + (uint32_t)(k % ROWS) * (PITCH * 8u) + col_off, v[b * PL::R + r]);
                    }
                }
            }
        }
        k1_cluster_sync();                                                 // all spectra of the cluster tile have landed
        {
            constexpr int HALF = COLS / 2, KSTEP = THREADS / HALF;
            const int c = (threadIdx.x % HALF) * 2;
            if (sc.world > 1) {
                const int rows_local = 1 << sc.rows_shift;
                const long long col0 = (long long)sc.rank * n_chirps + c0 + c;
                for (int r = threadIdx.x / HALF; r < ROWS; r += KSTEP) {
                    const int k = rank * ROWS + r;
                    float2* row = sc.peer[k >> sc.rows_shift] +
                                  (map * rows_local + (k & (rows_local - 1))) * (long long)sc.n_chirps_total + col0;
                    const float2 a = tile[r * PITCH + c], b2 = tile[r * PITCH + c + 1];
                    *reinterpret_cast<float4*>(row) = make_float4(a.x, a.y, b2.x, b2.y);
                }
            } else {
                float2* dst = x_rc + map * (long long)N * n_chirps + (long long)rank * ROWS * n_chirps + c0 + c;
                for (int r = threadIdx.x / HALF; r < ROWS; r += KSTEP) {
                    const float2 a = tile[r * PITCH + c], b2 = tile[r * PITCH + c + 1];
                    *reinterpret_cast<float4*>(dst + (long long)r * n_chirps) = make_float4(a.x, a.y, b2.x, b2.y);
                }
            }
        }
        k1_cluster_sync();                                                 // nobody overwrites a tile that is still being stored
    }
}
#endif

}  // namespace rdp
