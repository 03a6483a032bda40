// rdp_api.cu -- the C ABI of librdp (include/rdp.h) and the host-side launch logic.
//
// Built two ways from the same source:
//   nvcc -gencode arch=compute_100a,code=sm_100a ...      -> librdp.so      (the product)
//   g++  -x c++ -DRDP_EMU ...                              -> librdp_emu.so  (CPU test harness
//        for kernel logic only; see tests/emu/cuda_emu.h -- never loaded by the product)
#include "../../include/rdp.h"
#include "k1_range.cuh"
#include "k2_doppler.cuh"
#include "k3_cfar.cuh"
#include "k0_fixed.cuh"
#include "k6_tracker.cuh"

#ifndef RDP_EMU
#include <cuda.h>          // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved through the runtime
#endif
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace rdp;

namespace {

std::string g_create_error;

struct StageEvent { int stage; cudaEvent_t a, b; };

struct K1Launcher {
    int T, G, tc_max;
    size_t (*smem)(int tile_chirps);
    cudaError_t (*prepare)(size_t smem, int* ctas_per_sm);
    void (*launch)(int grid, size_t smem, cudaStream_t st, const void* cube, float2* x_rc, const float* win,
                   const float2* tw, int n_chirps, int tile_chirps, int tiles_per_map, long long total_tiles, K1Scatter sc,
                   int use_tma_store, const CUtensorMap* tmap, uint32_t* cpi_flags, int n_channels);
    int box_rows;                  // range lines per TMA store box
    // thread-block-cluster variant for long chirps (cs = CTAs per cluster; 1 = none)
    int cs;
    size_t smem_cluster;
    cudaError_t (*prepare_cluster)(size_t smem, int* max_clusters);
    cudaError_t (*launch_cluster)(int n_clusters, size_t smem, cudaStream_t st, const void* cube, float2* x_rc,
                                  const float* win, const float2* tw, int n_chirps, int tiles_per_map,
                                  long long total_tiles, K1Scatter sc);
};
constexpr int kStreamSpecDets = 24576;      // per CPI: head of the detection list a streaming slot copies back unasked

struct K2Launcher {
    int T, G;
    size_t smem;
    cudaError_t (*prepare)(size_t smem, int* ctas_per_sm);
    void (*launch)(int grid, size_t smem, cudaStream_t st, const float2* x_rc, float* rd_map, const float* win,
                   const float2* tw, long long n_rows, int mti_mode, int rows_per_map, int pad_rows, int combine, int n_combine,
                   int n_range, int n_beams, const float2* steer);
};
struct K3Launcher {
    int trows, tcols;            // 0,0 => generic kernel
    int srows, scols;            // shared-memory tile (+ halo) = TMA box
    size_t smem;
    cudaError_t (*prepare)(size_t smem, int* ctas_per_sm);
    void (*launch)(long long grid, size_t smem, cudaStream_t st, const float* rd_map, float* cfar_out,
                   uint32_t* det_bits, CfarArgs a, int tiles_r, int tiles_c, long long n_tiles, int use_tma,
                   const CUtensorMap* tmap, uint32_t* cpi_flags, int* tile_queue);
};

template <class K>
cudaError_t prepare_kernel(K kernel, int threads, size_t smem, int* ctas_per_sm) {
    cudaError_t e = cudaSuccess;
#ifndef RDP_EMU
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int n = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
    if (ctas_per_sm) *ctas_per_sm = n < 1 ? 1 : n;
#else
    (void)kernel; (void)threads; (void)smem;
    if (ctas_per_sm) *ctas_per_sm = 1;
#endif
    return e;
}

// ---- K1 dispatch ---------------------------------------------------------------------------
template <class Plan, int FMT>
struct K1Inst {
    static size_t smem(int tc) { return K1Shape<Plan>::smem_bytes(tc); }
    static cudaError_t prepare(size_t s, int* occ) {
        return prepare_kernel(k1_range_kernel<Plan, FMT>, K1Shape<Plan>::THREADS, s, occ);
    }
    static void launch(int grid, size_t s, cudaStream_t st, const void* cube, float2* x_rc, const float* win,
                       const float2* tw, int n_chirps, int tile_chirps, int tiles_per_map, long long total_tiles, K1Scatter sc,
                       int use_tma_store, const CUtensorMap* tmap, uint32_t* cpi_flags, int n_channels) {
        RDP_LAUNCH((k1_range_kernel<Plan, FMT>), grid, K1Shape<Plan>::THREADS, s, st, cube, x_rc, win, tw,
                   n_chirps, tile_chirps, tiles_per_map, total_tiles, sc, use_tma_store, *tmap, cpi_flags, n_channels);
    }
#ifndef RDP_EMU
    static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int n_clusters, size_t s, cudaStream_t st) {
        std::memset(cfg, 0, sizeof(*cfg));
        cfg->gridDim = dim3((unsigned)(n_clusters * K1Shape<Plan>::CS));
        cfg->blockDim = dim3(K1Shape<Plan>::THREADS);
        cfg->dynamicSmemBytes = s;
        cfg->stream = st;
        attr->id = cudaLaunchAttributeClusterDimension;
        attr->val.clusterDim.x = K1Shape<Plan>::CS;
        attr->val.clusterDim.y = 1;
        attr->val.clusterDim.z = 1;
        cfg->attrs = attr;
        cfg->numAttrs = 1;
    }
    static cudaError_t prepare_cluster(size_t s, int* max_clusters) {
        if constexpr (K1Shape<Plan>::CS > 1) {
            cudaError_t e = cudaFuncSetAttribute(k1_range_cluster_kernel<Plan, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s);
            if (e != cudaSuccess) return e;
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute attr;
            cluster_config(&cfg, &attr, 1, s, nullptr);
            int n = 0;
            e = cudaOccupancyMaxActiveClusters(&n, k1_range_cluster_kernel<Plan, FMT>, &cfg);
            *max_clusters = n;
            return e;
        } else {
            *max_clusters = 0;
            return cudaSuccess;
        }
    }
    static cudaError_t launch_cluster(int n_clusters, size_t s, cudaStream_t st, const void* cube, float2* x_rc,
                                      const float* win, const float2* tw, int n_chirps, int tiles_per_map,
                                      long long total_tiles, K1Scatter sc) {
        if constexpr (K1Shape<Plan>::CS > 1) {
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute attr;
            cluster_config(&cfg, &attr, n_clusters, s, st);
            return cudaLaunchKernelEx(&cfg, k1_range_cluster_kernel<Plan, FMT>, cube, x_rc, win, tw, n_chirps, tiles_per_map,
                                      total_tiles, sc);
        } else {
            return cudaErrorNotSupported;
        }
    }
    static K1Launcher get() {
        return K1Launcher{Plan::T, K1Shape<Plan>::G, K1Shape<Plan>::TC_MAX, &smem, &prepare, &launch, K1Shape<Plan>::BOX_ROWS,
                          K1Shape<Plan>::CS, K1Shape<Plan>::smem_bytes_cluster(), &prepare_cluster, &launch_cluster};
    }
#else
    static K1Launcher get() {
        return K1Launcher{Plan::T, K1Shape<Plan>::G, K1Shape<Plan>::TC_MAX, &smem, &prepare, &launch, K1Shape<Plan>::BOX_ROWS, 1, 0, nullptr, nullptr};
    }
#endif
};

template <class Plan, int MAG>
struct K2Inst {
    static cudaError_t prepare(size_t s, int* occ) {
        cudaError_t e = prepare_kernel(k2_doppler_kernel<Plan, MAG, 1>, K2Shape<Plan>::THREADS, s, nullptr);
        if (e != cudaSuccess) return e;
        e = prepare_kernel(k2_doppler_kernel<Plan, MAG, 2>, K2Shape<Plan>::THREADS, s, nullptr);
        if (e != cudaSuccess) return e;
        return prepare_kernel(k2_doppler_kernel<Plan, MAG, 0>, K2Shape<Plan>::THREADS, s, occ);
    }
    static void launch(int grid, size_t s, cudaStream_t st, const float2* x_rc, float* rd_map, const float* win,
                       const float2* tw, long long n_rows, int mti_mode, int rows_per_map, int pad_rows, int combine, int n_combine,
                       int n_range, int n_beams, const float2* steer) {
        // combine 1: non-coherent integration over n_combine channels, n_rows counts OUTPUT lines (cpi, range);
        // combine 2: coherent beams, n_rows counts (cpi, range, beam)
        if (combine == RDP_COMBINE_NCI)
            RDP_LAUNCH((k2_doppler_kernel<Plan, MAG, 1>), grid, K2Shape<Plan>::THREADS, s, st, x_rc, rd_map, win, tw,
                       n_rows, mti_mode, rows_per_map, pad_rows, n_combine, n_range, 1, steer);
        else if (combine == RDP_COMBINE_BEAMS)
            RDP_LAUNCH((k2_doppler_kernel<Plan, MAG, 2>), grid, K2Shape<Plan>::THREADS, s, st, x_rc, rd_map, win, tw,
                       n_rows, mti_mode, rows_per_map, pad_rows, n_combine, n_range, n_beams, steer);
        else
            RDP_LAUNCH((k2_doppler_kernel<Plan, MAG, 0>), grid, K2Shape<Plan>::THREADS, s, st, x_rc, rd_map, win, tw,
                       n_rows, mti_mode, rows_per_map, pad_rows, 1, n_range, 1, steer);
    }
    static K2Launcher get() {
        return K2Launcher{Plan::T, K2Shape<Plan>::G, K2Shape<Plan>::smem_bytes(), &prepare, &launch};
    }
};

// FFT plans per length (see fft_core.cuh): radix <= 8 with 8 points / thread up to 512,
// radix 16 with 16 points / thread (single-buffered exchange from 2048) above.
using Plan8 = FftPlan<8, 8, 8>;
using Plan16 = FftPlan<16, 8, 8, 2>;
using Plan32 = FftPlan<32, 8, 8, 4>;
using Plan64 = FftPlan<64, 8, 8, 8>;
using Plan128 = FftPlan<128, 8, 8, 4, 4>;
using Plan256 = FftPlan<256, 8, 8, 8, 4>;
using Plan512 = FftPlan<512, 8, 8, 8, 8>;
using Plan1024 = FftPlan<1024, 16, 16, 16, 4>;
using Plan2048 = FftPlan<2048, 16, 16, 16, 8, 1, 0>;
using Plan4096 = FftPlan<4096, 16, 16, 16, 16, 1, 0>;

template <int FMT>
bool k1_for(int n, K1Launcher* out) {
    switch (n) {
        case 8: *out = K1Inst<Plan8, FMT>::get(); return true;
        case 16: *out = K1Inst<Plan16, FMT>::get(); return true;
        case 32: *out = K1Inst<Plan32, FMT>::get(); return true;
        case 64: *out = K1Inst<Plan64, FMT>::get(); return true;
        case 128: *out = K1Inst<Plan128, FMT>::get(); return true;
        case 256: *out = K1Inst<Plan256, FMT>::get(); return true;
        case 512: *out = K1Inst<Plan512, FMT>::get(); return true;
        case 1024: *out = K1Inst<Plan1024, FMT>::get(); return true;
        case 2048: *out = K1Inst<Plan2048, FMT>::get(); return true;
        case 4096: *out = K1Inst<Plan4096, FMT>::get(); return true;
    }
    return false;
}
template <int MAG>
bool k2_for(int n, K2Launcher* out) {
    switch (n) {
        case 8: *out = K2Inst<Plan8, MAG>::get(); return true;
        case 16: *out = K2Inst<Plan16, MAG>::get(); return true;
        case 32: *out = K2Inst<Plan32, MAG>::get(); return true;
        case 64: *out = K2Inst<Plan64, MAG>::get(); return true;
        case 128: *out = K2Inst<Plan128, MAG>::get(); return true;
        case 256: *out = K2Inst<Plan256, MAG>::get(); return true;
        case 512: *out = K2Inst<Plan512, MAG>::get(); return true;
        case 1024: *out = K2Inst<Plan1024, MAG>::get(); return true;
        case 2048: *out = K2Inst<Plan2048, MAG>::get(); return true;
        case 4096: *out = K2Inst<Plan4096, MAG>::get(); return true;
    }
    return false;
}

// ---- K3 dispatch ---------------------------------------------------------------------------
template <int HL, int HC, int GL, int GC, int TR, int NSTAGE, int P0B = 1>
struct K3Inst {
    using Sh = K3Shape<HL, HC, GL, GC, TR, NSTAGE, P0B>;
    static cudaError_t prepare(size_t s, int* occ) {
        cudaError_t e = prepare_kernel(k3_cfar_kernel<HL, HC, GL, GC, TR, NSTAGE, P0B, true>, Sh::THREADS, s, nullptr);
        if (e != cudaSuccess) return e;
        return prepare_kernel(k3_cfar_kernel<HL, HC, GL, GC, TR, NSTAGE, P0B, false>, Sh::THREADS, s, occ);
    }
    static void launch(long long grid, size_t s, cudaStream_t st, const float* rd_map, float* cfar_out,
                       uint32_t* det_bits, CfarArgs a, int tiles_r, int tiles_c, long long n_tiles, int use_tma,
                       const CUtensorMap* tmap, uint32_t* cpi_flags, int* tile_queue) {
        if (a.int_mode)
            RDP_LAUNCH((k3_cfar_kernel<HL, HC, GL, GC, TR, NSTAGE, P0B, true>), (unsigned)grid, Sh::THREADS, s, st, rd_map, cfar_out,
                       det_bits, a, tiles_r, tiles_c, n_tiles, use_tma, *tmap, cpi_flags, tile_queue);
        else
            RDP_LAUNCH((k3_cfar_kernel<HL, HC, GL, GC, TR, NSTAGE, P0B, false>), (unsigned)grid, Sh::THREADS, s, st, rd_map, cfar_out,
                       det_bits, a, tiles_r, tiles_c, n_tiles, use_tma, *tmap, cpi_flags, tile_queue);
    }
    static K3Launcher get() {
        return K3Launcher{TR, Sh::TC, Sh::SR, Sh::SC, Sh::smem_bytes(), &prepare, &launch};
    }
};
struct K3Generic {
    static cudaError_t prepare(size_t, int* occ) { if (occ) *occ = 64; return cudaSuccess; }
    static void launch(long long grid, size_t, cudaStream_t st, const float* rd_map, float* cfar_out,
                       uint32_t* det_bits, CfarArgs a, int, int, long long, int cap, const CUtensorMap*, uint32_t*, int*) {
        long long n_words = grid;   // grid carries the bitmask word count for the generic kernel, `use_tma` its CTA cap
        long long blocks = (n_words + 7) / 8;
        if (blocks > cap) blocks = cap;
        RDP_LAUNCH(k3_cfar_generic_kernel, (unsigned)blocks, 256, 0, st, rd_map, cfar_out, det_bits, a, n_words);
    }
    static K3Launcher get() { return K3Launcher{0, 0, 0, 0, 0, &prepare, &launch}; }
};

// `variant` (RDP_K3_VARIANT, A/B measurements of the default window only): 0 = 64-line tiles, two raw stages (the
// default since the phase-0 rewrite: three CTAs per SM with 85 registers and the next tile's TMA load under the
// current tile beat four CTAs with 64 registers, 9.22 against 9.56 ms per 1024 automotive CPIs); 1 = 64 lines, one
// stage; 2 = 32 lines, one stage; 3 = 32 lines, two stages; 4 / 5 = 64 lines, two / one stage(s), phase 0 with two
// bands per thread on half of the warps
K3Launcher k3_for(int hl, int hc, int gl, int gc, int variant) {
#define RDP_K3_CASE(HL, HC, GL, GC, TR, NS) \
    if (hl == HL && hc == HC && gl == GL && gc == GC) return K3Inst<HL, HC, GL, GC, TR, NS>::get();
    if (variant == 1) { RDP_K3_CASE(5, 6, 1, 2, 64, 1) }
    if (variant == 2) { RDP_K3_CASE(5, 6, 1, 2, 32, 1) }
    if (variant == 3) { RDP_K3_CASE(5, 6, 1, 2, 32, 2) }
    if (variant == 4 && hl == 5 && hc == 6 && gl == 1 && gc == 2) return K3Inst<5, 6, 1, 2, 64, 2, 2>::get();
    if (variant == 5 && hl == 5 && hc == 6 && gl == 1 && gc == 2) return K3Inst<5, 6, 1, 2, 64, 1, 2>::get();
    RDP_K3_CASE(5, 6, 1, 2, 64, 2)      // radar_core.vhd:376-382 defaults (4,4,2,1), as built
    RDP_K3_CASE(6, 5, 2, 1, 64, 2)      // same, as named
    RDP_K3_CASE(3, 4, 1, 1, 64, 1)      // tb_os_cfar_2d.vhd:16-19 (3,2,1,1), as built
    RDP_K3_CASE(4, 3, 1, 1, 64, 1)      // same, as named
    RDP_K3_CASE(3, 3, 1, 1, 64, 1)      // legacy quick config (2,2,1,1)
    RDP_K3_CASE(10, 12, 2, 4, 64, 1)    // BASELINE config 4 "wide" (8,8,4,2), as built
    RDP_K3_CASE(12, 10, 4, 2, 64, 1)    // same, as named
#undef RDP_K3_CASE
    return K3Generic::get();
}

}  // namespace

// One in-flight submission of the streaming ingest path (BASELINE config 5)
struct StreamSlot {
    cudaStream_t copy = nullptr;            // H2D of the cube and D2H of the results
    cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
    long long spec = 0;                     // records copied back speculatively with the counts
    void* d_in = nullptr;
    rdp_det* d_dets = nullptr;
    rdp_det* d_packed = nullptr;
    int* d_counts = nullptr;
    uint32_t* d_status = nullptr;
    rdp_det* h_packed = nullptr;            // page-locked results
    int* h_counts = nullptr;
    uint32_t* h_status = nullptr;
    int n_cpi = 0;
    bool busy = false;
#ifndef RDP_EMU
    cudaGraphExec_t graph = nullptr;        // the slot's kernel chain (K1 .. K5) captured once per n_cpi, replayed per submission
    int graph_n_cpi = 0;
#endif
};

struct rdp_handle {
    rdp_config cfg;
    int device = 0;
    int sm_count = 0;
    int chunk = 1;
    int host_chunk = 1;            // CPIs per pass of rdp_process_host (<= chunk)
    cudaStream_t stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
    cudaStream_t cfar_stream = nullptr;    // overlap mode: K3/K4 of chunk i run here while K1/K2 of chunk i+1 run on `stream`
    cudaStream_t cur = nullptr;            // stream the run_* helpers launch on
    cudaEvent_t ev_map_ready[2] = {nullptr, nullptr}, ev_cfar_done[2] = {nullptr, nullptr};
    int overlap = 0;               // RDP_OVERLAP: with the K3 tile queue the two-stream pipeline costs 1.7 % (measured), so it is opt-in
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::vector<cudaEvent_t> ev_copied, ev_done, ev_mapout, ev_counts;
    float *win_r = nullptr, *win_d = nullptr;
    float2 *tw_r = nullptr, *tw_d = nullptr;
    float2* x_rc = nullptr;        // [chunk][channel][range][chirp]
    float* map_scratch[2] = {nullptr, nullptr};
    void* in_stage[2] = {nullptr, nullptr};
    size_t cfar_rows_cap = 0;      // range lines the bitmask / row-offset scratch can hold
    uint32_t* det_bits = nullptr;  // [chunk][channel][range][words_per_row] detection bitmask
    int* row_off = nullptr;        // [chunk][channel][range] record offset of each range line within its CPI
    int* counts = nullptr;         // [max_cpis]
    uint32_t* status = nullptr;    // [max_cpis]
    rdp_det* dets = nullptr;       // [max_cpis][max_dets]   (host path)
    rdp_det* packed = nullptr;     // packed list            (host path)
    void* l2_flush = nullptr;
    size_t l2_flush_bytes = 0;
    int64_t scratch_bytes = 0;
    int64_t launches = 0;
    K1Launcher k1;
    K2Launcher k2;
    K3Launcher k3;
    int k1_grid_cap = 1, k2_grid_cap = 1, k3_grid_cap = 1;
    int knob_k3_ctas = 0, knob_k12_ctas = 0, knob_k3_tiles_per_cta = 0, knob_priority = 1;
    int knob_k3_debug_skip = 0, knob_k3_skew_ns = 0;
    int knob_k1_cluster = -1;       // RDP_K1_CLUSTER: -1 = automatic (sharded mode only), 0 / 1 = forced off / on
    int stream_spec_dets = kStreamSpecDets;
    int knob_k3_queue = 1;          // RDP_K3_QUEUE: K3's persistent CTAs draw their tiles from a device-side queue
    int* k3_queue = nullptr;
    int knob_stream_graph = 1;      // RDP_STREAM_GRAPH: streaming submissions replay a CUDA graph of their kernel chain
    int k3_variant = 0;             // RDP_K3_VARIANT (tile shape / staging depth of the default CFAR window, A/B measurements)
    void* encode_tiled = nullptr;   // cuTensorMapEncodeTiled, resolved at create (NULL: manual tile loads)
    int cfar_tma = 0;               // last CFAR launch staged its tiles with TMA
    int range_tma = 0;              // last range-stage launch stored its tiles with TMA
    int knob_k1_tma = -1;           // RDP_K1_TMA: -1 = automatic (int16 input: measured equal or better; complex64: 4 % worse), 0 / 1 = forced
    int k1_tile_chirps = 1;
    size_t k1_smem = 0;
    int k1_max_clusters = 0;       // > 0: the cluster variant of K1 is usable for this shape
    CfarArgs cfar;
    TrackState* trk_state = nullptr;   // one track file per channel
    TrackerArgs trk_args;
    int* rom_dev = nullptr;         // mirrored Q15 Hamming coefficients of the last fixed-point window length (tap)
    int rom_n = 0;
    uint32_t* sat_dev = nullptr;
    int map_channels = 1;          // maps per CPI the CFAR stage sees: n_channels, 1 with combine = NCI, n_beams with BEAMS
    int n_beams = 1;               // combine = BEAMS: smallest power of two >= n_channels
    float2* steer = nullptr;       // exp(-j 2 pi m / n_beams), m < n_beams
    uint32_t* det_bits2 = nullptr; // peak_group: the grouped bitmask (same shape as det_bits)
    size_t det_bits2_cap = 0;
    uint32_t* cpi_flags[2] = {nullptr, nullptr};   // [chunk] status bits raised before the CFAR stage, by chunk parity
    // arith = FIXED16: ROMs and Q15 twiddles of both transforms, block exponents of the current chunk
    int *rom_r = nullptr, *rom_d = nullptr;
    uint32_t *btw_r = nullptr, *btw_d = nullptr;
    int *exp_r = nullptr, *exp_d = nullptr;
    uint32_t* btw_tap = nullptr;    // twiddles of the last rdp_fixed_bfp_fft length
    int btw_tap_n = 0;
    std::vector<StreamSlot> slots;
    int slot_cpis = 0;
    bool stage_timing = false;
    std::vector<StageEvent> stage_events;
    float stage_ms[4] = {0, 0, 0, 0};
    int64_t stage_launches[4] = {0, 0, 0, 0};
    std::string err;
};

namespace {

int fail(rdp_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define RDP_CUDA(h, expr)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(h, e_ == cudaErrorMemoryAllocation ? RDP_E_NOMEM : RDP_E_CUDA,        \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

size_t in_bytes_per_sample(const rdp_config& c) { return c.input_format == RDP_IN_INT16_IQ ? 4 : 8; }
size_t cells_per_cpi(const rdp_config& c) { return (size_t)c.n_channels * c.n_chirps * c.n_samples; }
// cells of the range-Doppler map(s) of one CPI: one map per channel, or a single combined map (combine = NCI)
int beams_for(int n_channels) { int b = 1; while (b < n_channels) b <<= 1; return b; }
int map_channels_of(const rdp_config& c) {
    return c.combine == RDP_COMBINE_NCI ? 1 : (c.combine == RDP_COMBINE_BEAMS ? beams_for(c.n_channels) : c.n_channels);
}
size_t map_cells_per_cpi(const rdp_config& c) { return (size_t)map_channels_of(c) * c.n_chirps * c.n_samples; }

template <class T>
int dev_alloc(rdp_handle* h, T** p, size_t count) {
    void* v = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    RDP_CUDA(h, cudaMalloc(&v, bytes));
    *p = reinterpret_cast<T*>(v);
    h->scratch_bytes += (int64_t)bytes;
    return RDP_OK;
}

void stage_begin(rdp_handle* h, int stage) {
    h->launches += 1;
    h->stage_launches[stage] += 1;
    if (!h->stage_timing) return;
    StageEvent se;
    se.stage = stage;
    cudaEventCreate(&se.a);
    cudaEventCreate(&se.b);
    cudaEventRecord(se.a, h->cur);
    h->stage_events.push_back(se);
}
void stage_end(rdp_handle* h) {
    if (!h->stage_timing) return;
    cudaEventRecord(h->stage_events.back().b, h->cur);
}

int launch_check(rdp_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, RDP_E_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return RDP_OK;
}

// ---- the three stages on device buffers (n_cpi <= any; no chunking here) -------------------
int run_range(rdp_handle* h, const void* cube, int n_cpi, float2* x_rc, const K1Scatter* scatter = nullptr,
              uint32_t* cpi_flags = nullptr) {
    const rdp_config& c = h->cfg;
    K1Scatter sc;
    std::memset(&sc, 0, sizeof(sc));
    sc.world = 1;
    if (scatter) sc = *scatter;
    const int tc = h->k1_tile_chirps;
    const int tiles_per_map = c.n_chirps / tc;
    const long long total = (long long)n_cpi * c.n_channels * tiles_per_map;
    if (total == 0) return RDP_OK;
    if (total >= (1ll << 31) || (long long)n_cpi * c.n_channels * c.n_chirps >= (1ll << 31))
        return fail(h, RDP_E_CAPACITY, "range stage: too many chirps in one launch; lower chunk_cpis");
    int grid = (int)(total < h->k1_grid_cap ? total : h->k1_grid_cap);
    stage_begin(h, 0);
    // The cluster variant pays for its 128-byte segments with 8-byte distributed-shared-memory stores: measured 2.4x
    // slower than the plain kernel into local HBM (1024 x 4096: 45 vs 19 ms per 1024 CPIs), 15 % faster when the
    // segments are peer stores over NVLink (profiles/r1_summary.md) -- so it serves the sharded mode only.
    // RDP_K1_CLUSTER=0 / 1 (read at create) forces it off / on (A/B measurements).
    bool use_cluster = h->k1_max_clusters > 0 && sc.world > 1;
    if (h->knob_k1_cluster >= 0) use_cluster = h->k1_max_clusters > 0 && h->knob_k1_cluster != 0;
    if (use_cluster) {
        const long long ctiles = total / h->k1.cs;
        const int n_clusters = (int)(ctiles < h->k1_max_clusters ? ctiles : h->k1_max_clusters);
        RDP_CUDA(h, h->k1.launch_cluster(n_clusters, h->k1.smem_cluster, h->cur, cube, x_rc, h->win_r, h->tw_r, c.n_chirps,
                                         tiles_per_map, total, sc));
    } else {
        CUtensorMap tmap;
        std::memset(&tmap, 0, sizeof(tmap));
        int use_tma_store = 0;
#ifndef RDP_EMU
        // 16-chirp tiles leave as TMA stores: x_rc seen as [map][range][2 * n_chirps floats], box = 128 bytes x box_rows
        // lines, 128-byte swizzle (k1_range.cuh).  Peer stores of the sharded mode and narrow tiles keep the store loop.
        const bool want_tma = h->knob_k1_tma < 0 ? c.input_format == RDP_IN_INT16_IQ : h->knob_k1_tma != 0;
        if (h->encode_tiled && want_tma && sc.world == 1 && tc == 16 && ((uintptr_t)x_rc & 15) == 0) {
            typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            const cuuint64_t maps = (cuuint64_t)n_cpi * c.n_channels;
            const cuuint64_t gdim[3] = {(cuuint64_t)c.n_chirps * 2, (cuuint64_t)c.n_samples, maps};
            const cuuint64_t gstr[2] = {(cuuint64_t)c.n_chirps * 8, (cuuint64_t)c.n_chirps * c.n_samples * 8};
            const cuuint32_t box[3] = {32, (cuuint32_t)h->k1.box_rows, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult rr = reinterpret_cast<encode_fn>(h->encode_tiled)(
                &tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x_rc, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            use_tma_store = (rr == CUDA_SUCCESS);
        }
#endif
        h->range_tma = use_tma_store;
        h->k1.launch(grid, h->k1_smem, h->cur, cube, x_rc, h->win_r, h->tw_r, c.n_chirps, tc, tiles_per_map, total, sc,
                     use_tma_store, &tmap, cpi_flags, c.n_channels);
    }
    stage_end(h);
    return launch_check(h, "k1_range_kernel");
}

int run_doppler(rdp_handle* h, const float2* x_rc, int n_cpi, float* rd_map, long long rows_override = -1,
                int rows_per_map = 0, int pad_rows = 0) {
    const rdp_config& c = h->cfg;
    const int combine = rows_override < 0 ? c.combine : RDP_COMBINE_NONE;
    // work items: output lines; with coherent beams one item yields the two beams of a pair
    const int items_per_line = (combine == RDP_COMBINE_BEAMS && h->n_beams > 1) ? h->n_beams / 2 : h->map_channels;
    const long long rows = rows_override >= 0 ? rows_override : (long long)n_cpi * items_per_line * c.n_samples;
    if (rows == 0) return RDP_OK;
    long long blocks = (rows + h->k2.G - 1) / h->k2.G;
    int grid = (int)(blocks < h->k2_grid_cap ? blocks : h->k2_grid_cap);
    stage_begin(h, 1);
    h->k2.launch(grid, h->k2.smem, h->cur, x_rc, rd_map, h->win_d, h->tw_d, rows, c.mti_mode, rows_per_map, pad_rows,
                 combine, c.n_channels, c.n_samples, h->n_beams, h->steer);
    stage_end(h);
    return launch_check(h, "k2_doppler_kernel");
}

// K3 + K4 on n_cpi <= chunk maps-worth of CPIs: bitmask, then ordered compaction into dets / counts / status
// (all three already offset to this chunk); cpi_label is the CPI index written into the records.
int run_cfar(rdp_handle* h, const float* rd_map, int n_cpi, float* cfar_out, rdp_det* dets, int* counts,
             uint32_t* status, int cpi_label, int int_mode = 0, int n_range_rows = 0, uint32_t* cpi_flags = nullptr) {
    rdp_config c = h->cfg;                                 // by value: n_samples may be overridden (sharded maps)
    c.n_channels = h->map_channels;                        // maps per CPI (1 after channel combining)
    CfarArgs args = h->cfar;
    args.int_mode = int_mode;
    if (n_range_rows > 0) {
        c.n_samples = n_range_rows;
        args.n_range = n_range_rows;
        // bitmask / row-offset scratch was sized for the configured map height; grow it if this one is taller
        const size_t need = (size_t)n_cpi * c.n_channels * n_range_rows;
        if (need > h->cfar_rows_cap) {
            h->scratch_bytes -= (int64_t)(h->cfar_rows_cap * ((size_t)args.words_per_row * sizeof(uint32_t) + sizeof(int)));
            if (h->det_bits) cudaFree(h->det_bits);
            if (h->row_off) cudaFree(h->row_off);
            h->det_bits = nullptr; h->row_off = nullptr;
            int rc = dev_alloc(h, &h->det_bits, need * args.words_per_row);
            if (rc) return rc;
            rc = dev_alloc(h, &h->row_off, need);
            if (rc) return rc;
            h->cfar_rows_cap = need;
        }
    }
    const long long maps = (long long)n_cpi * c.n_channels;
    if (maps == 0) return RDP_OK;
    if (maps * c.n_samples >= (1ll << 31)) return fail(h, RDP_E_CAPACITY, "CFAR stage: too many range lines in one launch; lower chunk_cpis");
    if (h->k3.trows != 0) {
        const long long tiles = maps * ((c.n_samples + h->k3.trows - 1) / h->k3.trows) * ((c.n_chirps + h->k3.tcols - 1) / h->k3.tcols);
        if (tiles >= (1ll << 31)) return fail(h, RDP_E_CAPACITY, "CFAR stage: too many tiles in one launch; lower chunk_cpis");
    }
    stage_begin(h, 2);
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    if (h->k3.trows == 0) {
        h->k3.launch(maps * c.n_samples * args.words_per_row, 0, h->cur, rd_map, cfar_out, h->det_bits, args, 0, 0,
                     0, h->k3_grid_cap, &tmap, nullptr, nullptr);
    } else {
        const int tiles_r = (c.n_samples + h->k3.trows - 1) / h->k3.trows;
        const int tiles_c = (c.n_chirps + h->k3.tcols - 1) / h->k3.tcols;
        const long long n_tiles = maps * tiles_r * tiles_c;
        int use_tma = 0;
#ifndef RDP_EMU
        // TMA stages the (tile + halo) box and zero-fills outside the map == cfar_edges ZERO; the flat-raster
        // edge rule wraps across range lines, which a tiled tensor map cannot express -> manual loads
        if (h->encode_tiled && !args.raster && ((uintptr_t)rd_map & 15) == 0 && maps < (1ll << 31)) {
            typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            const cuuint64_t gdim[3] = {(cuuint64_t)c.n_chirps, (cuuint64_t)c.n_samples, (cuuint64_t)maps};
            const cuuint64_t gstr[2] = {(cuuint64_t)c.n_chirps * 4, (cuuint64_t)c.n_chirps * c.n_samples * 4};
            const cuuint32_t box[3] = {(cuuint32_t)h->k3.scols, (cuuint32_t)h->k3.srows, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = reinterpret_cast<encode_fn>(h->encode_tiled)(
                &tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(rd_map), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            use_tma = (r == CUDA_SUCCESS);
        }
#endif
        h->cfar_tma = use_tma;
        long long grid = n_tiles < h->k3_grid_cap ? n_tiles : h->k3_grid_cap;
        // RDP_K3_TILES_PER_CTA > 0: more, shorter-lived CTAs than fit at once, so that SM resources turn over and the
        // FFT kernels of the next chunk (higher-priority stream) get placed between them
        if (h->knob_k3_tiles_per_cta > 0) {
            const long long want = (n_tiles + h->knob_k3_tiles_per_cta - 1) / h->knob_k3_tiles_per_cta;
            if (want > grid) grid = want;
        }
        // tiles differ several-fold in cost: the persistent CTAs draw them from a queue (RDP_K3_QUEUE=0: fixed stride)
        int* queue = nullptr;
        if (h->knob_k3_queue && h->k3_queue && n_tiles > grid) {
            RDP_CUDA(h, cudaMemsetAsync(h->k3_queue, 0, sizeof(int), h->cur));
            queue = h->k3_queue;
        }
        h->k3.launch(grid, h->k3.smem, h->cur, rd_map, cfar_out, h->det_bits, args, tiles_r, tiles_c, n_tiles, use_tma,
                     &tmap, int_mode ? cpi_flags : nullptr, queue);
    }
    stage_end(h);
    int rc = launch_check(h, "k3_cfar_kernel");
    if (rc) return rc;
    const uint32_t* bits = h->det_bits;
    stage_begin(h, 3);
    if (h->cfg.peak_group) {
        const long long n_words = maps * c.n_samples * args.words_per_row;
        if (!h->det_bits2 || (size_t)n_words > h->det_bits2_cap) {
            if (h->det_bits2) { cudaFree(h->det_bits2); h->scratch_bytes -= (int64_t)(h->det_bits2_cap * sizeof(uint32_t)); h->det_bits2 = nullptr; }
            const size_t cap = (size_t)n_words > h->cfar_rows_cap * (size_t)args.words_per_row ? (size_t)n_words : h->cfar_rows_cap * (size_t)args.words_per_row;
            int rc2 = dev_alloc(h, &h->det_bits2, cap);
            if (rc2) { stage_end(h); return rc2; }
            h->det_bits2_cap = cap;
        }
        h->launches += 1;
        RDP_LAUNCH(k3b_peak_group_kernel, (unsigned)((n_words + 255) / 256), 256, 0, h->cur, h->det_bits, h->det_bits2, rd_map, args, n_words);
        bits = h->det_bits2;
    }
    RDP_LAUNCH(k4_row_offsets_kernel, n_cpi, K4_THREADS, 0, h->cur, bits, h->row_off, counts, status, args, cpi_flags);
    const long long total_rows = maps * c.n_samples;
    int lanes_per_row = 1;
    while (lanes_per_row < args.words_per_row && lanes_per_row < 32) lanes_per_row <<= 1;
    const long long rows_per_cta = 8ll * (32 / lanes_per_row);
    RDP_LAUNCH(k4_emit_kernel, (unsigned)((total_rows + rows_per_cta - 1) / rows_per_cta), 256, 0, h->cur, bits,
               h->row_off, rd_map, dets, args, total_rows, cpi_label, lanes_per_row);
    h->launches += 1;
    stage_end(h);
    return launch_check(h, "k4 compaction kernels");
}

__global__ void __launch_bounds__(256)
k5_pack_dets_kernel(const rdp_det* __restrict__ dets, const int* __restrict__ counts, rdp_det* __restrict__ packed,
                    int max_dets, int cpi_offset, int cpi_first) {
    // one CTA per CPI (cpi_first + blockIdx.x): offset = sum of the clipped counts of the CPIs before it
    __shared__ long long offset_sh;
    const int cpi = cpi_first + blockIdx.x;
    if (threadIdx.x == 0) {
        long long off = 0;
        for (int i = 0; i < cpi; ++i) off += counts[i] < max_dets ? counts[i] : max_dets;
        offset_sh = off;
    }
    __syncthreads();
    const int n = counts[cpi] < max_dets ? counts[cpi] : max_dets;
    const rdp_det* row = dets + (long long)cpi * max_dets;
    for (int i = threadIdx.x; i < n; i += 256) {
        rdp_det d = row[i];
        d.cpi += cpi_offset;
        packed[offset_sh + i] = d;
    }
}

// ---- arith = FIXED16: tables and the three-kernel chain in front of the integer CFAR -------------------------------
// HAMMING_ROM (window_multiplier.vhd:34-49) read through the mirrored address (:97-102): one coefficient per sample
std::vector<int> fixed_window_rom(int n) {
    std::vector<int> coef((size_t)n);
    const double pi = 3.14159265358979323846;
    const int rom_size = n / 2;
    for (int i = 0; i < n; ++i) {
        int a = i < rom_size ? i : n - 1 - i;
        if (a > rom_size - 1) a = rom_size - 1;
        long v = (long)std::floor((0.54 - 0.46 * std::cos(2.0 * pi * (double)a / (double)(n - 1))) * 32767.0 + 0.5);
        coef[(size_t)i] = (int)(v > 32767 ? 32767 : (v < 0 ? 0 : v));
    }
    return coef;
}
// Q15 twiddles W_n^k, k < n / 2, as stream words {im[31:16], re[15:0]} (oracle/rtl_fixed.py:bfp_twiddles)
std::vector<uint32_t> bfp_twiddle_words(int n) {
    std::vector<uint32_t> w((size_t)(n / 2 > 0 ? n / 2 : 1));
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < n / 2; ++k) {
        const double ang = 2.0 * pi * (double)k / (double)n;
        const int re = (int)std::floor(std::cos(ang) * 32767.0 + 0.5), im = (int)std::floor(-std::sin(ang) * 32767.0 + 0.5);
        w[(size_t)k] = ((uint32_t)(uint16_t)(int16_t)im << 16) | (uint32_t)(uint16_t)(int16_t)re;
    }
    return w;
}
template <class T>
int upload_vec(rdp_handle* h, const std::vector<T>& v, T** dev) {
    int rc = dev_alloc(h, dev, v.size());
    if (rc) return rc;
    RDP_CUDA(h, cudaMemcpy(*dev, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return RDP_OK;
}

int launch_bfp(rdp_handle* h, const uint32_t* in, uint32_t* out, int* exps, const uint32_t* tw, int n, long long n_frames,
               const BfpFuse& fu) {
    if (n_frames == 0) return RDP_OK;
    int log2n = 0;
    while ((1 << log2n) < n) ++log2n;
    const int wpc = 8;                                            // one warp per frame, 8 frames in flight per CTA
    const size_t smem = (size_t)wpc * n * sizeof(uint32_t);
    long long blocks = (n_frames + wpc - 1) / wpc;
    const long long cap = (long long)h->sm_count * (smem > 64 * 1024 ? 1 : (smem > 32 * 1024 ? 3 : 6));
    if (blocks > cap) blocks = cap;
    h->launches += 1;
#ifndef RDP_EMU
#define RDP_BFP_ATTR(NB) if (smem > 48 * 1024) RDP_CUDA(h, cudaFuncSetAttribute(kf_bfp_fft_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
#else
#define RDP_BFP_ATTR(NB)
#endif
#define RDP_BFP_CASE(NB) \
    { RDP_BFP_ATTR(NB); RDP_LAUNCH(kf_bfp_fft_kernel<NB>, (unsigned)blocks, 256, smem, h->cur, in, out, exps, tw, n, log2n, n_frames, wpc, fu); }
    switch (n) {                                                  // 64 ... 1024 points: the stage's butterflies stay in registers
        case 64: RDP_BFP_CASE(1) break;
        case 128: RDP_BFP_CASE(2) break;
        case 256: RDP_BFP_CASE(4) break;
        case 512: RDP_BFP_CASE(8) break;
        case 1024: RDP_BFP_CASE(16) break;
        default: RDP_BFP_CASE(0) break;
    }
#undef RDP_BFP_CASE
#undef RDP_BFP_ATTR
    return launch_check(h, "kf_bfp_fft_kernel");
}

int ensure_fixed_tables(rdp_handle* h) {
    if (h->btw_r) return RDP_OK;
    const rdp_config& c = h->cfg;
    int rc = upload_vec(h, fixed_window_rom(c.n_samples), &h->rom_r);
    if (rc) return rc;
    rc = upload_vec(h, fixed_window_rom(c.n_chirps), &h->rom_d);
    if (rc) return rc;
    rc = upload_vec(h, bfp_twiddle_words(c.n_samples), &h->btw_r);
    if (rc) return rc;
    rc = upload_vec(h, bfp_twiddle_words(c.n_chirps), &h->btw_d);
    if (rc) return rc;
    rc = dev_alloc(h, &h->exp_r, (size_t)h->chunk * c.n_channels * c.n_chirps);
    if (rc) return rc;
    return dev_alloc(h, &h->exp_d, (size_t)h->chunk * c.n_channels * c.n_samples);
}

// window_multiplier -> range transform -> corner_turner -> doppler_notch -> window_multiplier -> Doppler transform ->
// magnitude_calc on n_cpi <= chunk CPIs (radar_core.vhd:267-374): three launches, the windows / MTI / magnitude fused
// into the load and store of the two block-floating-point transforms.  The range-major intermediate (sized for
// complex64) holds the two 32-bit-word buffers.  mag_map receives the 17-bit magnitude words.
int run_fixed_front(rdp_handle* h, const void* cube, int n_cpi, uint32_t* mag_map, uint32_t* cpi_flags, int* exp_r_out,
                    int* exp_d_out) {
    const rdp_config& c = h->cfg;
    if (n_cpi == 0) return RDP_OK;
    int rc = ensure_fixed_tables(h);
    if (rc) return rc;
    const long long cells = (long long)cells_per_cpi(c), maps = (long long)n_cpi * c.n_channels;
    uint32_t* a = reinterpret_cast<uint32_t*>(h->x_rc);
    uint32_t* b = a + (size_t)h->chunk * cells;
    BfpFuse fr;
    fr.adc_check = 1; fr.coef = h->rom_r; fr.sat_flags = cpi_flags; fr.frames_per_flag = (long long)c.n_channels * c.n_chirps;
    fr.flag_bits = RDP_STATUS_WINDOW_SAT; fr.mti_mode = 0; fr.mag_out = 0;
    stage_begin(h, 0);
    rc = launch_bfp(h, reinterpret_cast<const uint32_t*>(cube), a, exp_r_out ? exp_r_out : h->exp_r, h->btw_r, c.n_samples,
                    maps * c.n_chirps, fr);
    if (rc) { stage_end(h); return rc; }
    if (maps > 65535) { stage_end(h); return fail(h, RDP_E_CAPACITY, "fixed-point chain: more than 65535 maps per chunk; lower chunk_cpis"); }
    h->launches += 1;
    RDP_LAUNCH(kf_corner_turn_kernel, dim3((c.n_samples + 31) / 32, (c.n_chirps + 31) / 32, (unsigned)maps), 256, 0, h->cur, a, b,
               c.n_chirps, c.n_samples);
    stage_end(h);
    rc = launch_check(h, "kf_corner_turn_kernel");
    if (rc) return rc;
    BfpFuse fd;
    fd.adc_check = 0; fd.coef = h->rom_d; fd.sat_flags = cpi_flags; fd.frames_per_flag = (long long)c.n_channels * c.n_samples;
    fd.flag_bits = RDP_STATUS_WINDOW_SAT; fd.mti_mode = c.mti_mode; fd.mag_out = 1;
    stage_begin(h, 1);
    rc = launch_bfp(h, b, mag_map, exp_d_out ? exp_d_out : h->exp_d, h->btw_d, c.n_chirps, maps * c.n_samples, fd);
    stage_end(h);
    return rc;
}

// Everything in front of the CFAR stage for n_cpi <= chunk CPIs in the handle's arithmetic; `map` receives float32
// magnitudes (FLOAT32) or 17-bit magnitude words (FIXED16).  flags: this chunk's pre-CFAR status bits (zeroed here).
int run_front(rdp_handle* h, const void* cube, int n_cpi, float* map, uint32_t* flags, int* exp_r_out = nullptr,
              int* exp_d_out = nullptr) {
    if (flags) RDP_CUDA(h, cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)n_cpi, h->cur));
    if (h->cfg.arith == RDP_ARITH_FIXED16)
        return run_fixed_front(h, cube, n_cpi, reinterpret_cast<uint32_t*>(map), flags, exp_r_out, exp_d_out);
    int rc = run_range(h, cube, n_cpi, h->x_rc, nullptr, flags);
    if (rc) return rc;
    return run_doppler(h, h->x_rc, n_cpi, map);
}

int check_cfg(const rdp_config* c) {
    if (!c) return fail(nullptr, RDP_E_INVALID, "config is NULL");
    if (c->struct_size != (int32_t)sizeof(rdp_config))
        return fail(nullptr, RDP_E_INVALID, "rdp_config.struct_size %d != %d (ABI mismatch)", c->struct_size,
                    (int)sizeof(rdp_config));
    if (!is_pow2(c->n_chirps) || c->n_chirps < 8 || c->n_chirps > 4096)
        return fail(nullptr, RDP_E_INVALID, "n_chirps must be a power of two in [8, 4096], got %d", c->n_chirps);
    if (!is_pow2(c->n_samples) || c->n_samples < 8 || c->n_samples > 4096)
        return fail(nullptr, RDP_E_INVALID, "n_samples must be a power of two in [8, 4096], got %d", c->n_samples);
    if (c->n_channels < 1 || c->n_channels > 65535) return fail(nullptr, RDP_E_INVALID, "n_channels out of range");
    if (c->cfar_ref_r < 0 || c->cfar_ref_d < 0 || c->cfar_guard_r < 0 || c->cfar_guard_d < 0 ||
        c->cfar_ref_r + c->cfar_ref_d == 0)
        return fail(nullptr, RDP_E_INVALID, "bad CFAR extents");
    if (c->rank_pct < 0 || c->rank_pct > 100) return fail(nullptr, RDP_E_INVALID, "rank_pct must be 0..100");
    if (c->scale_min < 1 || c->scale_nom < 1 || c->scale_max < 1) return fail(nullptr, RDP_E_INVALID, "scales must be >= 1");
    if (c->scale_override < 0 || c->scale_override > 7)
        return fail(nullptr, RDP_E_INVALID, "scale_override is a 3-bit field (0..7), got %d", c->scale_override);
    if (c->mti_mode != 0 && c->mti_mode != 2 && c->mti_mode != 3)
        return fail(nullptr, RDP_E_INVALID, "mti_mode must be 0 (bypass), 2 or 3");
    if (c->input_format != RDP_IN_INT16_IQ && c->input_format != RDP_IN_COMPLEX64)
        return fail(nullptr, RDP_E_INVALID, "unknown input_format %d", c->input_format);
    if (c->mag_mode < 0 || c->mag_mode > 3) return fail(nullptr, RDP_E_INVALID, "unknown mag_mode %d", c->mag_mode);
    if (c->cfar_axes < 0 || c->cfar_axes > 1 || c->cfar_edges < 0 || c->cfar_edges > 1)
        return fail(nullptr, RDP_E_INVALID, "unknown cfar_axes / cfar_edges");
    if (c->max_dets_per_cpi < 1 || c->max_dets_per_cpi > (1 << 24))
        return fail(nullptr, RDP_E_INVALID, "max_dets_per_cpi must be 1..16777216");
    if (c->max_cpis < 1) return fail(nullptr, RDP_E_INVALID, "max_cpis must be >= 1");
    if (c->chunk_cpis < 0) return fail(nullptr, RDP_E_INVALID, "chunk_cpis must be >= 0");
    if (c->arith != RDP_ARITH_FLOAT32 && c->arith != RDP_ARITH_FIXED16) return fail(nullptr, RDP_E_INVALID, "unknown arith %d", c->arith);
    if (c->combine < RDP_COMBINE_NONE || c->combine > RDP_COMBINE_BEAMS) return fail(nullptr, RDP_E_INVALID, "unknown combine %d", c->combine);
    if (c->peak_group != 0 && c->peak_group != 1) return fail(nullptr, RDP_E_INVALID, "peak_group must be 0 or 1");
    if (c->combine != RDP_COMBINE_NONE && c->arith == RDP_ARITH_FIXED16)
        return fail(nullptr, RDP_E_INVALID, "channel combining is not part of the RTL chain: use arith FLOAT32");
    if (c->arith == RDP_ARITH_FIXED16) {
        if (c->input_format != RDP_IN_INT16_IQ || c->mag_mode != RDP_MAG_AMB || c->window_gain != 1.0f)
            return fail(nullptr, RDP_E_INVALID, "arith FIXED16 needs int16 IQ input, mag_mode AMB and window_gain 1.0");
        if (c->scale_min > 15 || c->scale_nom > 15 || c->scale_max > 15)
            return fail(nullptr, RDP_E_INVALID, "arith FIXED16: the RTL's threshold multiplier is 4 bits wide, scales must be 1..15");
    }
    return RDP_OK;
}

void release(rdp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto& se : h->stage_events) { cudaEventDestroy(se.a); cudaEventDestroy(se.b); }
    for (auto& sl : h->slots) {
        void* dev[] = {sl.d_in, sl.d_dets, sl.d_packed, sl.d_counts, sl.d_status};
        for (void* p : dev) if (p) cudaFree(p);
        void* hst[] = {sl.h_packed, sl.h_counts, sl.h_status};
        for (void* p : hst) if (p) cudaFreeHost(p);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.ev_out) cudaEventDestroy(sl.ev_out);
        if (sl.copy) cudaStreamDestroy(sl.copy);
#ifndef RDP_EMU
        if (sl.graph) cudaGraphExecDestroy(sl.graph);
#endif
    }
    h->slots.clear();
    for (auto e : h->ev_copied) cudaEventDestroy(e);
    for (auto e : h->ev_done) cudaEventDestroy(e);
    for (auto e : h->ev_mapout) cudaEventDestroy(e);
    for (auto e : h->ev_counts) cudaEventDestroy(e);
    if (h->t0) cudaEventDestroy(h->t0);
    if (h->t1) cudaEventDestroy(h->t1);
    void* ptrs[] = {h->win_r, h->win_d, h->tw_r, h->tw_d, h->x_rc, h->map_scratch[0], h->map_scratch[1],
                    h->in_stage[0], h->in_stage[1], h->det_bits, h->row_off, h->rom_dev, h->sat_dev, h->trk_state, h->counts, h->status, h->dets, h->packed, h->l2_flush,
                    h->cpi_flags[0], h->cpi_flags[1], h->det_bits2, h->steer, h->k3_queue, h->rom_r, h->rom_d, h->btw_r, h->btw_d, h->exp_r, h->exp_d, h->btw_tap};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->cfar_stream) cudaStreamDestroy(h->cfar_stream);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_map_ready[i]) cudaEventDestroy(h->ev_map_ready[i]);
        if (h->ev_cfar_done[i]) cudaEventDestroy(h->ev_cfar_done[i]);
    }
    delete h;
}

int upload_tables(rdp_handle* h, int n, float gain, float** win, float2** tw) {
    std::vector<float> w(n);
    std::vector<float2> t(n);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < n; ++i) {
        // Hamming, same formula as the ROM generator (window_multiplier.vhd:41-42); the float32
        // coefficient is formed first, then the gain, exactly as the oracle does
        float coef = (float)(0.54 - 0.46 * std::cos(2.0 * pi * (double)i / (double)(n - 1)));
        w[i] = (float)((double)coef * (double)gain);
        double ang = -2.0 * pi * (double)i / (double)n;
        t[i] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    int rc = dev_alloc(h, win, (size_t)n);
    if (rc) return rc;
    rc = dev_alloc(h, tw, (size_t)n);
    if (rc) return rc;
    RDP_CUDA(h, cudaMemcpy(*win, w.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    RDP_CUDA(h, cudaMemcpy(*tw, t.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    return RDP_OK;
}

int init_handle(rdp_handle* h) {
    const rdp_config& c = h->cfg;
    cudaDeviceProp prop;
    RDP_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    {   // every RDP_* tuning knob is read ONCE, here (rdp.h lists them); nothing on a launch path calls getenv
        const char* e;
        if ((e = std::getenv("RDP_OVERLAP"))) h->overlap = std::atoi(e);
        if ((e = std::getenv("RDP_K3_VARIANT"))) h->k3_variant = std::atoi(e);
        if ((e = std::getenv("RDP_K3_CTAS"))) h->knob_k3_ctas = std::atoi(e);
        if ((e = std::getenv("RDP_K3_TILES_PER_CTA"))) h->knob_k3_tiles_per_cta = std::atoi(e);
        if ((e = std::getenv("RDP_K12_CTAS"))) h->knob_k12_ctas = std::atoi(e);
        if ((e = std::getenv("RDP_K1_CLUSTER"))) h->knob_k1_cluster = std::atoi(e) != 0;
        if ((e = std::getenv("RDP_K1_TMA"))) h->knob_k1_tma = std::atoi(e);
        if ((e = std::getenv("RDP_STREAM_PRIORITY"))) h->knob_priority = std::atoi(e);
        if ((e = std::getenv("RDP_K3_DEBUG_SKIP"))) h->knob_k3_debug_skip = std::atoi(e);
        if ((e = std::getenv("RDP_K3_SKEW_NS"))) h->knob_k3_skew_ns = std::atoi(e);
        if ((e = std::getenv("RDP_STREAM_SPEC")) && std::atoi(e) > 0) h->stream_spec_dets = std::atoi(e);
        if ((e = std::getenv("RDP_STREAM_GRAPH"))) h->knob_stream_graph = std::atoi(e);
        if ((e = std::getenv("RDP_K3_QUEUE"))) h->knob_k3_queue = std::atoi(e);
    }
    // the FFT kernels (HBM-bound) run on `stream`, the CFAR kernels (ALU-bound) of the previous chunk on `cfar_stream`;
    // the FFT stream has the higher priority so that its CTAs are placed first whenever CFAR CTAs retire
    int prio_least = 0, prio_greatest = 0;
    RDP_CUDA(h, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    const int prio_fft = h->knob_priority ? prio_greatest : prio_least;
    RDP_CUDA(h, cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_fft));
    RDP_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    RDP_CUDA(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    RDP_CUDA(h, cudaStreamCreateWithPriority(&h->cfar_stream, cudaStreamNonBlocking, prio_least));
    h->cur = h->stream;
    for (int i = 0; i < 2; ++i) {
        RDP_CUDA(h, cudaEventCreateWithFlags(&h->ev_map_ready[i], cudaEventDefault));
        RDP_CUDA(h, cudaEventCreateWithFlags(&h->ev_cfar_done[i], cudaEventDefault));
    }
    RDP_CUDA(h, cudaEventCreate(&h->t0));
    RDP_CUDA(h, cudaEventCreate(&h->t1));

    int rc = upload_tables(h, c.n_samples, c.window_gain, &h->win_r, &h->tw_r);
    if (rc) return rc;
    rc = upload_tables(h, c.n_chirps, c.window_gain, &h->win_d, &h->tw_d);
    if (rc) return rc;

    // kernels
    bool ok = (c.input_format == RDP_IN_INT16_IQ) ? k1_for<0>(c.n_samples, &h->k1) : k1_for<1>(c.n_samples, &h->k1);
    if (!ok) return fail(h, RDP_E_INVALID, "no range-FFT plan for n_samples=%d", c.n_samples);
    switch (c.mag_mode) {
        case 0: ok = k2_for<0>(c.n_chirps, &h->k2); break;
        case 1: ok = k2_for<1>(c.n_chirps, &h->k2); break;
        case 2: ok = k2_for<2>(c.n_chirps, &h->k2); break;
        default: ok = k2_for<3>(c.n_chirps, &h->k2); break;
    }
    if (!ok) return fail(h, RDP_E_INVALID, "no Doppler-FFT plan for n_chirps=%d", c.n_chirps);
    h->k1_tile_chirps = h->k1.tc_max < c.n_chirps ? h->k1.tc_max : c.n_chirps;
    h->k1_smem = h->k1.smem(h->k1_tile_chirps);
    int occ = 1;
    RDP_CUDA(h, h->k1.prepare(h->k1_smem, &occ));
    h->k1_grid_cap = h->sm_count * occ;
    h->k1_max_clusters = 0;
    if (h->k1.cs > 1 && h->k1.prepare_cluster && h->k1_tile_chirps == h->k1.tc_max &&
        (c.n_chirps / h->k1.tc_max) % h->k1.cs == 0) {
        int n = 0;
        if (h->k1.prepare_cluster(h->k1.smem_cluster, &n) == cudaSuccess && n > 0) h->k1_max_clusters = n;
        else cudaGetLastError();                                  // fall back to the plain kernel
    }
    RDP_CUDA(h, h->k2.prepare(h->k2.smem, &occ));
    h->k2_grid_cap = h->sm_count * occ;

    // CFAR geometry in physical axes (os_cfar_2d.vhd:41-47)
    CfarArgs& a = h->cfar;
    h->map_channels = map_channels_of(c);
    h->n_beams = beams_for(c.n_channels);
    if (c.combine == RDP_COMBINE_BEAMS) {
        std::vector<float2> st((size_t)h->n_beams);
        const double pi = 3.14159265358979323846;
        for (int m = 0; m < h->n_beams; ++m) {
            const double ang = -2.0 * pi * (double)m / (double)h->n_beams;
            st[(size_t)m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        int rc0 = dev_alloc(h, &h->steer, st.size());
        if (rc0) return rc0;
        RDP_CUDA(h, cudaMemcpy(h->steer, st.data(), sizeof(float2) * st.size(), cudaMemcpyHostToDevice));
    }
    a.n_range = c.n_samples; a.n_doppler = c.n_chirps; a.n_channels = h->map_channels;
    if (c.cfar_axes == RDP_AXES_AS_BUILT) {
        a.half_l = c.cfar_ref_d + c.cfar_guard_d; a.guard_l = c.cfar_guard_d;
        a.half_c = c.cfar_ref_r + c.cfar_guard_r; a.guard_c = c.cfar_guard_r;
    } else {
        a.half_l = c.cfar_ref_r + c.cfar_guard_r; a.guard_l = c.cfar_guard_r;
        a.half_c = c.cfar_ref_d + c.cfar_guard_d; a.guard_c = c.cfar_guard_d;
    }
    const int win = (2 * a.half_l + 1) * (2 * a.half_c + 1);
    a.n_ref = win - (2 * a.guard_l + 1) * (2 * a.guard_c + 1);
    if (a.n_ref < 1) return fail(h, RDP_E_INVALID, "CFAR window has no reference cells");
    if (win > 32 * MAXQ_GENERIC) return fail(h, RDP_E_INVALID, "CFAR window of %d cells exceeds %d", win, 32 * MAXQ_GENERIC);
    a.rank_idx = (a.n_ref * c.rank_pct) / 100;                    // os_cfar_2d.vhd:181-182
    if (a.rank_idx >= a.n_ref) a.rank_idx = a.n_ref - 1;
    a.smin = (float)c.scale_min; a.snom = (float)c.scale_nom; a.smax = (float)c.scale_max;
    a.adaptive = c.scale_override == 0;
    if (a.adaptive) {
        int s1 = c.scale_min < c.scale_nom ? c.scale_min : c.scale_nom;
        if (c.scale_max < s1) s1 = c.scale_max;
        a.s1 = (float)s1;
    } else {
        a.s1 = (float)c.scale_override;
    }
    if (a.adaptive) {
        // detection => x_k < max(cut / smax, cut / snom, min(cut / smin, mean / 2)), each term with a 1e-3 margin
        const float big = a.snom < a.smax ? a.snom : a.smax;
        a.pf_a = 1.001f / big;
        a.pf_b = 1.001f / a.smin;
        a.pf_c = 0.5f * 1.002f / (float)a.n_ref;
    } else {
        // fixed scale: fl(ref * s1) < cut  =>  ref < cut / s1 * (1 + 2^-23)
        a.pf_a = 1.001f / a.s1; a.pf_b = 0.0f; a.pf_c = 0.0f;
    }
    if (a.pf_a < 1e-6f) a.pf_a = 1e-6f;                            // keeps thresholds of cuts >= 1e-30 out of the bf16 subnormals
    a.raster = c.cfar_edges == RDP_EDGES_RASTER;
    a.int_mode = 0;
    a.debug_skip = h->knob_k3_debug_skip;
    a.neg1 = -1;
    a.skew_ns = h->knob_k3_skew_ns;
    a.skew_div = h->sm_count > 0 ? h->sm_count : 1;
    a.is1 = (int)a.s1; a.ismin = c.scale_min; a.isnom = c.scale_nom; a.ismax = c.scale_max;
    a.max_dets = c.max_dets_per_cpi;
    a.words_per_row = (c.n_chirps + 31) / 32;
    h->k3 = k3_for(a.half_l, a.half_c, a.guard_l, a.guard_c, h->k3_variant);
    RDP_CUDA(h, h->k3.prepare(h->k3.smem, &occ));
    h->k3_grid_cap = h->sm_count * occ;
    // RDP_K3_CTAS / RDP_K12_CTAS cap the persistent CTAs per SM (co-residency experiments)
    if (h->knob_k3_ctas > 0) h->k3_grid_cap = h->sm_count * h->knob_k3_ctas;
    if (h->knob_k12_ctas > 0) {
        h->k1_grid_cap = h->sm_count * h->knob_k12_ctas;
        h->k2_grid_cap = h->sm_count * h->knob_k12_ctas;
    }
#ifndef RDP_EMU
    {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            h->encode_tiled = fn;
        cudaGetLastError();
    }
#endif

    // chunking.  Measured on B200 (profiles/r2_summary.md): every launch costs ~12-15 us of ramp-up and tail on top of
    // its work, so a pass over few CPIs loses more than an L2-resident intermediate would save; with the automotive
    // shape the device-resident path is flat from 128 CPIs per chunk (1.5 GiB of range-major intermediate) up and 6 %
    // slower at 32.  The HOST path moves its chunks through PCIe while the previous one is computed, so it wants
    // SHORT chunks (the last one's kernels are exposed): it walks the same scratch in sub-chunks of <= 384 MiB.
    const size_t xrc_per_cpi = cells_per_cpi(c) * sizeof(float2);
    int chunk = c.chunk_cpis;
    if (chunk == 0) {
        size_t budget = (size_t)1536 << 20;
        chunk = (int)(budget / xrc_per_cpi);
        if (chunk < 1) chunk = 1;
        if (chunk > 256) chunk = 256;
    }
    if (chunk > c.max_cpis) chunk = c.max_cpis;
    h->chunk = chunk;
    h->host_chunk = (int)(((size_t)384 << 20) / xrc_per_cpi);
    if (h->host_chunk < 1) h->host_chunk = 1;
    if (h->host_chunk > chunk || c.chunk_cpis != 0) h->host_chunk = chunk;   // an explicit chunk_cpis is honoured as given
    rc = dev_alloc(h, &h->x_rc, (size_t)chunk * cells_per_cpi(c));
    if (rc) return rc;
    for (int i = 0; i < 2; ++i) {
        rc = dev_alloc(h, &h->map_scratch[i], (size_t)chunk * map_cells_per_cpi(c));
        if (rc) return rc;
    }
    rc = dev_alloc(h, &h->det_bits, (size_t)chunk * h->map_channels * c.n_samples * h->cfar.words_per_row);
    if (rc) return rc;
    rc = dev_alloc(h, &h->row_off, (size_t)chunk * h->map_channels * c.n_samples);
    if (rc) return rc;
    h->cfar_rows_cap = (size_t)chunk * h->map_channels * c.n_samples;
    rc = dev_alloc(h, &h->counts, (size_t)c.max_cpis);
    if (rc) return rc;
    rc = dev_alloc(h, &h->status, (size_t)c.max_cpis);
    if (rc) return rc;
    for (int i = 0; i < 2; ++i) {
        rc = dev_alloc(h, &h->cpi_flags[i], (size_t)chunk);
        if (rc) return rc;
    }
    rc = dev_alloc(h, &h->k3_queue, 4);
    if (rc) return rc;
    if (c.arith == RDP_ARITH_FIXED16) {
        if (h->k3.trows == 0) return fail(h, RDP_E_INVALID, "arith FIXED16: no tiled integer CFAR instance for this window shape");
        rc = ensure_fixed_tables(h);
        if (rc) return rc;
    }
    return RDP_OK;
}

int ensure_host_path(rdp_handle* h) {
    const rdp_config& c = h->cfg;
    if (h->in_stage[0]) return RDP_OK;
    for (int i = 0; i < 2; ++i) {
        unsigned char* p = nullptr;
        int rc = dev_alloc(h, &p, (size_t)h->chunk * cells_per_cpi(c) * in_bytes_per_sample(c));
        if (rc) return rc;
        h->in_stage[i] = p;
    }
    int rc = dev_alloc(h, &h->dets, (size_t)c.max_cpis * c.max_dets_per_cpi);
    if (rc) return rc;
    if (!h->packed) rc = dev_alloc(h, &h->packed, (size_t)c.max_cpis * c.max_dets_per_cpi);
    if (rc) return rc;
    for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        RDP_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDefault)); h->ev_copied.push_back(e);
        RDP_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDefault)); h->ev_done.push_back(e);
        RDP_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDefault)); h->ev_mapout.push_back(e);
        RDP_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDefault)); h->ev_counts.push_back(e);
    }
    return RDP_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

#ifdef RDP_EMU
int rdp_emu_marker(void) { return 1; }   // the product loader refuses libraries exporting this
#endif

void rdp_default_config(rdp_config* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (int32_t)sizeof(rdp_config);
    cfg->n_chirps = 128;          // N_DOPPLER  radar_core.vhd:14
    cfg->n_samples = 1024;        // N_RANGE    radar_core.vhd:13
    cfg->n_channels = 1;
    cfg->cfar_ref_r = 4; cfg->cfar_ref_d = 4; cfg->cfar_guard_r = 2; cfg->cfar_guard_d = 1;   // :16-19
    cfg->rank_pct = 75; cfg->scale_min = 2; cfg->scale_nom = 4; cfg->scale_max = 6;           // :380
    cfg->scale_override = 0;
    cfg->mti_mode = 0;
    cfg->input_format = RDP_IN_INT16_IQ;
    cfg->mag_mode = RDP_MAG_AMB;
    cfg->cfar_axes = RDP_AXES_AS_BUILT;
    cfg->cfar_edges = RDP_EDGES_ZERO;
    cfg->window_gain = 1.0f;
    cfg->max_dets_per_cpi = 4096;
    cfg->max_cpis = 1;
    cfg->chunk_cpis = 0;
    cfg->arith = RDP_ARITH_FLOAT32;
    cfg->combine = RDP_COMBINE_NONE;
    cfg->peak_group = 0;
}

int rdp_create(const rdp_config* cfg, int32_t device, rdp_handle** out) {
    if (!out) return fail(nullptr, RDP_E_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_cfg(cfg);
    if (rc) return rc;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev < 1)
        return fail(nullptr, RDP_E_NO_DEVICE, "no CUDA device available (%s); librdp has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n_dev) return fail(nullptr, RDP_E_NO_DEVICE, "device %d out of range [0, %d)", device, n_dev);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, RDP_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    rdp_handle* h = new (std::nothrow) rdp_handle();
    if (!h) return fail(nullptr, RDP_E_NOMEM, "out of host memory");
    h->cfg = *cfg;
    h->device = device;
    rc = init_handle(h);
    if (rc) {
        g_create_error = h->err;
        release(h);
        return rc;
    }
    *out = h;
    return RDP_OK;
}

int rdp_destroy(rdp_handle* h) {
    if (!h) return RDP_E_INVALID;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    release(h);
    return RDP_OK;
}

const char* rdp_last_error(const rdp_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rdp_get_info(const rdp_handle* h, rdp_info* info) {
    if (!h || !info) return RDP_E_INVALID;
    std::memset(info, 0, sizeof(*info));
    info->abi_version = RDP_ABI_VERSION;
    info->device = h->device;
    info->sm_count = h->sm_count;
    info->chunk_cpis = h->chunk;
    info->scratch_bytes = h->scratch_bytes;
    info->kernel_launches = h->launches;
    info->cfar_tma = h->cfar_tma;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess)
        std::snprintf(info->device_name, sizeof(info->device_name), "%s", prop.name);
    return RDP_OK;
}

int rdp_sync(rdp_handle* h) {
    if (!h) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return RDP_OK;
}

int rdp_range_stage(rdp_handle* h, const void* cube, int32_t n_cpi, void* x_rc) {
    if (!h || !cube || !x_rc || n_cpi < 0) return fail(h, RDP_E_INVALID, "rdp_range_stage: bad argument");
    RDP_CUDA(h, cudaSetDevice(h->device));
    return run_range(h, cube, n_cpi, reinterpret_cast<float2*>(x_rc));
}

int rdp_doppler_stage(rdp_handle* h, const void* x_rc, int32_t n_cpi, float* rd_map) {
    if (!h || !x_rc || !rd_map || n_cpi < 0) return fail(h, RDP_E_INVALID, "rdp_doppler_stage: bad argument");
    RDP_CUDA(h, cudaSetDevice(h->device));
    return run_doppler(h, reinterpret_cast<const float2*>(x_rc), n_cpi, rd_map);
}

int rdp_cfar_stage(rdp_handle* h, const float* rd_map, int32_t n_cpi, float* cfar_out, rdp_det* dets,
                   int32_t* n_dets, uint32_t* status) {
    if (!h || !rd_map || !dets || !n_dets || !status || n_cpi < 0)
        return fail(h, RDP_E_INVALID, "rdp_cfar_stage: bad argument");
    const rdp_config& c = h->cfg;
    RDP_CUDA(h, cudaSetDevice(h->device));
    const size_t cells = map_cells_per_cpi(c);
    for (int base = 0; base < n_cpi; base += h->chunk) {
        const int n = (n_cpi - base) < h->chunk ? (n_cpi - base) : h->chunk;
        int rc = run_cfar(h, rd_map + (size_t)base * cells, n, cfar_out ? cfar_out + (size_t)base * cells : nullptr,
                          dets + (size_t)base * c.max_dets_per_cpi, n_dets + base, status + base, base);
        if (rc) return rc;
    }
    return RDP_OK;
}

int rdp_cfar_stage_u32(rdp_handle* h, const uint32_t* mag_map, int32_t n_cpi, uint32_t* cfar_out, rdp_det* dets,
                       int32_t* n_dets, uint32_t* status) {
    if (!h || !mag_map || !dets || !n_dets || !status || n_cpi < 0)
        return fail(h, RDP_E_INVALID, "rdp_cfar_stage_u32: bad argument");
    if (h->k3.trows == 0)
        return fail(h, RDP_E_INVALID, "rdp_cfar_stage_u32: no tiled kernel instance for this CFAR window shape");
    if (h->cfg.scale_min > 15 || h->cfg.scale_nom > 15 || h->cfg.scale_max > 15)
        return fail(h, RDP_E_INVALID, "rdp_cfar_stage_u32: the RTL's threshold multiplier is 4 bits wide (os_cfar_2d.vhd:204), scales must be 1..15");
    const rdp_config& c = h->cfg;
    RDP_CUDA(h, cudaSetDevice(h->device));
    const size_t cells = map_cells_per_cpi(c);
    for (int base = 0; base < n_cpi; base += h->chunk) {
        const int n = (n_cpi - base) < h->chunk ? (n_cpi - base) : h->chunk;
        // the kernels move 32-bit words; in integer mode they are read as unsigned integers
        RDP_CUDA(h, cudaMemsetAsync(h->cpi_flags[0], 0, sizeof(uint32_t) * (size_t)n, h->stream));   // RDP_STATUS_MAP_RANGE
        int rc = run_cfar(h, reinterpret_cast<const float*>(mag_map) + (size_t)base * cells, n,
                          cfar_out ? reinterpret_cast<float*>(cfar_out) + (size_t)base * cells : nullptr,
                          dets + (size_t)base * c.max_dets_per_cpi, n_dets + base, status + base, base, 1, 0, h->cpi_flags[0]);
        if (rc) return rc;
    }
    return RDP_OK;
}

// scope guard: the run_* helpers launch on h->cur; whatever path leaves a function, it is the main stream again
struct CurStreamGuard {
    rdp_handle* h;
    explicit CurStreamGuard(rdp_handle* hh) : h(hh) {}
    ~CurStreamGuard() { h->cur = h->stream; }
};

static int process_device_impl(rdp_handle* h, const void* cube, int32_t n_cpi, float* rd_map, int32_t* exp_r, int32_t* exp_d,
                               rdp_det* dets, int32_t* n_dets, uint32_t* status) {
    const rdp_config& c = h->cfg;
    RDP_CUDA(h, cudaSetDevice(h->device));
    const size_t cells = map_cells_per_cpi(c);
    const size_t in_stride = cells_per_cpi(c) * in_bytes_per_sample(c);
    const int int_mode = c.arith == RDP_ARITH_FIXED16;
    const size_t er_stride = (size_t)c.n_channels * c.n_chirps, ed_stride = (size_t)c.n_channels * c.n_samples;
    CurStreamGuard guard(h);
    if (h->overlap && !h->stage_timing) {      // per-kernel timing wants the kernels alone on the GPU
        // two-stream chunk pipeline: FFT kernels (HBM-bound) of chunk i+1 on `stream`, CFAR (ALU-bound) of chunk i
        // on `cfar_stream`; the map scratch and the flag words are double-buffered, det_bits / row_off belong to the
        // CFAR stream
        int k = 0;
        for (int base = 0; base < n_cpi; base += h->chunk, ++k) {
            const int n = (n_cpi - base) < h->chunk ? (n_cpi - base) : h->chunk;
            const int b = k & 1;
            const unsigned char* in = reinterpret_cast<const unsigned char*>(cube) + (size_t)base * in_stride;
            float* map = rd_map ? rd_map + (size_t)base * cells : h->map_scratch[b];
            h->cur = h->stream;
            if (k >= 2) RDP_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_cfar_done[b], 0));
            int rc = run_front(h, in, n, map, h->cpi_flags[b], exp_r ? exp_r + (size_t)base * er_stride : nullptr,
                               exp_d ? exp_d + (size_t)base * ed_stride : nullptr);
            if (rc) return rc;
            RDP_CUDA(h, cudaEventRecord(h->ev_map_ready[b], h->stream));
            h->cur = h->cfar_stream;
            RDP_CUDA(h, cudaStreamWaitEvent(h->cfar_stream, h->ev_map_ready[b], 0));
            rc = run_cfar(h, map, n, nullptr, dets + (size_t)base * c.max_dets_per_cpi, n_dets + base, status + base, base,
                          int_mode, 0, h->cpi_flags[b]);
            if (rc) return rc;
            RDP_CUDA(h, cudaEventRecord(h->ev_cfar_done[b], h->cfar_stream));
        }
        if (k > 0) RDP_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_cfar_done[(k - 1) & 1], 0));   // rdp_sync / timers see the end
        return RDP_OK;
    }
    for (int base = 0; base < n_cpi; base += h->chunk) {
        const int n = (n_cpi - base) < h->chunk ? (n_cpi - base) : h->chunk;
        const unsigned char* in = reinterpret_cast<const unsigned char*>(cube) + (size_t)base * in_stride;
        float* map = rd_map ? rd_map + (size_t)base * cells : h->map_scratch[0];
        int rc = run_front(h, in, n, map, h->cpi_flags[0], exp_r ? exp_r + (size_t)base * er_stride : nullptr,
                           exp_d ? exp_d + (size_t)base * ed_stride : nullptr);
        if (rc) return rc;
        rc = run_cfar(h, map, n, nullptr, dets + (size_t)base * c.max_dets_per_cpi, n_dets + base, status + base, base, int_mode,
                      0, h->cpi_flags[0]);
        if (rc) return rc;
    }
    return RDP_OK;
}

int rdp_process_device(rdp_handle* h, const void* cube, int32_t n_cpi, float* rd_map, rdp_det* dets,
                       int32_t* n_dets, uint32_t* status) {
    if (!h || !cube || !dets || !n_dets || !status || n_cpi < 0)
        return fail(h, RDP_E_INVALID, "rdp_process_device: bad argument");
    return process_device_impl(h, cube, n_cpi, rd_map, nullptr, nullptr, dets, n_dets, status);
}

int rdp_process_fixed_device(rdp_handle* h, const void* cube, int32_t n_cpi, uint32_t* mag_map, int32_t* exp_range,
                             int32_t* exp_doppler, rdp_det* dets, int32_t* n_dets, uint32_t* status) {
    if (!h || !cube || !dets || !n_dets || !status || n_cpi < 0)
        return fail(h, RDP_E_INVALID, "rdp_process_fixed_device: bad argument");
    if (h->cfg.arith != RDP_ARITH_FIXED16)
        return fail(h, RDP_E_INVALID, "rdp_process_fixed_device: the handle was created with arith FLOAT32");
    return process_device_impl(h, cube, n_cpi, reinterpret_cast<float*>(mag_map), exp_range, exp_doppler, dets, n_dets, status);
}

int rdp_process_host(rdp_handle* h, const void* cube, int32_t n_cpi, float* rd_map, rdp_det* dets,
                     int32_t* n_dets, uint32_t* status) {
    if (!h || !cube || !dets || !n_dets || !status || n_cpi < 0)
        return fail(h, RDP_E_INVALID, "rdp_process_host: bad argument");
    const rdp_config& c = h->cfg;
    if (n_cpi > c.max_cpis) return fail(h, RDP_E_CAPACITY, "n_cpi %d exceeds max_cpis %d", n_cpi, c.max_cpis);
    if (n_cpi == 0) return RDP_OK;
    RDP_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_host_path(h);
    if (rc) return rc;
    const size_t cells = map_cells_per_cpi(c);
    const size_t in_stride = cells_per_cpi(c) * in_bytes_per_sample(c);
    // The detection list goes back chunk by chunk on the D2H stream while later chunks are still being copied in
    // (PCIe is full duplex): after chunk k is queued, the host waits for the counts of chunk k-1 -- long finished,
    // its successor's input copy takes several times its compute -- and queues the copy of its packed records.
    long long list_off = 0;
    auto drain = [&](int base0, int n0, int b0) -> int {
        RDP_CUDA(h, cudaEventSynchronize(h->ev_counts[b0]));
        long long total = 0;
        for (int i = base0; i < base0 + n0; ++i) total += n_dets[i] < c.max_dets_per_cpi ? n_dets[i] : c.max_dets_per_cpi;
        if (total > 0)
            RDP_CUDA(h, cudaMemcpyAsync(dets + list_off, h->packed + list_off, sizeof(rdp_det) * (size_t)total,
                                        cudaMemcpyDeviceToHost, h->d2h_stream));
        list_off += total;
        return RDP_OK;
    };
    int k = 0, prev_base = 0, prev_n = 0;
    for (int base = 0; base < n_cpi; base += h->host_chunk, ++k) {
        const int n = (n_cpi - base) < h->host_chunk ? (n_cpi - base) : h->host_chunk;
        const int b = k & 1;
        // H2D of chunk k on the copy stream, once the kernels that read this staging buffer (chunk k-2) are done
        if (k >= 2) RDP_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));
        RDP_CUDA(h, cudaMemcpyAsync(h->in_stage[b], reinterpret_cast<const unsigned char*>(cube) + (size_t)base * in_stride,
                                    (size_t)n * in_stride, cudaMemcpyHostToDevice, h->copy_stream));
        RDP_CUDA(h, cudaEventRecord(h->ev_copied[b], h->copy_stream));
        RDP_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copied[b], 0));
        if (rd_map && k >= 2) RDP_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_mapout[b], 0));
        float* map = h->map_scratch[b];
        rc = run_front(h, h->in_stage[b], n, map, h->cpi_flags[b]);
        if (rc) return rc;
        rc = run_cfar(h, map, n, nullptr, h->dets + (size_t)base * c.max_dets_per_cpi, h->counts + base,
                      h->status + base, base, c.arith == RDP_ARITH_FIXED16, 0, h->cpi_flags[b]);
        if (rc) return rc;
        RDP_CUDA(h, cudaEventRecord(h->ev_done[b], h->stream));
        if (rd_map) {
            RDP_CUDA(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_done[b], 0));
            RDP_CUDA(h, cudaMemcpyAsync(rd_map + (size_t)base * cells, map, (size_t)n * cells * sizeof(float),
                                        cudaMemcpyDeviceToHost, h->d2h_stream));
            RDP_CUDA(h, cudaEventRecord(h->ev_mapout[b], h->d2h_stream));
        }
        // pack this chunk's records behind those of the chunks before it; counts / status to the host
        h->launches += 1;
        RDP_LAUNCH(k5_pack_dets_kernel, n, 256, 0, h->stream, h->dets, h->counts, h->packed, c.max_dets_per_cpi, 0, base);
        rc = launch_check(h, "k5_pack_dets_kernel");
        if (rc) return rc;
        RDP_CUDA(h, cudaMemcpyAsync(n_dets + base, h->counts + base, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        RDP_CUDA(h, cudaMemcpyAsync(status + base, h->status + base, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        RDP_CUDA(h, cudaEventRecord(h->ev_counts[b], h->stream));
        if (k >= 1) {
            rc = drain(prev_base, prev_n, b ^ 1);
            if (rc) return rc;
        }
        prev_base = base; prev_n = n;
    }
    rc = drain(prev_base, prev_n, (k - 1) & 1);
    if (rc) return rc;
    RDP_CUDA(h, cudaStreamSynchronize(h->d2h_stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return RDP_OK;
}

// Device detection rows [n_cpi][max_dets_per_cpi] + counts -> one packed host list (pack kernel + a single D2H copy).
int rdp_gather_dets(rdp_handle* h, const rdp_det* dets, const int32_t* n_dets, int32_t n_cpi, rdp_det* dets_host,
                    int32_t* counts_host, int64_t* total_out) {
    if (!h || !dets || !n_dets || !dets_host || !counts_host || !total_out || n_cpi < 0)
        return fail(h, RDP_E_INVALID, "rdp_gather_dets: bad argument");
    const rdp_config& c = h->cfg;
    if (n_cpi > c.max_cpis) return fail(h, RDP_E_CAPACITY, "rdp_gather_dets: %d CPIs exceed max_cpis = %d", n_cpi, c.max_cpis);
    *total_out = 0;
    if (n_cpi == 0) return RDP_OK;
    RDP_CUDA(h, cudaSetDevice(h->device));
    h->launches += 1;
    if (!h->packed) {                                             // scratch of the host path, allocated on first use
        int rc0 = dev_alloc(h, &h->packed, (size_t)c.max_cpis * c.max_dets_per_cpi);
        if (rc0) return rc0;
    }
    RDP_LAUNCH(k5_pack_dets_kernel, n_cpi, 256, 0, h->stream, dets, n_dets, h->packed, c.max_dets_per_cpi, 0, 0);
    int rc = launch_check(h, "k5_pack_dets_kernel");
    if (rc) return rc;
    RDP_CUDA(h, cudaMemcpyAsync(counts_host, n_dets, sizeof(int) * (size_t)n_cpi, cudaMemcpyDeviceToHost, h->stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    long long total = 0;
    for (int i = 0; i < n_cpi; ++i) total += counts_host[i] < c.max_dets_per_cpi ? counts_host[i] : c.max_dets_per_cpi;
    if (total > 0) {
        RDP_CUDA(h, cudaMemcpyAsync(dets_host, h->packed, sizeof(rdp_det) * (size_t)total, cudaMemcpyDeviceToHost, h->stream));
        RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    *total_out = total;
    return RDP_OK;
}

// ---- 16-bit fixed-point stage taps (bit-exact restatements of the hand-written RTL stages) ------------
int rdp_fixed_window(rdp_handle* h, const void* in, int64_t n_rows, int32_t n, void* out, uint32_t* saturated) {
    if (!h || !in || !out || n_rows < 0 || n < 2 || (n & 1)) return fail(h, RDP_E_INVALID, "rdp_fixed_window: bad argument");
    RDP_CUDA(h, cudaSetDevice(h->device));
    if (h->rom_n != n) {
        if (h->rom_dev) { cudaFree(h->rom_dev); h->rom_dev = nullptr; h->scratch_bytes -= (int64_t)sizeof(int) * h->rom_n; }
        h->rom_n = 0;
        int rc = upload_vec(h, fixed_window_rom(n), &h->rom_dev);
        if (rc) return rc;
        h->rom_n = n;
    }
    if (!h->sat_dev) { int rc = dev_alloc(h, &h->sat_dev, 1); if (rc) return rc; }
    RDP_CUDA(h, cudaMemsetAsync(h->sat_dev, 0, sizeof(uint32_t), h->stream));
    const long long total = (long long)n_rows * n;
    if (total > 0) {
        long long blocks = (total + 255) / 256;
        if (blocks > h->sm_count * 32) blocks = h->sm_count * 32;
        h->launches += 1;
        RDP_LAUNCH(kf_window_kernel, (unsigned)blocks, 256, 0, h->stream, reinterpret_cast<const uint32_t*>(in),
                   reinterpret_cast<uint32_t*>(out), h->rom_dev, n, total, h->sat_dev, total, 1u);
        int rc = launch_check(h, "kf_window_kernel");
        if (rc) return rc;
    }
    if (saturated) {
        RDP_CUDA(h, cudaMemcpyAsync(saturated, h->sat_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return RDP_OK;
}

int rdp_fixed_mti(rdp_handle* h, const void* in, int64_t n_lines, int32_t n, int32_t mode, void* out) {
    if (!h || !in || !out || n_lines < 0 || n < 1 || (mode != 0 && mode != 2 && mode != 3))
        return fail(h, RDP_E_INVALID, "rdp_fixed_mti: bad argument");
    RDP_CUDA(h, cudaSetDevice(h->device));
    const long long total = (long long)n_lines * n;
    if (total == 0) return RDP_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > h->sm_count * 32) blocks = h->sm_count * 32;
    h->launches += 1;
    RDP_LAUNCH(kf_mti_kernel, (unsigned)blocks, 256, 0, h->stream, reinterpret_cast<const uint32_t*>(in),
               reinterpret_cast<uint32_t*>(out), n, total, mode);
    return launch_check(h, "kf_mti_kernel");
}

int rdp_fixed_magnitude(rdp_handle* h, const void* in, int64_t n_words, uint32_t* out) {
    if (!h || !in || !out || n_words < 0) return fail(h, RDP_E_INVALID, "rdp_fixed_magnitude: bad argument");
    RDP_CUDA(h, cudaSetDevice(h->device));
    if (n_words == 0) return RDP_OK;
    long long blocks = (n_words + 255) / 256;
    if (blocks > h->sm_count * 32) blocks = h->sm_count * 32;
    h->launches += 1;
    RDP_LAUNCH(kf_magnitude_kernel, (unsigned)blocks, 256, 0, h->stream, reinterpret_cast<const uint32_t*>(in), out,
               (long long)n_words);
    return launch_check(h, "kf_magnitude_kernel");
}

int rdp_fixed_corner_turn(rdp_handle* h, const void* in, int32_t n_maps, int32_t n_chirps, int32_t n_samples, void* out) {
    if (!h || !in || !out || n_maps < 0 || n_maps > 65535 || n_chirps < 1 || n_samples < 1)
        return fail(h, RDP_E_INVALID, "rdp_fixed_corner_turn: bad argument");
    RDP_CUDA(h, cudaSetDevice(h->device));
    if (n_maps == 0) return RDP_OK;
    h->launches += 1;
    RDP_LAUNCH(kf_corner_turn_kernel, dim3((n_samples + 31) / 32, (n_chirps + 31) / 32, n_maps), 256, 0, h->stream,
               reinterpret_cast<const uint32_t*>(in), reinterpret_cast<uint32_t*>(out), n_chirps, n_samples);
    return launch_check(h, "kf_corner_turn_kernel");
}

int rdp_fixed_bfp_fft(rdp_handle* h, const void* in, int64_t n_frames, int32_t n, void* out, int32_t* exps) {
    if (!h || !in || !out || n_frames < 0 || n < 2 || n > 4096 || !is_pow2(n))
        return fail(h, RDP_E_INVALID, "rdp_fixed_bfp_fft: bad argument (n must be a power of two in [2, 4096])");
    RDP_CUDA(h, cudaSetDevice(h->device));
    if (h->btw_tap_n != n) {
        if (h->btw_tap) { cudaFree(h->btw_tap); h->btw_tap = nullptr; h->scratch_bytes -= (int64_t)sizeof(uint32_t) * (h->btw_tap_n / 2); }
        h->btw_tap_n = 0;
        int rc = upload_vec(h, bfp_twiddle_words(n), &h->btw_tap);
        if (rc) return rc;
        h->btw_tap_n = n;
    }
    BfpFuse fu;
    std::memset(&fu, 0, sizeof(fu));
    return launch_bfp(h, reinterpret_cast<const uint32_t*>(in), reinterpret_cast<uint32_t*>(out), exps, h->btw_tap, n,
                      (long long)n_frames, fu);
}

int rdp_tracker_reset(rdp_handle* h, const rdp_tracker_config* cfg) {
    if (!h) return RDP_E_INVALID;
    rdp_tracker_config d = {32, 2, 5, 10, 5, 128, 64};          // radar_core.vhd:425-429
    if (cfg) d = *cfg;
    if (d.max_tracks < 1 || d.max_tracks > 32 || d.init_hits < 0 || d.coast_max < 0 || d.assoc_gate_r < 0 || d.assoc_gate_d < 0)
        return fail(h, RDP_E_INVALID, "rdp_tracker_reset: bad generics (max_tracks must be 1..32)");
    RDP_CUDA(h, cudaSetDevice(h->device));
    const int n_ch = h->cfg.n_channels;
    if (!h->trk_state) { int rc = dev_alloc(h, &h->trk_state, (size_t)n_ch); if (rc) return rc; }
    std::vector<TrackState> init((size_t)n_ch);
    std::memset(init.data(), 0, sizeof(TrackState) * (size_t)n_ch);
    for (auto& s : init) { s.best_det_idx = 63; s.best_distance = 0xffff; }      // signal initialisers, tws_tracker.vhd:81-83
    RDP_CUDA(h, cudaMemcpyAsync(h->trk_state, init.data(), sizeof(TrackState) * (size_t)n_ch, cudaMemcpyHostToDevice, h->stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    h->trk_args = TrackerArgs{d.max_tracks, d.init_hits, d.coast_max, d.assoc_gate_r, d.assoc_gate_d, d.alpha_gain, d.beta_gain,
                              n_ch, h->cfg.max_dets_per_cpi};
    return RDP_OK;
}

int rdp_track_device(rdp_handle* h, const rdp_det* dets, const int32_t* n_dets, int32_t n_cpi, rdp_track* tracks,
                     int32_t* n_tracks, int32_t* active) {
    if (!h || !dets || !n_dets || !tracks || !n_tracks || !active || n_cpi < 0)
        return fail(h, RDP_E_INVALID, "rdp_track_device: bad argument");
    if (!h->trk_state) { int rc = rdp_tracker_reset(h, nullptr); if (rc) return rc; }
    if (n_cpi == 0) return RDP_OK;
    RDP_CUDA(h, cudaSetDevice(h->device));
    h->launches += 1;
    RDP_LAUNCH(k6_tracker_kernel, h->cfg.n_channels, 32, 0, h->stream, dets, n_dets, n_cpi, h->trk_args, h->trk_state, tracks,
               n_tracks, active, 0);
    return launch_check(h, "k6_tracker_kernel");
}

// ---- sharded large CPI (BASELINE config 4): peer-store corner turn + row-count overrides + CUDA IPC -------------
int rdp_range_stage_scatter(rdp_handle* h, const void* cube, int32_t n_cpi, const rdp_scatter* sc) {
    if (!h || !cube || !sc || n_cpi < 0 || sc->world < 1 || sc->world > 16 || sc->rank < 0 || sc->rank >= sc->world)
        return fail(h, RDP_E_INVALID, "rdp_range_stage_scatter: bad argument");
    const rdp_config& c = h->cfg;
    if (c.combine != RDP_COMBINE_NONE || c.arith != RDP_ARITH_FLOAT32)
        return fail(h, RDP_E_INVALID, "the sharded-CPI entry points support neither channel combining nor arith FIXED16");
    if (c.n_samples % sc->world != 0 || !is_pow2(c.n_samples / sc->world))
        return fail(h, RDP_E_INVALID, "n_samples / world must be a power of two");
    K1Scatter k;
    std::memset(&k, 0, sizeof(k));
    for (int i = 0; i < sc->world; ++i) {
        if (!sc->peer[i]) return fail(h, RDP_E_INVALID, "rdp_range_stage_scatter: peer[%d] is NULL", i);
        k.peer[i] = reinterpret_cast<float2*>(sc->peer[i]);
    }
    k.world = sc->world;
    k.rank = sc->rank;
    k.rows_shift = 0;
    while ((1 << k.rows_shift) < c.n_samples / sc->world) ++k.rows_shift;
    k.n_chirps_total = c.n_chirps * sc->world;
    RDP_CUDA(h, cudaSetDevice(h->device));
    if (sc->world == 1) return run_range(h, cube, n_cpi, k.peer[0]);
    return run_range(h, cube, n_cpi, k.peer[sc->rank], &k);
}

int rdp_doppler_stage_rows(rdp_handle* h, const void* x_rc, int64_t n_rows, int32_t rows_per_map, int32_t pad_rows,
                           float* rd_map) {
    if (!h || !x_rc || !rd_map || n_rows < 0 || rows_per_map < 0 || pad_rows < 0)
        return fail(h, RDP_E_INVALID, "rdp_doppler_stage_rows: bad argument");
    RDP_CUDA(h, cudaSetDevice(h->device));
    return run_doppler(h, reinterpret_cast<const float2*>(x_rc), 0, rd_map, n_rows, rows_per_map, pad_rows);
}

int rdp_cfar_stage_rows(rdp_handle* h, const float* rd_map, int32_t n_cpi, int32_t n_range_rows, float* cfar_out,
                        rdp_det* dets, int32_t* n_dets, uint32_t* status) {
    if (!h || !rd_map || !dets || !n_dets || !status || n_cpi < 0 || n_range_rows < 1)
        return fail(h, RDP_E_INVALID, "rdp_cfar_stage_rows: bad argument");
    const rdp_config& c = h->cfg;
    if (c.combine != RDP_COMBINE_NONE) return fail(h, RDP_E_INVALID, "rdp_cfar_stage_rows: not with channel combining");
    RDP_CUDA(h, cudaSetDevice(h->device));
    const size_t cells = (size_t)c.n_channels * n_range_rows * c.n_chirps;
    for (int base = 0; base < n_cpi; base += h->chunk) {
        const int n = (n_cpi - base) < h->chunk ? (n_cpi - base) : h->chunk;
        int rc = run_cfar(h, rd_map + (size_t)base * cells, n, cfar_out ? cfar_out + (size_t)base * cells : nullptr,
                          dets + (size_t)base * c.max_dets_per_cpi, n_dets + base, status + base, base, 0, n_range_rows);
        if (rc) return rc;
    }
    return RDP_OK;
}

int rdp_ipc_export(rdp_handle* h, void* device_ptr, unsigned char handle_bytes[64]) {
    if (!h || !device_ptr || !handle_bytes) return RDP_E_INVALID;
#ifdef RDP_EMU
    std::memset(handle_bytes, 0, 64);
    std::memcpy(handle_bytes, &device_ptr, sizeof(void*));       // single-process emulation: the pointer itself
    return RDP_OK;
#else
    RDP_CUDA(h, cudaSetDevice(h->device));
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    RDP_CUDA(h, cudaIpcGetMemHandle(&mh, device_ptr));
    std::memcpy(handle_bytes, &mh, 64);
    return RDP_OK;
#endif
}

int rdp_ipc_open(rdp_handle* h, const unsigned char handle_bytes[64], void** device_ptr) {
    if (!h || !handle_bytes || !device_ptr) return RDP_E_INVALID;
#ifdef RDP_EMU
    std::memcpy(device_ptr, handle_bytes, sizeof(void*));
    return RDP_OK;
#else
    RDP_CUDA(h, cudaSetDevice(h->device));
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, handle_bytes, 64);
    RDP_CUDA(h, cudaIpcOpenMemHandle(device_ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    return RDP_OK;
#endif
}

int rdp_ipc_close(rdp_handle* h, void* device_ptr) {
    if (!h || !device_ptr) return RDP_E_INVALID;
#ifndef RDP_EMU
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaIpcCloseMemHandle(device_ptr));
#endif
    return RDP_OK;
}

int rdp_copy_d2d(rdp_handle* h, void* dst, const void* src, size_t bytes) {
    if (!h || (!dst && bytes) || (!src && bytes)) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return RDP_OK;
}

int rdp_copy_d2d_rows(rdp_handle* h, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes,
                      size_t n_rows) {
    if (!h || !dst || !src || row_bytes > dst_pitch || row_bytes > src_pitch) return RDP_E_INVALID;
    if (!row_bytes || !n_rows) return RDP_OK;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, n_rows, cudaMemcpyDeviceToDevice, h->stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return RDP_OK;
}

int rdp_device_memset(rdp_handle* h, void* dst, int value, size_t bytes) {
    if (!h || (!dst && bytes)) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaMemsetAsync(dst, value, bytes, h->stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return RDP_OK;
}

int rdp_stream_open(rdp_handle* h, int32_t n_slots, int32_t slot_cpis) {
    if (!h || n_slots < 1 || n_slots > 16 || slot_cpis < 1) return fail(h, RDP_E_INVALID, "rdp_stream_open: bad argument");
    if (!h->slots.empty()) return fail(h, RDP_E_INVALID, "rdp_stream_open: stream already open");
    if (slot_cpis > h->chunk) return fail(h, RDP_E_CAPACITY, "slot_cpis %d exceeds the chunk size %d", slot_cpis, h->chunk);
    const rdp_config& c = h->cfg;
    RDP_CUDA(h, cudaSetDevice(h->device));
    h->slot_cpis = slot_cpis;
    h->slots.resize(n_slots);
    const size_t in_bytes = (size_t)slot_cpis * cells_per_cpi(c) * in_bytes_per_sample(c);
    const size_t n_rec = (size_t)slot_cpis * c.max_dets_per_cpi;
    for (auto& sl : h->slots) {
        RDP_CUDA(h, cudaStreamCreateWithFlags(&sl.copy, cudaStreamNonBlocking));
        RDP_CUDA(h, cudaEventCreateWithFlags(&sl.ev_in, cudaEventDefault));
        RDP_CUDA(h, cudaEventCreateWithFlags(&sl.ev_done, cudaEventDefault));
        RDP_CUDA(h, cudaEventCreateWithFlags(&sl.ev_out, cudaEventDefault));
        unsigned char* p = nullptr;
        int rc = dev_alloc(h, &p, in_bytes); if (rc) return rc;
        sl.d_in = p;
        rc = dev_alloc(h, &sl.d_dets, n_rec); if (rc) return rc;
        rc = dev_alloc(h, &sl.d_packed, n_rec); if (rc) return rc;
        rc = dev_alloc(h, &sl.d_counts, (size_t)slot_cpis); if (rc) return rc;
        rc = dev_alloc(h, &sl.d_status, (size_t)slot_cpis); if (rc) return rc;
        RDP_CUDA(h, cudaHostAlloc(reinterpret_cast<void**>(&sl.h_packed), n_rec * sizeof(rdp_det), cudaHostAllocDefault));
        RDP_CUDA(h, cudaHostAlloc(reinterpret_cast<void**>(&sl.h_counts), (size_t)slot_cpis * sizeof(int), cudaHostAllocDefault));
        RDP_CUDA(h, cudaHostAlloc(reinterpret_cast<void**>(&sl.h_status), (size_t)slot_cpis * sizeof(uint32_t), cudaHostAllocDefault));
    }
    return RDP_OK;
}

int rdp_stream_submit(rdp_handle* h, int32_t slot, const void* cube, int32_t n_cpi) {
    if (!h || !cube || slot < 0 || slot >= (int)h->slots.size() || n_cpi < 1 || n_cpi > h->slot_cpis)
        return fail(h, RDP_E_INVALID, "rdp_stream_submit: bad argument");
    StreamSlot& sl = h->slots[slot];
    if (sl.busy) return fail(h, RDP_E_INVALID, "rdp_stream_submit: slot %d still in flight (call rdp_stream_wait)", slot);
    const rdp_config& c = h->cfg;
    RDP_CUDA(h, cudaSetDevice(h->device));
    const size_t in_bytes = (size_t)n_cpi * cells_per_cpi(c) * in_bytes_per_sample(c);
    // ingest on the slot's copy stream; the kernels of all slots share the compute stream (and its scratch), so
    // the copy of submission i+1 overlaps the kernels of submission i
    RDP_CUDA(h, cudaMemcpyAsync(sl.d_in, cube, in_bytes, cudaMemcpyHostToDevice, sl.copy));
    RDP_CUDA(h, cudaEventRecord(sl.ev_in, sl.copy));
    RDP_CUDA(h, cudaStreamWaitEvent(h->stream, sl.ev_in, 0));
    // the kernel chain of one submission: front end, CFAR, compaction, pack -- all on the compute stream
    auto chain = [&]() -> int {
        int rc0 = run_front(h, sl.d_in, n_cpi, h->map_scratch[0], h->cpi_flags[0]);
        if (rc0) return rc0;
        rc0 = run_cfar(h, h->map_scratch[0], n_cpi, nullptr, sl.d_dets, sl.d_counts, sl.d_status, 0, c.arith == RDP_ARITH_FIXED16, 0,
                       h->cpi_flags[0]);
        if (rc0) return rc0;
        h->launches += 1;
        RDP_LAUNCH(k5_pack_dets_kernel, n_cpi, 256, 0, h->stream, sl.d_dets, sl.d_counts, sl.d_packed, c.max_dets_per_cpi, 0, 0);
        return launch_check(h, "k5_pack_dets_kernel");
    };
    int rc = RDP_OK;
#ifndef RDP_EMU
    // Every buffer of the chain belongs to the slot or the handle, so the chain is the same from submission to
    // submission: it is captured into a CUDA graph the first time a slot sees an n_cpi and replayed afterwards -- one
    // launch instead of seven, tensor maps and launch parameters baked in (RDP_STREAM_GRAPH=0 turns it off; per-kernel
    // stage timing needs the individual launches).
    if (h->knob_stream_graph && !h->stage_timing) {
        if (!sl.graph || sl.graph_n_cpi != n_cpi) {
            if (sl.graph) { cudaGraphExecDestroy(sl.graph); sl.graph = nullptr; }
            const int64_t launches0 = h->launches;
            cudaGraph_t g = nullptr;
            RDP_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
            rc = chain();
            cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
            if (rc == RDP_OK && ce == cudaSuccess && g) ce = cudaGraphInstantiate(&sl.graph, g, 0);
            if (g) cudaGraphDestroy(g);
            h->launches = launches0;                                // counted per replay below
            if (rc != RDP_OK || ce != cudaSuccess) {                 // fall back to plain launches for good
                cudaGetLastError();
                sl.graph = nullptr;
                h->knob_stream_graph = 0;
                rc = RDP_OK;
            } else {
                sl.graph_n_cpi = n_cpi;
            }
        }
    }
    if (h->knob_stream_graph && !h->stage_timing && sl.graph) {
        RDP_CUDA(h, cudaGraphLaunch(sl.graph, h->stream));
        h->launches += (c.arith == RDP_ARITH_FIXED16 ? 3 : 2) + 3 + 1 + (c.peak_group ? 1 : 0);   // front end, K3 + 2 x K4, K5, K3b
    } else
#endif
    rc = chain();
    if (rc) return rc;
    RDP_CUDA(h, cudaEventRecord(sl.ev_done, h->stream));
    RDP_CUDA(h, cudaStreamWaitEvent(sl.copy, sl.ev_done, 0));
    RDP_CUDA(h, cudaMemcpyAsync(sl.h_counts, sl.d_counts, sizeof(int) * (size_t)n_cpi, cudaMemcpyDeviceToHost, sl.copy));
    RDP_CUDA(h, cudaMemcpyAsync(sl.h_status, sl.d_status, sizeof(uint32_t) * (size_t)n_cpi, cudaMemcpyDeviceToHost, sl.copy));
    // the head of the packed list is copied speculatively together with the counts, so that one synchronisation
    // suffices for ordinary scenes; rdp_stream_wait fetches the rest of an unusually long list
    {
        size_t spec = (size_t)n_cpi * c.max_dets_per_cpi;
        const int per_cpi = h->stream_spec_dets;                  // RDP_STREAM_SPEC at create (tests shrink it)
        if (spec > (size_t)n_cpi * per_cpi) spec = (size_t)n_cpi * per_cpi;
        sl.spec = (long long)spec;
        RDP_CUDA(h, cudaMemcpyAsync(sl.h_packed, sl.d_packed, sizeof(rdp_det) * spec, cudaMemcpyDeviceToHost, sl.copy));
    }
    RDP_CUDA(h, cudaEventRecord(sl.ev_out, sl.copy));
    sl.n_cpi = n_cpi;
    sl.busy = true;
    return RDP_OK;
}

int rdp_stream_wait(rdp_handle* h, int32_t slot, const rdp_det** dets, const int32_t** n_dets, const uint32_t** status,
                    int32_t* total) {
    if (!h || slot < 0 || slot >= (int)h->slots.size()) return fail(h, RDP_E_INVALID, "rdp_stream_wait: bad argument");
    StreamSlot& sl = h->slots[slot];
    if (!sl.busy) return fail(h, RDP_E_INVALID, "rdp_stream_wait: slot %d has no submission in flight", slot);
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaEventSynchronize(sl.ev_out));
    sl.busy = false;
    long long t = 0;
    for (int i = 0; i < sl.n_cpi; ++i)
        t += sl.h_counts[i] < h->cfg.max_dets_per_cpi ? sl.h_counts[i] : h->cfg.max_dets_per_cpi;
    if (t > sl.spec) {
        RDP_CUDA(h, cudaMemcpyAsync(sl.h_packed + sl.spec, sl.d_packed + sl.spec, sizeof(rdp_det) * (size_t)(t - sl.spec),
                                    cudaMemcpyDeviceToHost, sl.copy));
        RDP_CUDA(h, cudaStreamSynchronize(sl.copy));
    }
    if (dets) *dets = sl.h_packed;
    if (n_dets) *n_dets = sl.h_counts;
    if (status) *status = sl.h_status;
    if (total) *total = (int32_t)t;
    return RDP_OK;
}

int rdp_stream_close(rdp_handle* h) {
    if (!h) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaDeviceSynchronize());
    for (auto& sl : h->slots) {
        void* dev[] = {sl.d_in, sl.d_dets, sl.d_packed, sl.d_counts, sl.d_status};
        for (void* p : dev) if (p) cudaFree(p);
        void* hst[] = {sl.h_packed, sl.h_counts, sl.h_status};
        for (void* p : hst) if (p) cudaFreeHost(p);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.ev_out) cudaEventDestroy(sl.ev_out);
        if (sl.copy) cudaStreamDestroy(sl.copy);
#ifndef RDP_EMU
        if (sl.graph) cudaGraphExecDestroy(sl.graph);
#endif
    }
    h->slots.clear();
    return RDP_OK;
}

int rdp_device_alloc(rdp_handle* h, size_t bytes, void** ptr) {
    if (!h || !ptr) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaMalloc(ptr, bytes ? bytes : 16));
    return RDP_OK;
}
int rdp_device_free(rdp_handle* h, void* ptr) {
    if (!h) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaFree(ptr));
    return RDP_OK;
}
int rdp_host_alloc(rdp_handle* h, size_t bytes, void** ptr) {
    if (!h || !ptr) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return RDP_OK;
}
int rdp_host_free(rdp_handle* h, void* ptr) {
    if (!h) return RDP_E_INVALID;
    RDP_CUDA(h, cudaFreeHost(ptr));
    return RDP_OK;
}
int rdp_copy_h2d(rdp_handle* h, void* dst, const void* src, size_t bytes) {
    if (!h || (!dst && bytes) || (!src && bytes)) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return RDP_OK;
}
int rdp_copy_d2h(rdp_handle* h, void* dst, const void* src, size_t bytes) {
    if (!h || (!dst && bytes) || (!src && bytes)) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    return RDP_OK;
}

int rdp_timer_start(rdp_handle* h) {
    if (!h) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaEventRecord(h->t0, h->stream));
    return RDP_OK;
}
int rdp_timer_stop(rdp_handle* h, float* elapsed_ms) {
    if (!h || !elapsed_ms) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaEventRecord(h->t1, h->stream));
    RDP_CUDA(h, cudaEventSynchronize(h->t1));
    RDP_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->t0, h->t1));
    return RDP_OK;
}
int rdp_set_stage_timing(rdp_handle* h, int32_t enable) {
    if (!h) return RDP_E_INVALID;
    h->stage_timing = enable != 0;
    return RDP_OK;
}
int rdp_get_stage_timing(rdp_handle* h, float ms[4], int64_t launches[4]) {
    if (!h || !ms || !launches) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    RDP_CUDA(h, cudaStreamSynchronize(h->stream));
    for (auto& se : h->stage_events) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, se.a, se.b) == cudaSuccess) h->stage_ms[se.stage] += t;
        cudaEventDestroy(se.a);
        cudaEventDestroy(se.b);
    }
    h->stage_events.clear();
    for (int i = 0; i < 4; ++i) {
        ms[i] = h->stage_ms[i];
        launches[i] = h->stage_launches[i];
        h->stage_ms[i] = 0.f;
        h->stage_launches[i] = 0;
    }
    return RDP_OK;
}
int rdp_flush_l2(rdp_handle* h) {
    if (!h) return RDP_E_INVALID;
    RDP_CUDA(h, cudaSetDevice(h->device));
    if (!h->l2_flush) {
        cudaDeviceProp prop;
        RDP_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
        h->l2_flush_bytes = (size_t)prop.l2CacheSize * 2 + (64u << 20);
        RDP_CUDA(h, cudaMalloc(&h->l2_flush, h->l2_flush_bytes));
    }
    RDP_CUDA(h, cudaMemsetAsync(h->l2_flush, 0x5a, h->l2_flush_bytes, h->stream));
    return RDP_OK;
}

}  // extern "C"
