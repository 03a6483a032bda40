/*
 * rdp.h -- C ABI of the B200-native range-Doppler + OS-CFAR path ("librdp.so").
 *
 * This is the drop-in boundary for the hot path of the reference FPGA design
 * Aurellia-Beam/fpga-fmcw-radar-processor.  The reference has no software API; the
 * boundary it defines is the `radar_core` entity (rtl/src/radar_core.vhd:11-57:
 * generics, ADC stream in, detection stream out, control pins, status flag) and the
 * text records its testbenches write (rtl/src/tb_radar_core.vhd:149-160).  Each entry
 * point below cites the part of that interface it replaces.  Plain C: pointers, sizes
 * and PODs only -- no CUDA, torch or C++ types appear in any signature.
 *
 * Threading: a handle is bound to one CUDA device and is not thread-safe; use one
 * handle per GPU per host thread (the reference is a single in-order clock domain).
 * All entry points return RDP_OK (0) or a negative RDP_E* code and never throw or
 * exit; rdp_last_error() returns the message of the last failure.
 * There is NO CPU fallback: rdp_create() fails with RDP_E_NO_DEVICE without a GPU.
 */
#ifndef RDP_H_
#define RDP_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDP_ABI_VERSION 2

enum {
    RDP_OK = 0,
    RDP_E_INVALID = -1,      /* bad argument / unsupported configuration          */
    RDP_E_NO_DEVICE = -2,    /* no CUDA device, or device index out of range       */
    RDP_E_CUDA = -3,         /* a CUDA runtime call or kernel launch failed        */
    RDP_E_NOMEM = -4,        /* device or pinned-host allocation failed            */
    RDP_E_CAPACITY = -5      /* n_cpi exceeds the handle's configured capacity     */
};

/* input_format */
enum { RDP_IN_INT16_IQ = 0, RDP_IN_COMPLEX64 = 1 };
/* mag_mode: RDP_MAG_AMB is the reference's alpha-max-beta-min (rtl/src/magnitude_calc.vhd:55-81) */
enum { RDP_MAG_AMB = 0, RDP_MAG_ABS = 1, RDP_MAG_POW = 2, RDP_MAG_LOGPOW = 3 };
/* RDP_MAG_LOGPOW writes 10 log10(|X|^2) dB (a null cell = -inf is clamped to -300 dB).  The CFAR stage keeps the
 * reference's MULTIPLICATIVE rule cut > scale * x_k on whatever words the map holds (os_cfar_2d.vhd:200-207 knows
 * nothing of the map's unit): on a dB map that is a well-defined but unusual statistic, and where the map has
 * negative words the tile is decided by the exact sign-safe path (slower, identical to the oracle).  A CUT equal to
 * zero is never a detection (the RTL's output word would be 0 either way).                                        */
/* arith: FLOAT32 = windows / FFTs / magnitude in float32 (the reference's FFT IP arithmetic is not defined in its
 * tree); FIXED16 = the 16-bit word-for-word chain of radar_core with the two FFT IP cores replaced by the
 * self-defined block-floating-point transform of rdp_fixed_bfp_fft ("self-consistent, not Xilinx-exact") and the
 * integer CFAR of rdp_cfar_stage_u32.  FIXED16 needs int16 input, mag_mode AMB and window_gain 1.0 (the as-built
 * x2 gain is part of the fixed-point window itself).                                                             */
enum { RDP_ARITH_FLOAT32 = 0, RDP_ARITH_FIXED16 = 1 };
/* combine (optional stage in front of the CFAR; NOT in the reference, which has one receive channel -- SURVEY.md 8f
 * row 3): NONE = one range-Doppler map and one CFAR per channel; NCI = non-coherent integration, the magnitudes (in
 * the configured mag_mode; RDP_MAG_POW gives the classical power integration) of the n_channels channels of a CPI
 * are added cell by cell (float32, channel 0 first) inside the Doppler stage and the CFAR runs on the ONE combined
 * map: rd_map is then [n_cpi][1][range][doppler] and every record has channel = 0.  FLOAT32 arithmetic only.
 * BEAMS = coherent digital beamforming of a uniform linear array: n_beams = the smallest power of two >= n_channels,
 * beam b = sum over channels c of x_c * exp(-j 2 pi b c / n_beams), formed on the range-major intermediate inside the
 * Doppler stage (the transform is linear: steering before the Doppler FFT equals an angle FFT across the channels after
 * it); rd_map is [n_cpi][n_beams][range][doppler], the CFAR runs per beam and every record's channel = its beam.   */
enum { RDP_COMBINE_NONE = 0, RDP_COMBINE_NCI = 1, RDP_COMBINE_BEAMS = 2 };
/* cfar_axes: AS_BUILT = the RTL's *_RANGE generics act along Doppler and *_DOPPLER along
 * range (rtl/src/os_cfar_2d.vhd:140-146,157-159); AS_NAMED = what the names say.        */
enum { RDP_AXES_AS_BUILT = 0, RDP_AXES_AS_NAMED = 1 };
/* cfar_edges: ZERO = window clipped to the CPI, outside cells read 0; RASTER = window on
 * the flat range-major raster exactly as the RTL line buffer sees it (wraps into the
 * neighbouring range line at Doppler edges; 0 before / after the CPI).                   */
enum { RDP_EDGES_ZERO = 0, RDP_EDGES_RASTER = 1 };

/* status word bits (one word per CPI) -- the sticky `status_overflow` of
 * rtl/src/radar_core.vhd:447-456 split by cause.                                         */
enum {
    RDP_STATUS_DET_OVERFLOW = 1u << 0,  /* more than max_dets_per_cpi detections: the list holds the
                                           first max_dets_per_cpi in (channel, range, doppler) order */
    RDP_STATUS_WINDOW_SAT = 1u << 1,    /* FIXED16: a window_multiplier output clipped to 16 bits somewhere in the
                                           CPI (saturation_flag, rtl/src/window_multiplier.vhd:151-158, which
                                           radar_core.vhd:447-456 ORs into status_overflow)                  */
    RDP_STATUS_ADC_FULL_SCALE = 1u << 2,/* int16 input: a sample of the CPI sits on a rail (+32767 / -32768), i.e.
                                           the ADC (or the synthetic generator) clipped before the path saw it */
    RDP_STATUS_MAP_RANGE = 1u << 3      /* rdp_cfar_stage_u32: a magnitude word >= 2^24 was seen -- outside the
                                           supported range, detections of this CPI may be missing               */
};

/*
 * Parameters = the generics of radar_core (rtl/src/radar_core.vhd:12-20), the fixed
 * generics it passes to os_cfar_2d (:376-382) and doppler_notch (:329-331), and its two
 * run-time control pins mti_bypass / cfar_scale_ovr (:48-49).
 */
typedef struct rdp_config {
    int32_t struct_size;       /* = sizeof(rdp_config); guards ABI drift                  */
    int32_t n_chirps;          /* N_DOPPLER: chirps per CPI = Doppler FFT length (8..4096, power of 2) */
    int32_t n_samples;         /* N_RANGE: samples per chirp = range FFT length (8..4096, power of 2)  */
    int32_t n_channels;        /* receive channels per CPI (the RTL has 1)                */
    int32_t cfar_ref_r;        /* CFAR_REF_R   (default 4)                                */
    int32_t cfar_ref_d;        /* CFAR_REF_D   (default 4)                                */
    int32_t cfar_guard_r;      /* CFAR_GUARD_R (default 2)                                */
    int32_t cfar_guard_d;      /* CFAR_GUARD_D (default 1)                                */
    int32_t rank_pct;          /* RANK_PCT     (75)                                       */
    int32_t scale_min;         /* SCALE_MIN    (2)                                        */
    int32_t scale_nom;         /* SCALE_NOM    (4)                                        */
    int32_t scale_max;         /* SCALE_MAX    (6)                                        */
    int32_t scale_override;    /* cfar_scale_ovr: 0 = adaptive, 1..7 = fixed scale        */
    int32_t mti_mode;          /* 0 = mti_bypass, 2 = two-pulse (NOTCH_MODE), 3 = three-pulse */
    int32_t input_format;      /* RDP_IN_*                                                */
    int32_t mag_mode;          /* RDP_MAG_*                                               */
    int32_t cfar_axes;         /* RDP_AXES_*                                              */
    int32_t cfar_edges;        /* RDP_EDGES_*                                             */
    float   window_gain;       /* 1.0; 2.0 mirrors the as-built x2 of window_multiplier.vhd:146-149 */
    int32_t max_dets_per_cpi;  /* detection-list capacity per CPI                         */
    int32_t max_cpis;          /* largest n_cpi a single rdp_process* call may pass       */
    int32_t chunk_cpis;        /* CPIs per internal pass (0 = auto: 1.5 GiB of range-major intermediate for
                                  device-resident cubes, 384 MiB sub-chunks on the host path: the measured
                                  optima on B200)                                          */
    int32_t arith;             /* RDP_ARITH_*                                             */
    int32_t combine;           /* RDP_COMBINE_*                                           */
    int32_t peak_group;        /* 1 = peak grouping of the detections (optional, not in the reference): a detection
                                  is dropped when one of its 8 neighbouring cells is a detection of larger
                                  magnitude (ties: the first in (range, doppler) order stays), so a target that
                                  spreads over adjacent cells is reported once, at its peak.  Counts, overflow
                                  and the list refer to the grouped detections; cfar_out stays the CFAR output. */
} rdp_config;

/* One detection = one `r d mag` record (rtl/src/tb_radar_core.vhd:153-159), with the CPI
 * and channel it belongs to and TRUE cell coordinates (the RTL's 3-beat label skew,
 * radar_core.vhd:396-411, is not reproduced).                                            */
typedef struct rdp_det {
    int32_t cpi;
    int32_t channel;
    int32_t range;
    int32_t doppler;
    float   mag;
} rdp_det;

/* One reported track = one `TRK <id> R=<q2> D=<q2> Q=<quality>` line (rtl/src/tb_radar_core.vhd:163-181):
 * the trk_* outputs of rtl/src/tws_tracker.vhd:30-37; positions are Q2 (bins * 4) signed as in the RTL. */
typedef struct rdp_track {
    int32_t cpi;
    int32_t channel;
    int32_t id;
    int32_t range_q2;
    int32_t doppler_q2;
    int32_t vel_r;
    int32_t vel_d;
    int32_t quality;
    int32_t status;            /* 2 = FIRM, 3 = COAST (only these are reported, tws_tracker.vhd:280-281) */
} rdp_track;

/* generics of tws_tracker (rtl/src/tws_tracker.vhd:10-20) as radar_core sets them (radar_core.vhd:425-429) */
typedef struct rdp_tracker_config {
    int32_t max_tracks;        /* MAX_TRACKS (<= 32)          default 32  */
    int32_t init_hits;         /* INIT_HITS                   default 2   */
    int32_t coast_max;         /* COAST_MAX                   default 5   */
    int32_t assoc_gate_r;      /* ASSOC_GATE_R (bins)         default 10  */
    int32_t assoc_gate_d;      /* ASSOC_GATE_D (bins)         default 5   */
    int32_t alpha_gain;        /* ALPHA_GAIN (Q8)             default 128 */
    int32_t beta_gain;         /* BETA_GAIN (Q8)              default 64  */
} rdp_tracker_config;

/* Destination of the distributed corner turn (rdp_range_stage_scatter): the range-major intermediate of every
 * rank, [map][n_samples / world][n_chirps * world] complex64, as device pointers valid in THIS process (the
 * rank's own buffer and the peers' buffers opened with rdp_ipc_open).                                      */
typedef struct rdp_scatter {
    void*   peer[16];
    int32_t world;
    int32_t rank;
} rdp_scatter;

typedef struct rdp_info {
    int32_t abi_version;
    int32_t device;
    int32_t sm_count;
    int32_t chunk_cpis;        /* resolved chunk size                                     */
    int32_t cfar_tma;          /* 1 if the last CFAR launch staged its tiles with TMA     */
    int32_t reserved;
    int64_t scratch_bytes;     /* device memory owned by the handle                       */
    int64_t kernel_launches;   /* kernels launched by this handle since creation          */
    char    device_name[128];
} rdp_info;

/* Environment variables read by rdp_create (measurement knobs; the defaults are what the benchmarks use):
 *   RDP_OVERLAP=0|1       CFAR kernels of chunk i on a second stream under the FFT kernels of chunk i+1 (default 0:
 *                         since K3 balances its tiles through a queue there is no tail left to fill, measured -1.7 %)
 *   RDP_K3_CTAS=n         cap the persistent CFAR CTAs per SM;  RDP_K12_CTAS=n  the same for the two FFT kernels
 *   RDP_K1_CLUSTER=0|1    force the thread-block-cluster variant of the range stage off / on (default: only for the
 *                         peer-store corner turn of rdp_range_stage_scatter with 2048- / 4096-sample chirps)
 *   RDP_STREAM_SPEC=n     records per CPI a streaming slot copies back together with the counts (tests)
 *   RDP_STREAM_GRAPH=0|1  streaming submissions replay a CUDA graph of their kernel chain (default 1)
 *   RDP_K3_QUEUE=0|1      the persistent CFAR CTAs draw their tiles from a device-side queue (default 1) / fixed stride
 *   RDP_K1_TMA=0|1        force the TMA tile store of the range stage off / on (default: on for int16 input)       */
typedef struct rdp_handle rdp_handle;

/* Fills *cfg with the reference defaults (radar_core.vhd:12-20: 1024 x 128, CFAR 4/4/2/1,
 * rank 75 %, scales 2/4/6, adaptive scale, MTI bypassed, int16 IQ input).                */
void rdp_default_config(rdp_config* cfg);

/* Replaces reset + the one-shot FFT configuration FSM (radar_core.vhd:279-301): builds
 * window / twiddle tables, allocates scratch and the stream.                              */
int rdp_create(const rdp_config* cfg, int32_t device, rdp_handle** out);
int rdp_destroy(rdp_handle* h);
const char* rdp_last_error(const rdp_handle* h);   /* h may be NULL: last create error  */
int rdp_get_info(const rdp_handle* h, rdp_info* info);

/*
 * The whole path on DEVICE-resident buffers (the s_axis ADC stream in, the det_* stream
 * out: radar_core.vhd:25-35).  Asynchronous on the handle's stream; rdp_sync() waits.
 *   cube     [n_cpi][channel][chirp][sample] of {I,Q} int16 (I first) or complex64
 *   rd_map   [n_cpi][channel][range][doppler] float32, or NULL to keep it internal
 *   dets     [n_cpi][max_dets_per_cpi] records; the first n_dets[i] of row i are valid and
 *            sorted by (channel, range, doppler)
 *   n_dets   [n_cpi] detections found per CPI (may exceed max_dets_per_cpi => overflow bit)
 *   status   [n_cpi] RDP_STATUS_* bits
 */
int rdp_process_device(rdp_handle* h, const void* cube, int32_t n_cpi, float* rd_map,
                       rdp_det* dets, int32_t* n_dets, uint32_t* status);

/* Same with HOST buffers (pinned recommended): chunked, host->device and device->host
 * copies overlapped with the kernels on internal streams.  Synchronous.  n_cpi <= max_cpis.
 * The detection list comes back PACKED, not in rows: the records of CPI i start at record offset
 *     sum over k < i of min(n_dets[k], max_dets_per_cpi)
 * of `dets` (CPI 0 at offset 0; every record carries its cpi index), so `dets` needs room for that sum -- at most
 * n_cpi * max_dets_per_cpi records -- and only the records found cross PCIe.  rd_map / n_dets / status as above.  */
int rdp_process_host(rdp_handle* h, const void* cube, int32_t n_cpi, float* rd_map,
                     rdp_det* dets, int32_t* n_dets, uint32_t* status);

int rdp_sync(rdp_handle* h);

/* arith = FIXED16 only: rdp_process_device with the intermediate words exposed (all DEVICE pointers, nullable except
 * the detection outputs).  mag_map [n_cpi][channel][range][doppler] uint32 = the 17-bit words magnitude_calc hands
 * to os_cfar_2d; exp_range [n_cpi][channel][chirp] / exp_doppler [n_cpi][channel][range] int32 = the block exponents
 * of the range / Doppler transform of each frame -- the xfft core's m_axis_data_tuser, which radar_core leaves open
 * (rtl/src/radar_core.vhd:310,358): the RTL's CFAR therefore compares mantissas of differently scaled range lines,
 * and so does this mode; a caller that wants calibrated magnitudes multiplies by 2^(exp_range + exp_doppler).
 * With arith = FIXED16, rdp_process_device / _host / the streaming calls run this chain too; their rd_map then
 * receives the same uint32 words (cast the pointer), and every record's mag is the word converted to float.     */
int rdp_process_fixed_device(rdp_handle* h, const void* cube, int32_t n_cpi, uint32_t* mag_map, int32_t* exp_range,
                             int32_t* exp_doppler, rdp_det* dets, int32_t* n_dets, uint32_t* status);

/* Stage taps on DEVICE buffers -- each RTL stage is separately testable (rtl/src/tb_*.vhd):
 *   range stage   = window_multiplier + range FFT + corner_turner  (radar_core.vhd:267-327)
 *                   cube -> x_rc [n_cpi][channel][range][chirp] complex64
 *   doppler stage = doppler_notch + window_multiplier + Doppler FFT + magnitude_calc (:329-374)
 *                   x_rc -> rd_map [n_cpi][channel][range][doppler] float32 ([n_cpi][1] / [n_cpi][n_beams] maps per
 *                   CPI with combine = NCI / BEAMS: size the buffer for them)
 *   cfar stage    = os_cfar_2d + non-zero filter + coordinates (:376-418)
 *                   rd_map -> detections; cfar_out (nullable) is the RTL's dense output
 *                   word per cell: CUT magnitude or 0.                                    */
int rdp_range_stage(rdp_handle* h, const void* cube, int32_t n_cpi, void* x_rc);
int rdp_doppler_stage(rdp_handle* h, const void* x_rc, int32_t n_cpi, float* rd_map);
int rdp_cfar_stage(rdp_handle* h, const float* rd_map, int32_t n_cpi, float* cfar_out,
                   rdp_det* dets, int32_t* n_dets, uint32_t* status);

/* Fixed-point CFAR stage: the map holds UNSIGNED INTEGER magnitudes (the 17-bit words of
 * magnitude_calc; HARD PRECONDITION: every word < 2^24 -- the pre-filter converts words to float and is only
 * conservative while that is exact; a CPI that breaks it gets RDP_STATUS_MAP_RANGE and may lose detections) and every decision is the bit-exact integer
 * arithmetic of rtl/src/os_cfar_2d.vhd:152-217 -- integer sum, mean = floor(sum / N_REF)
 * truncated to 17 bits, lo = mean >> 1, hi = (mean + lo) mod 2^17, exact product x_k * scale.
 * cfar_out (nullable) is the RTL's output word per cell: the CUT magnitude or 0.  Records carry
 * the magnitude converted to float.  Needs a tiled kernel instance for the window shape
 * (the radar_core, tb_os_cfar_2d, legacy and "wide" shapes have one).                        */
int rdp_cfar_stage_u32(rdp_handle* h, const uint32_t* mag_map, int32_t n_cpi, uint32_t* cfar_out,
                       rdp_det* dets, int32_t* n_dets, uint32_t* status);

/* 16-bit fixed-point stage taps on DEVICE buffers: bit-exact restatements of the hand-written RTL
 * stages (the FFT IP's arithmetic is not defined in the reference tree, so K1/K2 stay float32).
 * A sample is the 32-bit stream word {Q[31:16], I[15:0]} = int16 I then int16 Q in memory.
 *   rdp_fixed_window      window_multiplier (rtl/src/window_multiplier.vhd:34-49,97-158) along rows of n
 *                         words: y = sat16((x * coef + 2^14) >> 14); *saturated (host, nullable) = the
 *                         sticky saturation_flag
 *   rdp_fixed_mti         doppler_notch (rtl/src/doppler_notch.vhd:70-104) along lines of n words,
 *                         mode 0 = bypass, 2, 3
 *   rdp_fixed_magnitude   magnitude_calc (rtl/src/magnitude_calc.vhd:55-81): word -> 17-bit unsigned
 *   rdp_fixed_corner_turn corner_turner (rtl/src/corner_turner.vhd:79-80): [map][chirp][sample] ->
 *                         [map][range][chirp] of 32-bit words                                       */
int rdp_fixed_window(rdp_handle* h, const void* in, int64_t n_rows, int32_t n, void* out, uint32_t* saturated);
int rdp_fixed_mti(rdp_handle* h, const void* in, int64_t n_lines, int32_t n, int32_t mode, void* out);
int rdp_fixed_magnitude(rdp_handle* h, const void* in, int64_t n_words, uint32_t* out);
int rdp_fixed_corner_turn(rdp_handle* h, const void* in, int32_t n_maps, int32_t n_chirps, int32_t n_samples,
                          void* out);
/* The transform FIXED16 uses in place of the two Xilinx xfft v9.1 cores (16-bit data and twiddles, block floating
 * point, convergent rounding: vivado_proj/.../ip/xfft_0_1/xfft_0.xci:12-27,59; their RTL is not in the reference
 * tree, so this is SELF-DEFINED and stated bit for bit in oracle/rtl_fixed.py:bfp_fft): log2(n) radix-2 DIF stages,
 * exact integer butterflies with Q15 twiddles and convergent >> 15, one convergent shift per stage for the whole
 * frame so that it fits 16 bits again; exps[frame] (DEVICE, nullable) = the sum of the shifts, DFT(x) ~= X * 2^exp.
 * n_frames frames of n words (n = 2..4096, power of two), natural order in and out.                               */
int rdp_fixed_bfp_fft(rdp_handle* h, const void* in, int64_t n_frames, int32_t n, void* out, int32_t* exps);

/* Streaming ingest (BASELINE config 5: pinned-host, multi-buffered chirp ingest).  n_slots
 * submissions may be in flight; each carries up to slot_cpis CPIs (<= the handle's chunk size).
 * rdp_stream_submit() is asynchronous: the host->device copy of submission i+1 overlaps the
 * kernels of submission i (the reference streams one word per clock with no host at all:
 * radar_core.vhd:25-29).  rdp_stream_wait() blocks until the slot's results are in page-locked host
 * memory owned by the handle and returns pointers to them, valid until the slot is submitted again:
 * dets = *total records of all the submission's CPIs, canonical order, cpi = index within the
 * submission; n_dets / status per CPI.  `cube` must stay valid until the wait returns; use pinned
 * memory (rdp_host_alloc) or the copy will not overlap.                                         */
int rdp_stream_open(rdp_handle* h, int32_t n_slots, int32_t slot_cpis);
int rdp_stream_submit(rdp_handle* h, int32_t slot, const void* cube, int32_t n_cpi);
int rdp_stream_wait(rdp_handle* h, int32_t slot, const rdp_det** dets, const int32_t** n_dets,
                    const uint32_t** status, int32_t* total);
int rdp_stream_close(rdp_handle* h);

/* Track-while-scan processor on the DEVICE detection list (the consumer of the path: tws_tracker in
 * rtl/src/radar_core.vhd:424-439), bit-exact with rtl/src/tws_tracker.vhd including its as-built
 * signal-semantics behaviour.  One tracker instance per channel; one scan per CPI (the RTL's top level
 * ends a scan at every range line's tlast, radar_core.vhd:434 -- a wiring defect that is not reproduced).
 *   rdp_tracker_reset  (re)initialise every instance (aresetn); cfg NULL = the radar_core generics
 *   rdp_track_device   dets / n_dets as produced by rdp_process_device for n_cpi CPIs (device pointers);
 *                      tracks [n_cpi][channel][32] records, n_tracks [n_cpi][channel] reported per scan,
 *                      active [n_cpi][channel] = active_tracks after the scan (all device pointers).
 *                      Track files persist between calls.                                           */
int rdp_tracker_reset(rdp_handle* h, const rdp_tracker_config* cfg);
int rdp_track_device(rdp_handle* h, const rdp_det* dets, const int32_t* n_dets, int32_t n_cpi, rdp_track* tracks,
                     int32_t* n_tracks, int32_t* active);

/* Sharded large CPI (BASELINE config 4: one CPI spread over the GPUs of a box).  Each rank holds n_chirps
 * of the CPI's n_chirps * world chirps and owns n_samples / world of its range lines:
 *   rdp_range_stage_scatter  window + range FFT of the local chirps, each range line stored DIRECTLY into the
 *                            owning rank's intermediate (peer stores over NVLink: the all-to-all corner turn is
 *                            fused into K1's transposed store, no separate collective or repack); the caller
 *                            barriers across ranks before the Doppler stage reads
 *   rdp_doppler_stage_rows   Doppler stage on an explicit number of range lines (handle: n_chirps = the full
 *                            chirp count); rows_per_map > 0 writes every group of rows_per_map lines with
 *                            pad_rows untouched lines above and below it (room for the CFAR halo)
 *   rdp_cfar_stage_rows      CFAR stage on maps of an explicit number of range lines (own lines + halo lines
 *                            copied from the neighbouring ranks)
 *   rdp_ipc_export / open / close   CUDA IPC handles (64 opaque bytes) to map a peer rank's buffer
 *   rdp_copy_d2d, rdp_copy_d2d_rows, rdp_device_memset   device-side copy (also from a mapped peer buffer;
 *                            _rows: n_rows pieces of row_bytes at the given pitches, e.g. the halo lines of every
 *                            map of a batch in one call) and fill                                              */
int rdp_range_stage_scatter(rdp_handle* h, const void* cube, int32_t n_cpi, const rdp_scatter* sc);
int rdp_doppler_stage_rows(rdp_handle* h, const void* x_rc, int64_t n_rows, int32_t rows_per_map, int32_t pad_rows,
                           float* rd_map);
int rdp_cfar_stage_rows(rdp_handle* h, const float* rd_map, int32_t n_cpi, int32_t n_range_rows, float* cfar_out,
                        rdp_det* dets, int32_t* n_dets, uint32_t* status);
int rdp_ipc_export(rdp_handle* h, void* device_ptr, unsigned char handle_bytes[64]);
int rdp_ipc_open(rdp_handle* h, const unsigned char handle_bytes[64], void** device_ptr);
int rdp_ipc_close(rdp_handle* h, void* device_ptr);
int rdp_copy_d2d(rdp_handle* h, void* dst_device, const void* src_device, size_t bytes);
int rdp_copy_d2d_rows(rdp_handle* h, void* dst_device, size_t dst_pitch, const void* src_device, size_t src_pitch,
                      size_t row_bytes, size_t n_rows);
int rdp_device_memset(rdp_handle* h, void* dst_device, int value, size_t bytes);

/* Device detection rows [n_cpi][max_dets_per_cpi] + per-CPI counts (as rdp_process_device / rdp_cfar_stage* leave
 * them) -> one packed host list in (cpi, channel, range, doppler) order: dets_host needs room for
 * sum(min(count, max_dets_per_cpi)) records (at most n_cpi * max_dets_per_cpi), n_cpi <= max_cpis.  The non-zero
 * filter's output FIFO read side (rtl/src/radar_core.vhd:396-418).                                                */
int rdp_gather_dets(rdp_handle* h, const rdp_det* dets_device, const int32_t* n_dets_device, int32_t n_cpi,
                    rdp_det* dets_host, int32_t* counts_host, int64_t* total);

/* Device / pinned-host memory owned by the caller (so a ctypes host needs no other CUDA binding). */
int rdp_device_alloc(rdp_handle* h, size_t bytes, void** ptr);
int rdp_device_free(rdp_handle* h, void* ptr);
int rdp_host_alloc(rdp_handle* h, size_t bytes, void** ptr);
int rdp_host_free(rdp_handle* h, void* ptr);
int rdp_copy_h2d(rdp_handle* h, void* dst_device, const void* src_host, size_t bytes);
int rdp_copy_d2h(rdp_handle* h, void* dst_host, const void* src_device, size_t bytes);

/* Measurement: CUDA-event timing on the handle's stream, per-stage accumulation, L2 flush. */
int rdp_timer_start(rdp_handle* h);
int rdp_timer_stop(rdp_handle* h, float* elapsed_ms);            /* synchronises            */
int rdp_set_stage_timing(rdp_handle* h, int32_t enable);         /* events around each kernel; also turns
                                                                    off the CFAR / FFT two-stream overlap
                                                                    so that each kernel is timed alone  */
/* ms[0..3] = range, doppler, cfar, compaction kernels; launches[0..3] likewise; resets the sums. */
int rdp_get_stage_timing(rdp_handle* h, float ms[4], int64_t launches[4]);
int rdp_flush_l2(rdp_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* RDP_H_ */
