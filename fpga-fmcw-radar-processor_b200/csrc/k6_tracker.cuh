// k6_tracker.cuh -- track-while-scan processor consuming the detection list on the device.
//
// Bit-exact restatement of rtl/src/tws_tracker.vhd (the consumer of the hot path's detection stream,
// rtl/src/radar_core.vhd:424-439), including its as-built behaviour: association compares with the
// PREVIOUS track's best distance and keeps the last hit (:157-178, VHDL signal semantics), Q2 positions
// wrap in their 12-bit / 9-bit signed signals (:52-53,188-189,242-245), det_count is a 6-bit counter
// (:75,131-137), 4-bit hit / miss / quality counters.  One warp per tracker instance (= receive
// channel); scans (= CPIs) are processed in order, the track file lives in registers (lane = track)
// and is kept in global memory between launches.  Sequential by nature and tiny: not a throughput kernel.
#pragma once
#include "rdp_common.cuh"
#include "../../include/rdp.h"

namespace rdp {

struct TrackerArgs {
    int max_tracks, init_hits, coast_max, gate_r, gate_d, alpha, beta;
    int n_channels, max_dets;
};

struct TrackState {            // one tracker instance (channel), persistent between launches
    int active[32], status[32], range_pos[32], dopp_pos[32], range_vel[32], dopp_vel[32];
    int hit[32], miss[32], quality[32], age[32], last_mag[32];
    int best_det_idx, best_distance;
};

__device__ __forceinline__ int wrap_s(int v, int bits) {
    v &= (1 << bits) - 1;
    return (v >> (bits - 1)) ? v - (1 << bits) : v;
}

__global__ void __launch_bounds__(32)
k6_tracker_kernel(const rdp_det* __restrict__ dets, const int* __restrict__ counts, int n_cpi, TrackerArgs a,
                  TrackState* __restrict__ state, rdp_track* __restrict__ trk_out, int* __restrict__ n_trk_out,
                  int* __restrict__ active_out, int cpi_label_base) {
    __shared__ int b_r[64], b_d[64], b_mag[64], b_valid[64], b_assoc[64];
    const int ch = blockIdx.x, lane = threadIdx.x;
    TrackState* st = state + ch;
    const bool is_trk = lane < a.max_tracks;
    int active = is_trk ? st->active[lane] : 0, status = st->status[lane], range_pos = st->range_pos[lane];
    int dopp_pos = st->dopp_pos[lane], range_vel = st->range_vel[lane], dopp_vel = st->dopp_vel[lane];
    int hit = st->hit[lane], miss = st->miss[lane], quality = st->quality[lane], age = st->age[lane];
    int last_mag = st->last_mag[lane];
    int best_det_idx = st->best_det_idx, best_distance = st->best_distance;

    for (int cpi = 0; cpi < n_cpi; ++cpi) {
        // ---- ST_COLLECT (:130-141): this channel's run of the canonical (channel, range, doppler) list --------
        const rdp_det* row = dets + (long long)cpi * a.max_dets;
        const int n = counts[cpi] < a.max_dets ? counts[cpi] : a.max_dets;
        int lo = 0, hi = n;                               // first record with channel >= ch
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid].channel < ch) lo = mid + 1; else hi = mid; }
        const int first = lo;
        hi = n;                                           // first record with channel > ch
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid].channel <= ch) lo = mid + 1; else hi = mid; }
        const int len = lo - first;
        const int det_count = len & 63;                   // 6-bit counter: the 64th detection wraps it (:75,137)
        for (int e = lane; e < 64; e += 32) {
            const bool v = e < len;                       // entry e ends up holding the LAST detection with index = e mod 64
            const int src = v ? first + e + 64 * ((len - 1 - e) / 64) : first;
            b_valid[e] = v ? 1 : 0;
            b_assoc[e] = 0;
            if (v) {
                const rdp_det d = row[src];
                int mag = (int)(d.mag + 0.5f);
                b_r[e] = d.range & 0x3ff;                 // det_range is 10 bits, det_doppler 7, det_magnitude 17 (:26-28)
                b_d[e] = d.doppler & 0x7f;
                b_mag[e] = mag < 0 ? 0 : (mag > 0x1ffff ? 0x1ffff : mag);
            } else {
                b_r[e] = 0; b_d[e] = 0; b_mag[e] = 0;
            }
        }
        __syncwarp();
        // ---- ST_PREDICT (:143-157) -------------------------------------------------------------------------------
        if (active) {
            range_pos = wrap_s(range_pos + range_vel, 12);
            dopp_pos = wrap_s(dopp_pos + dopp_vel, 9);
            age = (age + 1) & 0xff;
        }
        // ---- ST_ASSOCIATE / ST_UPDATE, track by track (:159-236) ----------------------------------------------------
        for (int ti = 0; ti < a.max_tracks; ++ti) {
            if (!__shfl_sync(0xffffffffu, active, ti)) continue;
            const int rp = __shfl_sync(0xffffffffu, range_pos, ti), dp = __shfl_sync(0xffffffffu, dopp_pos, ti);
            const int old_best = best_distance;          // the OLD signal value: the previous track's result
            int dist[2];
            bool ok[2];
            for (int k = 0; k < 2; ++k) {
                const int e = lane + 32 * k;
                int dr = rp - ((b_r[e] << 2) & 0xfff), dd = dp - ((b_d[e] << 2) & 0x1ff);
                dr = dr < 0 ? -dr : dr;
                dd = dd < 0 ? -dd : dd;
                dist[k] = (dr + dd) & 0xffff;
                ok[k] = b_valid[e] && !b_assoc[e] && dr < a.gate_r * 4 && dd < a.gate_d * 4 && dist[k] < old_best;
            }
            const unsigned v_hi = __ballot_sync(0xffffffffu, ok[1]), v_lo = __ballot_sync(0xffffffffu, ok[0]);
            int win = -1, win_dist = 0xffff;             // every hit overwrites: the LAST qualifying entry wins
            if (v_hi) { const int l = 31 - __clz((int)v_hi); win = 32 + l; win_dist = __shfl_sync(0xffffffffu, dist[1], l); }
            else if (v_lo) { const int l = 31 - __clz((int)v_lo); win = l; win_dist = __shfl_sync(0xffffffffu, dist[0], l); }
            best_distance = win >= 0 ? win_dist : 0xffff;
            best_det_idx = win >= 0 ? win : 63;
            const bool assoc = best_det_idx < 64 && best_distance < 0xffff;
            if (lane == ti) {
                if (assoc) {
                    const int meas_r = wrap_s(b_r[best_det_idx] << 2, 12), meas_d = wrap_s(b_d[best_det_idx] << 2, 9);
                    const int innov_r = wrap_s(meas_r - range_pos, 12), innov_d = wrap_s(meas_d - dopp_pos, 9);
                    range_pos = wrap_s(range_pos + wrap_s((innov_r * a.alpha) >> 8, 12), 12);
                    dopp_pos = wrap_s(dopp_pos + wrap_s((innov_d * a.alpha) >> 8, 9), 9);
                    range_vel = wrap_s(range_vel + wrap_s((innov_r * a.beta) >> 8, 10), 10);
                    dopp_vel = wrap_s(dopp_vel + wrap_s((innov_d * a.beta) >> 8, 8), 8);
                    const int old_hit = hit;
                    hit = (hit + 1) & 0xf;
                    miss = 0;
                    last_mag = b_mag[best_det_idx];
                    if (status == 1 && old_hit >= a.init_hits) status = 2;       // TENTATIVE -> FIRM
                    else if (status == 3) status = 2;                            // COAST -> FIRM
                    if (quality < 15) quality += 1;
                } else {
                    const int old_miss = miss;
                    miss = (miss + 1) & 0xf;
                    if (status == 2) status = 3;                                 // FIRM -> COAST
                    if (old_miss >= a.coast_max) { active = 0; status = 0; }
                    if (quality > 0) quality -= 1;
                }
            }
            if (assoc && lane == 0) b_assoc[best_det_idx] = 1;
            __syncwarp();
        }
        // ---- ST_INITIATE (:238-268): entries 0 .. det_count-1 (entry 0 even when det_count = 0) -------------------
        for (int idx = 0;; idx = (idx + 1) & 63) {
            if (b_valid[idx] && !b_assoc[idx]) {
                const unsigned free_mask = __ballot_sync(0xffffffffu, is_trk && !active);
                if (free_mask && lane == __ffs((int)free_mask) - 1) {
                    active = 1; status = 1;
                    range_pos = wrap_s(b_r[idx] << 2, 12);
                    dopp_pos = wrap_s(b_d[idx] << 2, 9);
                    range_vel = 0; dopp_vel = 0; hit = 1; miss = 0; quality = 1; age = 0;
                    last_mag = b_mag[idx];
                }
            }
            if (det_count == 0 || idx == ((det_count - 1) & 63)) break;
        }
        // ---- ST_MAINTAIN (:270-276), ST_OUTPUT (:278-300) ----------------------------------------------------------
        const unsigned act = __ballot_sync(0xffffffffu, active != 0);
        const bool report = active && (status == 2 || status == 3);
        const unsigned rep = __ballot_sync(0xffffffffu, report);
        const long long scan = (long long)cpi * a.n_channels + ch;
        if (report) {
            rdp_track t;
            t.cpi = cpi_label_base + cpi; t.channel = ch; t.id = lane;
            t.range_q2 = range_pos; t.doppler_q2 = dopp_pos; t.vel_r = range_vel; t.vel_d = dopp_vel;
            t.quality = quality; t.status = status;
            trk_out[scan * 32 + __popc(rep & ((1u << lane) - 1u))] = t;
        }
        if (lane == 0) { n_trk_out[scan] = __popc(rep); active_out[scan] = __popc(act); }
        __syncwarp();
    }
    st->active[lane] = active; st->status[lane] = status; st->range_pos[lane] = range_pos; st->dopp_pos[lane] = dopp_pos;
    st->range_vel[lane] = range_vel; st->dopp_vel[lane] = dopp_vel; st->hit[lane] = hit; st->miss[lane] = miss;
    st->quality[lane] = quality; st->age[lane] = age; st->last_mag[lane] = last_mag;
    if (lane == 0) { st->best_det_idx = best_det_idx; st->best_distance = best_distance; }
}

}  // namespace rdp
