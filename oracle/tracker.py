"""Bit-exact restatement of the track-while-scan processor.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/rtl/src/tws_tracker.vhd state by state with VHDL *signal* semantics
(a signal assigned in a clock cycle is read with its OLD value in that cycle).  As-built behaviour
that a reader of the comments would not expect, all reproduced here:

  * ST_ASSOCIATE (:157-178) compares each candidate's distance with the OLD `best_distance` signal --
    the previous track's result -- not with the 0xFFFF assigned at the top of the same cycle, and every
    hit overwrites the scheduled value: the winner is the LAST in-gate detection whose distance is
    below the previous track's best distance.
  * positions are Q2 in 12-bit / 9-bit SIGNED signals (:52-53): a range bin >= 512 or a Doppler bin
    >= 64 wraps negative when a track is initiated or updated (:188-189,242-245), while the
    association compares against the UNSIGNED detection position (:165-168), so such tracks never
    associate again.
  * `det_count` is a 6-bit counter (:75): the 64th detection wraps it to 0; ST_INITIATE walks only
    det_count entries (:254).
  * hit / miss / quality counters are 4-bit and compare with their old values (:200-203,217).
"""
from __future__ import annotations

FREE, TENTATIVE, FIRM, COAST = 0, 1, 2, 3
MAX_DETS = 64


def _wrap_s(v: int, bits: int) -> int:
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


class TwsTracker:
    def __init__(self, max_tracks=32, init_hits=2, coast_max=5, assoc_gate_r=10, assoc_gate_d=5,
                 alpha_gain=128, beta_gain=64):
        self.max_tracks, self.init_hits, self.coast_max = max_tracks, init_hits, coast_max
        self.gate_r, self.gate_d, self.alpha, self.beta = assoc_gate_r, assoc_gate_d, alpha_gain, beta_gain
        # reset values, tws_tracker.vhd:106-126 (signal initialisers :81-83 for best_*)
        self.trk = [dict(active=0, status=FREE, range_pos=0, dopp_pos=0, range_vel=0, dopp_vel=0, hit=0, miss=0,
                         quality=0, age=0, last_mag=0) for _ in range(max_tracks)]
        self.best_det_idx = 63
        self.best_distance = 0xFFFF
        self.num_active = 0

    def scan(self, detections):
        """One scan: `detections` = iterable of (range_bin, doppler_bin, magnitude) in arrival order, then
        det_last.  Returns (tracks, active) with tracks = list of dicts for every FIRM / COAST track
        (id, range_q2, doppler_q2, vel_r, vel_d, quality, status) in id order, as ST_OUTPUT emits them."""
        # ST_COLLECT (:130-141)
        buf = [dict(valid=0, r=0, d=0, mag=0, assoc=0) for _ in range(MAX_DETS)]
        det_count = 0
        for (r, d, mag) in detections:
            buf[det_count] = dict(valid=1, r=int(r) & 0x3FF, d=int(d) & 0x7F, mag=int(mag) & 0x1FFFF, assoc=0)
            det_count = (det_count + 1) & 63
        # ST_PREDICT (:143-157)
        for t in self.trk:
            if t["active"]:
                t["range_pos"] = _wrap_s(t["range_pos"] + t["range_vel"], 12)
                t["dopp_pos"] = _wrap_s(t["dopp_pos"] + t["dopp_vel"], 9)
                t["age"] = (t["age"] + 1) & 0xFF
        # ST_ASSOCIATE / ST_UPDATE per track (:159-236)
        for t in self.trk:
            if not t["active"]:
                continue
            old_best = self.best_distance
            new_dist, new_idx = 0xFFFF, 63
            for i, e in enumerate(buf):
                if e["valid"] and not e["assoc"]:
                    dist_r = abs(t["range_pos"] - ((e["r"] << 2) & 0xFFF))
                    dist_d = abs(t["dopp_pos"] - ((e["d"] << 2) & 0x1FF))
                    if dist_r < self.gate_r * 4 and dist_d < self.gate_d * 4:
                        distance = (dist_r + dist_d) & 0xFFFF
                        if distance < old_best:
                            new_dist, new_idx = distance, i
            self.best_distance, self.best_det_idx = new_dist, new_idx
            if self.best_det_idx < MAX_DETS and self.best_distance < 0xFFFF:
                e = buf[self.best_det_idx]
                e["assoc"] = 1
                meas_r = _wrap_s(e["r"] << 2, 12)
                meas_d = _wrap_s(e["d"] << 2, 9)
                innov_r = _wrap_s(meas_r - t["range_pos"], 12)
                innov_d = _wrap_s(meas_d - t["dopp_pos"], 9)
                t["range_pos"] = _wrap_s(t["range_pos"] + _wrap_s((innov_r * self.alpha) >> 8, 12), 12)
                t["dopp_pos"] = _wrap_s(t["dopp_pos"] + _wrap_s((innov_d * self.alpha) >> 8, 9), 9)
                t["range_vel"] = _wrap_s(t["range_vel"] + _wrap_s((innov_r * self.beta) >> 8, 10), 10)
                t["dopp_vel"] = _wrap_s(t["dopp_vel"] + _wrap_s((innov_d * self.beta) >> 8, 8), 8)
                old_hit = t["hit"]
                t["hit"] = (t["hit"] + 1) & 0xF
                t["miss"] = 0
                t["last_mag"] = e["mag"]
                if t["status"] == TENTATIVE and old_hit >= self.init_hits:
                    t["status"] = FIRM
                elif t["status"] == COAST:
                    t["status"] = FIRM
                if t["quality"] < 15:
                    t["quality"] += 1
            else:
                old_miss = t["miss"]
                t["miss"] = (t["miss"] + 1) & 0xF
                if t["status"] == FIRM:
                    t["status"] = COAST
                if old_miss >= self.coast_max:
                    t["active"], t["status"] = 0, FREE
                if t["quality"] > 0:
                    t["quality"] -= 1
        # ST_INITIATE (:238-268): entries 0 .. det_count-1 (entry 0 is looked at even when det_count = 0)
        idx = 0
        while True:
            e = buf[idx]
            if e["valid"] and not e["assoc"]:
                for t in self.trk:
                    if not t["active"]:
                        t.update(active=1, status=TENTATIVE, range_pos=_wrap_s(e["r"] << 2, 12),
                                 dopp_pos=_wrap_s(e["d"] << 2, 9), range_vel=0, dopp_vel=0, hit=1, miss=0,
                                 quality=1, age=0, last_mag=e["mag"])
                        break
            if det_count == 0 or idx == ((det_count - 1) & 63):
                break
            idx = (idx + 1) & 63
        # ST_MAINTAIN (:270-276), ST_OUTPUT (:278-300)
        self.num_active = sum(t["active"] for t in self.trk)
        out = [dict(id=i, range_q2=t["range_pos"], doppler_q2=t["dopp_pos"], vel_r=t["range_vel"], vel_d=t["dopp_vel"],
                    quality=t["quality"], status=t["status"])
               for i, t in enumerate(self.trk) if t["active"] and t["status"] in (FIRM, COAST)]
        return out, self.num_active


def track_lines(tracks, active):
    """The testbench's log lines (rtl/src/tb_radar_core.vhd:163-181): `TRK <id> R=<q2> D=<q2> Q=<quality>` per
    reported track, then `SCAN_END ACTIVE=<n>`; parsed by model/ADR_visualize.py:139-166."""
    lines = [f"TRK {t['id']} R={t['range_q2']} D={t['doppler_q2']} Q={t['quality']}" for t in tracks]
    lines.append(f"SCAN_END ACTIVE={active}")
    return lines
