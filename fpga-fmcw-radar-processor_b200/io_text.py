"""The reference's detection text format: one `range_bin doppler_bin magnitude` line per
detection, decimal integers (writer: rtl/src/tb_radar_core.vhd:149-160; reader:
model/ADR_visualize.py:109-122).  Files written here can be fed to the reference's plotters."""
from __future__ import annotations

import numpy as np


def write_detections(path, dets, cpi=None, channel=None) -> int:
    """Write `r d mag` lines for the detections of one (cpi, channel) (all if None).
    Magnitudes are rounded to integers as the RTL's 17-bit word would be.  Returns the line count."""
    sel = np.ones(len(dets), dtype=bool)
    if cpi is not None:
        sel &= dets["cpi"] == cpi
    if channel is not None:
        sel &= dets["channel"] == channel
    d = dets[sel]
    with open(path, "w") as f:
        for r, c, m in zip(d["range"], d["doppler"], d["mag"]):
            f.write(f"{int(r)} {int(c)} {int(round(float(m)))}\n")
    return int(sel.sum())


def read_detections(path) -> np.ndarray:
    """Parse `r d mag` lines the way model/ADR_visualize.py:109-122 does (3 integers, others skipped)."""
    rows = []
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) != 3:
                continue
            try:
                rows.append(tuple(int(p) for p in parts))
            except ValueError:
                continue
    return np.asarray(rows, dtype=np.int64).reshape(-1, 3)


def write_tracks(path, tracks, active, channel=0, full=False) -> int:
    """The reference's track logs for one channel.  `tracks` / `active` as returned by RadarCore.track().
    quick format (rtl/src/tb_radar_core.vhd:163-181): `TRK <id> R=<range q2> D=<doppler q2> Q=<quality>`;
    full format  (rtl/src/tb_tactical.vhd:350-363):   `TRK <id> R=.. D=.. VR=<range rate> Q=.. S=<status bits>`
    (`write(L, trk_status)` prints the 2-bit std_logic_vector as binary digits); both followed by
    `SCAN_END ACTIVE=<n>` per scan and both parsed by model/ADR_visualize.py:124-168.  Returns the number of lines."""
    n = 0
    with open(path, "w") as f:
        for cpi in range(active.shape[0]):
            for t in tracks[(tracks["cpi"] == cpi) & (tracks["channel"] == channel)]:
                if full:
                    f.write(f"TRK {int(t['id'])} R={int(t['range_q2'])} D={int(t['doppler_q2'])} VR={int(t['vel_r'])} "
                            f"Q={int(t['quality'])} S={int(t['status']) & 3:02b}\n")
                else:
                    f.write(f"TRK {int(t['id'])} R={int(t['range_q2'])} D={int(t['doppler_q2'])} Q={int(t['quality'])}\n")
                n += 1
            f.write(f"SCAN_END ACTIVE={int(active[cpi, channel])}\n")
            n += 1
    return n


def read_tracks(path):
    """Parse either track-log format the way model/ADR_visualize.py:124-168 (`load_tracks`) does.
    -> ({track id: {"range_q2": [...], "doppler_q2": [...], "quality": [...], "scan": [...]}}, [active per scan])."""
    tracks, scan_counts, scan = {}, [], 0
    with open(path) as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            if parts[0] == "TRK":
                t = tracks.setdefault(int(parts[1]), {"range_q2": [], "doppler_q2": [], "quality": [], "scan": []})
                q = 0
                for p in parts[4:]:
                    if p.startswith("Q="):
                        q = int(p.split("=")[1])
                        break
                t["range_q2"].append(int(parts[2].split("=")[1]))
                t["doppler_q2"].append(int(parts[3].split("=")[1]))
                t["quality"].append(q)
                t["scan"].append(scan)
            elif parts[0] == "SCAN_END":
                scan_counts.append(int(parts[1].split("=")[1]))
                scan += 1
    return tracks, scan_counts
