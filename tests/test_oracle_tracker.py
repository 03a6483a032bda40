"""Oracle tracker (bit-exact restatement of rtl/src/tws_tracker.vhd) against the scripted scenario and the
assertions of the reference's own testbench rtl/src/tb_tws_tracker.vhd:103-171.  CPU only."""
from oracle import tracker as tk


def tb_scan_detections(scan):
    """tb_tws_tracker.vhd:108-128: target 1 closes 5 bins / scan from range 200 (Doppler 40); target 2 sits at
    (600, 80); target 3 appears in scans 4..7; a false alarm at (900, 10) every third scan."""
    dets = []
    t1_r = 200 - (scan - 1) * 5
    if t1_r > 0:
        dets.append((t1_r, 40, 5000))
    dets.append((600, 80, 8000))
    if 4 <= scan <= 7:
        dets.append((400 + (scan - 4) * 3, 60, 3000))
    if scan % 3 == 0:
        dets.append((900, 10, 2000))
    return dets


def run_tb():
    trk = tk.TwsTracker(max_tracks=16, init_hits=2, coast_max=3, assoc_gate_r=10, assoc_gate_d=5)   # :11-18
    hist = []
    for scan in range(1, 13):
        out, active = trk.scan(tb_scan_detections(scan))
        hist.append((scan, active, out))
    return hist


def test_tb_assertions_hold():
    hist = {s: (a, o) for s, a, o in run_tb()}
    assert hist[2][0] >= 2          # :134-139
    assert hist[3][0] >= 2          # :141-146
    assert hist[6][0] >= 3          # :148-153


def test_target1_is_tracked_and_closing():
    """Target 1 (range < 512, Doppler < 64: inside the signed Q2 range) becomes FIRM and its track follows it."""
    hist = run_tb()
    firm = [[t for t in o if t["status"] == tk.FIRM and abs(t["range_q2"] - 4 * (200 - (s - 1) * 5)) <= 16] for s, a, o in hist]
    assert all(len(f) == 1 for f in firm[4:])                  # firm from scan 5 on (hit_count compares with its OLD value)
    r = [f[0]["range_q2"] for f in firm[4:]]
    assert all(b < a for a, b in zip(r, r[1:]))                # closing
    assert firm[-1][0]["vel_r"] < 0


def test_as_built_signed_wrap_never_associates():
    """Target 2 at range 600 / Doppler 80 initiates a track whose Q2 position wraps negative in the 12-bit / 9-bit
    signed signals (tws_tracker.vhd:242-245), so it never associates again: a fresh tentative track per scan."""
    trk = tk.TwsTracker(max_tracks=16, coast_max=3)
    for _ in range(4):
        out, active = trk.scan([(600, 80, 8000)])
    pos = [t["range_pos"] for t in trk.trk if t["active"]]
    assert all(p == 600 * 4 - 4096 for p in pos) and len(pos) >= 3
    assert out == []                                            # never FIRM


def test_log_lines_format():
    out, active = tk.TwsTracker().scan([(10, 5, 100)])
    assert tk.track_lines(out, active) == ["SCAN_END ACTIVE=1"]
    lines = tk.track_lines([dict(id=3, range_q2=400, doppler_q2=-12, vel_r=0, vel_d=0, quality=7, status=tk.FIRM)], 2)
    assert lines == ["TRK 3 R=400 D=-12 Q=7", "SCAN_END ACTIVE=2"]      # tb_radar_core.vhd:168-179


def test_det_count_wraps_at_64():
    trk = tk.TwsTracker()
    dets = [(i, i % 60, 100 + i) for i in range(70)]            # 70 detections: entries 0..5 are overwritten
    trk.scan(dets)
    # det_count wrapped to 6: ST_INITIATE looked at entries 0..5 only -> 6 tentative tracks
    assert sum(t["active"] for t in trk.trk) == 6
    assert sorted(t["range_pos"] // 4 for t in trk.trk if t["active"]) == [64, 65, 66, 67, 68, 69]
