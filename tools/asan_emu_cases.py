import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import fpga_fmcw_radar_processor_b200 as rdp
import parity_cases as pc
EMU = os.environ.get('RDP_ASAN_EMU', '/tmp/rdp_asan/librdp_emu.so')
def make_core(**kw):
    return rdp.RadarCore(library=EMU, _test_allow_emulator=True, **kw)
cases = [
    ("range", lambda: pc.case_range_stage(make_core, 64, 32)),
    ("doppler", lambda: pc.case_doppler_stage(make_core, 16, 64, 2, "amb")),
    ("cfar", lambda: pc.case_cfar_stage(make_core, 64, 128)),
    ("cfar raster partial", lambda: pc.case_cfar_stage(make_core, 16, 32, cfar_edges="raster")),
    ("cfar wide", lambda: pc.case_cfar_stage(make_core, 32, 64, cfar_ref_r=8, cfar_ref_d=8, cfar_guard_r=4, cfar_guard_d=2)),
    ("cfar generic", lambda: pc.case_cfar_stage(make_core, 32, 32, cfar_ref_r=1, cfar_ref_d=2, cfar_guard_r=0, cfar_guard_d=1)),
    ("signed", lambda: pc.case_cfar_signed_map(make_core)),
    ("crowded", lambda: pc.case_cfar_crowded_tile(make_core)),
    ("fixed cfar", lambda: pc.case_cfar_fixed_point(make_core)),
    ("e2e", lambda: pc.case_end_to_end(make_core, 2, 2, 32, 64)),
    ("bfp", lambda: pc.case_bfp_fft(make_core, sizes=(2, 8, 64))),
    ("fixed e2e", lambda: pc.case_fixed_end_to_end(make_core, mti_mode=2)),
    ("combine", lambda: pc.case_combine_and_group(make_core)),
    ("adc", lambda: pc.case_adc_full_scale_flag(make_core)),
    ("overflow", lambda: pc.case_overflow_flag(make_core)),
    ("streaming", lambda: pc.case_streaming(make_core)),
    ("tracker", lambda: pc.case_tracker(make_core)),
    ("fixed stages", lambda: pc.case_fixed_stages(make_core)),
    ("range flag", lambda: pc.case_cfar_fixed_range_flag(make_core)),
]
for name, fn in cases:
    fn(); print("ok", name, flush=True)
