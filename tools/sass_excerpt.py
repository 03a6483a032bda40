#!/usr/bin/env python
"""profiles/r2_sass_excerpt.txt: per-kernel instruction histogram and one example line of every TMA / mbarrier /
cluster / packed-bf16 instruction of the in-tree librdp.so.   usage: python tools/sass_excerpt.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fpga-fmcw-radar-processor_b200", "librdp.so")
FEATURE = r"UTMALDG|UTMASTG|UTMAPF|UTMACMDFLUSH|UBLKCP|HSET2\.BF16|SYNCS|UCGABAR|STG\.E\.128|LDS\.128|STS\.128|BAR\.SYNC|FENCE|ATOMS|REDUX|DADD|F2F\.F64|VOTE|ST\.E\.64|MAPA|PREFETCH|CCTL"
KERNELS = [
    ("K1 range stage, 512 samples, int16 input (the bench shape)", r"k1_range_kernelINS_7FftPlanILi512E.*EELi0EE"),
    ("K1 cluster variant, 4096 samples, complex64 (sharded mode)", r"k1_range_cluster_kernelINS_7FftPlanILi4096E.*EELi1EE"),
    ("K2 Doppler stage, 256 chirps, alpha-max-beta-min", r"k2_doppler_kernelINS_7FftPlanILi256E.*EELi0ELi0EE"),
    ("K2 with non-coherent channel integration", r"k2_doppler_kernelINS_7FftPlanILi256E.*EELi0ELi1EE"),
    ("K2 with coherent beams", r"k2_doppler_kernelINS_7FftPlanILi256E.*EELi0ELi2EE"),
    ("K3 OS-CFAR, radar_core window (5,6,1,2), float", r"k3_cfar_kernelILi5ELi6ELi1ELi2ELi64ELi1ELi1ELb0E"),
    ("K3 OS-CFAR, integer mode", r"k3_cfar_kernelILi5ELi6ELi1ELi2ELi64ELi1ELi1ELb1E"),
    ("K3b peak grouping", r"k3b_peak_group"), ("K4 row offsets", r"k4_row_offsets"), ("K4 emit", r"k4_emit"),
    ("K5 pack", r"k5_pack"), ("K6 tracker", r"k6_tracker"), ("block-floating-point FFT (arith FIXED16)", r"kf_bfp_fft"),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.split("\n")
    funcs, cur = [], None
    for line in sass:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = [m.group(1), collections.Counter(), []]
            funcs.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            cur[1][m.group(2).split(".")[0]] += 1
            if re.search(FEATURE, m.group(2)):
                cur[2].append((m.group(2), re.sub(r"\s+/\* 0x[0-9a-f]+ \*/$", "", line.strip())))
    print("# SASS excerpt of the in-tree librdp.so (cuobjdump -sass, sm_100a) -- made by tools/sass_excerpt.py")
    print("# per kernel: instruction count, op histogram (top 14), one example line of each TMA / mbarrier / cluster /")
    print("# packed-bf16 / wide shared-memory instruction with the number of times it occurs\n")
    for title, pat in KERNELS:
        for name, hist, feat in funcs:
            if re.search(pat, name):
                dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
                print(f"## {title}\n{dem[:240]}")
                print(f"{sum(hist.values())} SASS instructions; " + ", ".join("%s %d" % kv for kv in hist.most_common(14)))
                cnt = collections.Counter(op for op, _ in feat)
                seen = set()
                for op, line in feat:
                    if op not in seen:
                        seen.add(op)
                        print(f"    x{cnt[op]:<5d} {line}")
                print()
                break
    tot = collections.Counter()
    for _, _, feat in funcs:
        for op, _ in feat:
            if re.match(r"UTMA|HSET2|SYNCS|UCGABAR|UBLKCP|MAPA", op):
                tot[op] += 1
    print(f"## whole library ({len(funcs)} kernels): totals")
    print(", ".join("%s %d" % kv for kv in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
