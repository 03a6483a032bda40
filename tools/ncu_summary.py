#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` / `--page source --csv` exports (tools/gpu/session2b.sh) for profiles/.
usage: python tools/ncu_summary.py raw  <raw.csv>
       python tools/ncu_summary.py src  <source.csv> <file-substring> [top-n]      (instructions executed by source line)"""
import csv
import sys

RAW_KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    for k in RAW_KEYS:
        if k in idx:
            print(f"{k:80s} {vals[idx[k]]} {units[idx[k]]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(vals[i]) >= 0.05:
                    print(f"{h:80s} {vals[i]}")
            except ValueError:
                pass


def src(path, needle, top=40):
    rows = list(csv.reader(open(path)))
    # find the header row of the source table
    for r0, r in enumerate(rows):
        if "Source" in r and any("Instructions Executed" in c for c in r):
            break
    hdr = rows[r0]
    ci = {h: i for i, h in enumerate(hdr)}
    col = ci.get("Instructions Executed")
    out = []
    for r in rows[r0 + 1:]:
        if len(r) <= col:
            continue
        try:
            out.append((float(r[col]), r))
        except ValueError:
            pass
    total = sum(v for v, _ in out)
    print("total warp instructions", total)
    return hdr, ci, out, total


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        hdr, ci, out, total = src(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
        print(hdr)
        for v, r in sorted(out, key=lambda x: -x[0])[: int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
            print("%5.2f%%" % (100 * v / total), r[:3])
