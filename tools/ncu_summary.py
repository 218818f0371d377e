#!/usr/bin/env python
"""Headline counters of the `ncu --set full` raw pages kept under profiles/ -> profiles/r01_ncu_summary.json.
Usage: python tools/ncu_summary.py   (reads profiles/r01_*_ncu_full_raw.csv)"""
import csv
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_global_red.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_*_ncu_full_raw.csv"))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        name = os.path.basename(path)[len("r01_"):-len("_ncu_full_raw.csv")]
        out[name] = {m: ("%s %s" % d[m]).strip() for m in METRICS if m in d}
    with open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v.get("gpu__time_duration.sum") for k, v in out.items()}))


if __name__ == "__main__":
    main()
