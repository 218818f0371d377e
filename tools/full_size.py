#!/usr/bin/env python
"""Full-size single-GPU runs of BASELINE.json configs 2-5 (10 M / 60 M / 200 M / 1 B rows), columns resident in HBM.

The keyspaces, queries and checks live in tools/workloads.py (device-generated columns handed to the library through
n1gpu_table_set_column_device; every result checked against torch reductions over the regenerated chunks - exact for
counts, integer sums, min / max, group and DISTINCT sets, 1e-12 against exact integer-scaled totals for float sums).

Prints one JSON line per config.  Usage: python tools/full_size.py [config2 config3 config4 config5] ; FS_SCALE=0.01
shrinks every row count (smoke run)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import query_b200 as q  # noqa: E402
import workloads as wl  # noqa: E402

PEAK = 6547.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(query, reps=3):
    """median device time of the scan and median wall time of launch + collect + fetch of the groups"""
    query.set_timing(True)
    res = query.execute()  # warm-up (sizes the hash tables; may grow and rerun)
    ns, wall = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        query.launch()
        res = query.collect()
        n_groups = res.num_groups  # forces the finalisation
        wall.append(time.perf_counter() - t0)
        ns.append(query.last_scan_ns)
    return res, sorted(ns)[len(ns) // 2], sorted(wall)[len(wall) // 2], n_groups


def run(name, scale=None):
    scale = float(os.environ.get("FS_SCALE", "1")) if scale is None else scale
    w = wl.CONFIGS[name](scale=scale)
    t = w.sealed_table()
    qq = w.query(t)
    res, scan_ns, wall, ng = timed(qq, reps=3)
    ref = w.reference()
    check = w.check(res, ref)
    info = qq.info
    bpr = info["scan_bytes_per_row"]
    gbs = bpr * w.rows / scan_ns
    out = {"config": name, "rows": w.rows, "mode": info["mode"], "registers": info["registers"], "grid": info["grid"], "block": info["block"],
           "scan_bytes_per_row": bpr, "survey_bytes_per_row": w.survey_bytes_per_row, "scan_us": scan_ns / 1e3, "rows_per_s": w.rows / (scan_ns * 1e-9),
           "gb_per_s": gbs, "hbm_peak_gb_per_s": PEAK, "roofline_frac": gbs / PEAK, "scan_plus_finalize_wall_ms": wall * 1e3, "groups": ng, "check": check}
    print(json.dumps(out))
    sys.stdout.flush()
    return out


if __name__ == "__main__":
    q.init(0)
    for name in sys.argv[1:] or ["config3", "config4", "config5"]:
        run(name)
        torch.cuda.empty_cache()
