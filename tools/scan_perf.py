#!/usr/bin/env python
"""Kernel-level timing sweep (not the bench): nq_scan duration vs rows for the BASELINE config shapes, with
pre-shredded resident columns.  Prints one line per (shape, rows): kernel us (CUDA events around the launch,
queue kept busy), GB/s of column bytes, rows/s.   Usage: python tools/scan_perf.py [shape ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import query_b200 as q  # noqa: E402


def table_config2(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1_000_000, n, dtype=np.int64)
    a[0], a[1] = 0, 999_999
    f = rng.integers(1, 1_000_000, n).astype(np.float64) / 1e6
    t = q.Table(["n", "f"])
    t.set_column("n", a)
    t.set_column("f", f, tags=np.full(n, 5, dtype=np.uint8))
    return t.seal()


def table_config3(n, seed):
    rng = np.random.default_rng(seed)
    t = q.Table(["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"])
    dates = ["%04d-%02d-%02d" % (y, m, d) for y in range(1992, 1999) for m in range(1, 13) for d in range(1, 29)]
    t.set_column("l_shipdate", rng.integers(0, len(dates), n).astype(np.uint32), dictionary=dates)
    t.set_column("l_returnflag", rng.integers(0, 3, n).astype(np.uint32), dictionary=["A", "N", "R"])
    t.set_column("l_linestatus", rng.integers(0, 2, n).astype(np.uint32), dictionary=["F", "O"])
    t.set_column("l_quantity", rng.integers(1, 51, n, dtype=np.int64))
    price = (rng.integers(90000, 10500000, n).astype(np.float64) + 0.5) / 100.0
    t.set_column("l_extendedprice", price, tags=np.full(n, 5, dtype=np.uint8))
    disc = rng.integers(1, 11, n).astype(np.float64) / 100.0
    t.set_column("l_discount", disc, tags=np.full(n, 5, dtype=np.uint8))
    tax = rng.integers(1, 9, n).astype(np.float64) / 100.0
    t.set_column("l_tax", tax, tags=np.full(n, 5, dtype=np.uint8))
    return t.seal()


def table_config4(n, seed):
    rng = np.random.default_rng(seed)
    t = q.Table(["g", "x"])
    t.set_column("g", rng.integers(0, 1_000_000, n, dtype=np.int64))
    t.set_column("x", rng.integers(0, 1000, n, dtype=np.int64))
    return t.seal()


def table_config5(n, seed):
    rng = np.random.default_rng(seed)
    vocab = 100_000
    words = sorted("w%06d-%x" % (i, (i * 2654435761) & 0xffffff) for i in range(vocab))
    # Zipf(s = 1.1) over the vocabulary by inverse CDF, popularity rank -> dictionary code through a fixed permutation
    # (the distribution tools/full_size.py and tools/sweep_config5.py use for BASELINE config 5)
    w = np.arange(1, vocab + 1, dtype=np.float64) ** -1.1
    cdf = np.cumsum(w / w.sum())
    perm = np.random.default_rng(4).permutation(vocab).astype(np.uint32)
    rank = perm[np.minimum(np.searchsorted(cdf, rng.random(n)), vocab - 1)]
    ktag = np.full(n, 6, dtype=np.uint8)
    r = rng.integers(0, 10, n)
    ktag[r == 0] = 0
    ktag[r == 1] = 1
    v = rng.integers(-1000, 1_000_000, n, dtype=np.int64)
    vtag = np.full(n, 4, dtype=np.uint8)
    r = rng.integers(0, 10, n)
    vtag[r == 0] = 0
    vtag[r == 1] = 1
    t = q.Table(["k", "v"])
    t.set_column("k", rank, tags=ktag, dictionary=words)
    t.set_column("v", v, tags=vtag)
    return t.seal()


SHAPES = {
    "config2": (table_config2, "d", "((`d`.`n`) between 250000 and 749999)", [],
                ["count(*)", "count((`d`.`n`))", "sum((`d`.`n`))", "avg((`d`.`n`))", "min((`d`.`n`))", "max((`d`.`n`))", "sum((`d`.`f`))"]),
    "config2_sel1": (table_config2, "d", "((`d`.`n`) between 0 and 9999)", [],
                     ["count(*)", "count((`d`.`n`))", "sum((`d`.`n`))", "avg((`d`.`n`))", "min((`d`.`n`))", "max((`d`.`n`))", "sum((`d`.`f`))"]),
    "config3": (table_config3, "l", "((`l`.`l_shipdate`) <= \"1998-09-02\")", ["(`l`.`l_returnflag`)", "(`l`.`l_linestatus`)"],
                ["sum((`l`.`l_quantity`))", "sum((`l`.`l_extendedprice`))", "sum(((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))))",
                 "sum((((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))) * (1 + (`l`.`l_tax`))))", "avg((`l`.`l_quantity`))",
                 "avg((`l`.`l_extendedprice`))", "avg((`l`.`l_discount`))", "count(*)"]),
    "config4": (table_config4, "d", None, ["(`d`.`g`)"], ["count(distinct (`d`.`x`))", "sum(distinct (`d`.`x`))", "count(*)"]),
    "config4_nodistinct": (table_config4, "d", None, ["(`d`.`g`)"], ["sum((`d`.`x`))", "count(*)"]),
    "config5": (table_config5, "d", "((`d`.`v`) is not missing)", ["(`d`.`k`)"],
                ["count(*)", "count((`d`.`v`))", "sum((`d`.`v`))", "min((`d`.`v`))", "max((`d`.`v`))"]),
}


def main():
    q.init(0)
    shapes = sys.argv[1:] or ["config2"]
    sizes = [int(x) for x in os.environ.get("ROWS", "10000000,40000000").split(",")]
    for name in shapes:
        mk, alias, where, keys, aggs = SHAPES[name]
        for n in sizes:
            tabs = [mk(n, s + 1) for s in range(2)]
            qs = [q.Query(tabs[i % 2], alias, where, keys, aggs) for i in range(4)]
            for qq in qs:
                qq.set_stream(0)
            for qq in qs:
                qq.execute()
            ns = []
            t0 = time.perf_counter()
            for rep in range(5):
                for qq in qs:
                    qq.launch()
                for qq in qs:
                    r = qq.collect()
                    ns.append(qq.last_scan_ns)
            wall = (time.perf_counter() - t0) / 20
            info = qs[0].info
            us = sorted(ns)[len(ns) // 2] / 1e3
            gbs = info["scan_bytes_per_row"] * n / (us * 1e3)
            print("%-20s rows=%-10d mode=%-20s regs=%-3d grid=%-5d blk=%-4d B/row=%-3d kernel=%9.1f us  %7.1f GB/s  %.3e rows/s  wall/step=%.1f us groups=%d" % (
                name, n, info["mode"], info["registers"], info["grid"], info["block"], info["scan_bytes_per_row"], us, gbs, n / (us * 1e-6), wall * 1e6, r.num_groups))
            sys.stdout.flush()
            del qs, tabs


if __name__ == "__main__":
    main()
