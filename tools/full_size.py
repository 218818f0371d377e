#!/usr/bin/env python
"""Full-size single-GPU runs of BASELINE.json configs 3, 4 and 5 (60 M / 200 M / 1 B rows), columns resident in HBM.

The columns are generated ON the device with torch (seeded) and handed to the library through
n1gpu_table_set_column_device, so no host staging of 14 GB is needed.  The oracle cannot run these sizes; each result
is checked against torch reductions over the very same device tensors (exact for counts, integer sums, min/max,
group sets and DISTINCT sets; 1e-9 relative for float64 sums, whose summation order differs) - size-independent
properties, the bit-exact parity against the oracle is tests/test_gpu_parity.py at small sizes.

Prints one JSON line per config.  Usage: python tools/full_size.py [config3 config4 config5] ; FS_SCALE=0.01 shrinks
every row count (smoke run)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import query_b200 as q  # noqa: E402

DEV = torch.device("cuda:0")
SCALE = float(os.environ.get("FS_SCALE", "1"))
PEAK = 6547.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

C_MISSING, C_NULL, C_INT, C_FLOAT, C_STRING = 0, 1, 4, 5, 6


def timed(query, reps=3):
    query.set_timing(True)
    res = query.execute()  # warm-up (sizes the hash tables; may grow and rerun)
    ns = []
    wall = []
    for _ in range(reps):
        t0 = time.perf_counter()
        query.launch()
        res = query.collect()
        n_groups = res.num_groups  # forces the finalisation
        wall.append(time.perf_counter() - t0)
        ns.append(query.last_scan_ns)
    return res, sorted(ns)[len(ns) // 2], sorted(wall)[len(wall) // 2], n_groups


def line(name, n, query, res, scan_ns, wall_s, extra):
    info = query.info
    bpr = info["scan_bytes_per_row"]
    gbs = bpr * n / scan_ns
    out = {"config": name, "rows": n, "mode": info["mode"], "registers": info["registers"], "grid": info["grid"], "block": info["block"],
           "scan_bytes_per_row": bpr, "scan_us": scan_ns / 1e3, "rows_per_s": n / (scan_ns * 1e-9), "gb_per_s": gbs, "hbm_peak_gb_per_s": PEAK,
           "roofline_frac": gbs / PEAK, "scan_plus_finalize_wall_ms": wall_s * 1e3, "groups": res.num_groups}
    out.update(extra)
    print(json.dumps(out))
    sys.stdout.flush()


def ints(res, col):
    res._fetch()
    assert (res.agg_cls[:, col] == C_INT).all(), "aggregate %d is not an int everywhere" % col
    return res.agg_val[:, col]


# ---- config 3: TPC-H Q1 shape ----------------------------------------------------------------------------------------
def config3():
    n = int(60_000_000 * SCALE)
    g = torch.Generator(device=DEV).manual_seed(2)
    dates = ["%04d-%02d-%02d" % (y, m, d) for y in range(1992, 1999) for m in range(1, 13) for d in range(1, 29)]
    ship = torch.randint(0, len(dates), (n,), generator=g, device=DEV, dtype=torch.int32)
    rf = torch.randint(0, 3, (n,), generator=g, device=DEV, dtype=torch.int32)
    ls = torch.randint(0, 2, (n,), generator=g, device=DEV, dtype=torch.int32)
    qty = torch.randint(1, 51, (n,), generator=g, device=DEV, dtype=torch.int64)
    price = (torch.randint(90000, 10500000, (n,), generator=g, device=DEV, dtype=torch.int64).double() + 0.5) / 100.0
    disc = torch.randint(1, 11, (n,), generator=g, device=DEV, dtype=torch.int64).double() / 100.0
    tax = torch.randint(1, 9, (n,), generator=g, device=DEV, dtype=torch.int64).double() / 100.0
    ftag = torch.full((n,), C_FLOAT, dtype=torch.uint8, device=DEV)
    t = q.Table(["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"])
    t.set_column_device("l_shipdate", ship, dictionary=dates)
    t.set_column_device("l_returnflag", rf, dictionary=["A", "N", "R"])
    t.set_column_device("l_linestatus", ls, dictionary=["F", "O"])
    t.set_column_device("l_quantity", qty)
    t.set_column_device("l_extendedprice", price, tags=ftag)
    t.set_column_device("l_discount", disc, tags=ftag)
    t.set_column_device("l_tax", tax, tags=ftag)
    t.seal()
    aggs = ["sum((`l`.`l_quantity`))", "sum((`l`.`l_extendedprice`))", "sum(((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))))",
            "sum((((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))) * (1 + (`l`.`l_tax`))))", "avg((`l`.`l_quantity`))",
            "avg((`l`.`l_extendedprice`))", "avg((`l`.`l_discount`))", "count(*)"]
    qq = q.Query(t, "l", "((`l`.`l_shipdate`) <= \"1998-09-02\")", ["(`l`.`l_returnflag`)", "(`l`.`l_linestatus`)"], aggs)
    res, scan_ns, wall, ng = timed(qq)
    # torch reference on the same tensors
    cutoff = sum(1 for d in dates if d <= "1998-09-02")  # ranks below this pass (sorted dictionary)
    mask = ship < cutoff
    gid = (rf.long() * 2 + ls.long())[mask]
    cnt = torch.bincount(gid, minlength=6).cpu().numpy()
    sq = torch.zeros(6, dtype=torch.int64, device=DEV).index_add_(0, gid, qty[mask]).cpu().numpy()
    sp = torch.zeros(6, dtype=torch.float64, device=DEV).index_add_(0, gid, price[mask]).cpu().numpy()
    dp = price[mask] * (1 - disc[mask])
    sdp = torch.zeros(6, dtype=torch.float64, device=DEV).index_add_(0, gid, dp).cpu().numpy()
    sch = torch.zeros(6, dtype=torch.float64, device=DEV).index_add_(0, gid, dp * (1 + tax[mask])).cpu().numpy()
    rows = {tuple(k): a for k, a in res.rows()}
    assert len(rows) == int((cnt > 0).sum()), (len(rows), cnt)
    for r, rfv in enumerate("ANR"):
        for l, lsv in enumerate("FO"):
            i = r * 2 + l
            if cnt[i] == 0:
                continue
            a = rows[(rfv, lsv)]
            assert a[7] == int(cnt[i]) and a[0] == int(sq[i]), (rfv, lsv, a, cnt[i], sq[i])
            for got, want in ((a[1], sp[i]), (a[2], sdp[i]), (a[3], sch[i]), (a[5], sp[i] / cnt[i])):
                assert abs(got - want) <= 1e-9 * abs(want), (rfv, lsv, got, want)
            assert abs(a[4] - sq[i] / cnt[i]) <= 1e-12 * (sq[i] / cnt[i])
    line("config3", n, qq, res, scan_ns, wall, {"check": "6 groups: COUNT/SUM(int) exact, float SUM/AVG within 1e-9 of torch.index_add_ on the same tensors"})


# ---- config 4: 1 M groups, COUNT(DISTINCT) + SUM(DISTINCT) ---------------------------------------------------------------
def config4():
    n = int(200_000_000 * SCALE)
    ngroups = max(1000, int(1_000_000 * SCALE))
    g = torch.Generator(device=DEV).manual_seed(3)
    gg = torch.randint(0, ngroups, (n,), generator=g, device=DEV, dtype=torch.int64)
    xx = torch.randint(0, 1000, (n,), generator=g, device=DEV, dtype=torch.int64)
    t = q.Table(["g", "x"])
    t.set_column_device("g", gg)
    t.set_column_device("x", xx)
    t.seal()
    qq = q.Query(t, "d", None, ["(`d`.`g`)"], ["count(distinct (`d`.`x`))", "sum(distinct (`d`.`x`))", "count(*)"])
    res, scan_ns, wall, ng = timed(qq, reps=2)
    res._fetch()
    assert (res.key_cls[:, 0] == C_INT).all()
    keys = torch.from_numpy(res.key_val[:, 0].copy()).to(DEV)
    cnt = torch.bincount(gg, minlength=ngroups)
    assert ng == int((cnt > 0).sum()), (ng, int((cnt > 0).sum()))
    got_cnt = torch.from_numpy(ints(res, 2).copy()).to(DEV)
    assert torch.equal(got_cnt, cnt[keys]), "COUNT(*) per group"
    pairs = torch.unique(gg * 1000 + xx)
    del gg, xx
    pg = torch.div(pairs, 1000, rounding_mode="floor")
    cd = torch.bincount(pg, minlength=ngroups)
    sd = torch.zeros(ngroups, dtype=torch.int64, device=DEV).index_add_(0, pg, pairs - pg * 1000)
    assert torch.equal(torch.from_numpy(ints(res, 0).copy()).to(DEV), cd[keys]), "COUNT(DISTINCT x) per group"
    assert torch.equal(torch.from_numpy(ints(res, 1).copy()).to(DEV), sd[keys]), "SUM(DISTINCT x) per group"
    line("config4", n, qq, res, scan_ns, wall, {"distinct_entries": int(pairs.numel()), "hash_inserts_per_s": 2 * n / (scan_ns * 1e-9),
                                                "check": "every group: COUNT(*), COUNT(DISTINCT x), SUM(DISTINCT x) exact vs torch.unique/bincount on the same tensors"})


def tags_of(r, present):
    """class bytes: 10 % MISSING, 10 % NULL, else `present` (r uniform over 0..9)"""
    tg = torch.full(r.shape, present, dtype=torch.uint8, device=r.device)
    tg[r == 0] = C_MISSING
    tg[r == 1] = C_NULL
    return tg


# ---- config 5: Zipf string keys, 20 % MISSING/NULL -----------------------------------------------------------------------
CONFIG5_AGGS = ["count(*)", "count((`d`.`v`))", "sum((`d`.`v`))", "min((`d`.`v`))", "max((`d`.`v`))"]


def config5_table(n):
    """The config-5 keyspace partition of n rows, generated on the device: (table, words, device tensors)."""
    vocab = 100_000
    words = sorted("w%06d-%x" % (i, (i * 2654435761) & 0xffffff) for i in range(vocab))
    g = torch.Generator(device=DEV).manual_seed(4)
    # Zipf(s = 1.1) over the vocabulary by inverse CDF; popularity rank -> dictionary code through a fixed permutation
    w = torch.arange(1, vocab + 1, dtype=torch.float64, device=DEV).pow(-1.1)
    cdf = torch.cumsum(w / w.sum(), 0)
    perm = torch.randperm(vocab, generator=g, device=DEV).int()
    code = torch.empty(n, dtype=torch.int32, device=DEV)
    ktag = torch.empty(n, dtype=torch.uint8, device=DEV)
    v = torch.empty(n, dtype=torch.int64, device=DEV)
    vtag = torch.empty(n, dtype=torch.uint8, device=DEV)
    chunk = 50_000_000
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        u = torch.rand(m, generator=g, device=DEV, dtype=torch.float64)
        rank = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
        code[lo:lo + m] = perm[rank]
        r = torch.randint(0, 10, (m,), generator=g, device=DEV)
        ktag[lo:lo + m] = tags_of(r, C_STRING)
        v[lo:lo + m] = torch.randint(-1000, 1_000_000, (m,), generator=g, device=DEV, dtype=torch.int64)
        r = torch.randint(0, 10, (m,), generator=g, device=DEV)
        vtag[lo:lo + m] = tags_of(r, C_INT)
        del u, rank, r
    t = q.Table(["k", "v"])
    t.set_column_device("k", code, tags=ktag, dictionary=words)
    t.set_column_device("v", v, tags=vtag)
    t.set_global_rows(n)  # one partition: the exact row bound lets two row counters share a table word
    t.seal()
    return t, words, (code, ktag, v, vtag)


def config5_reference(words, tensors):
    """torch reductions over the same device tensors: group id 0 = MISSING key, 1 = NULL key, 2 + code = string"""
    code, ktag, v, vtag = tensors
    vocab = len(words)
    passing = vtag != C_MISSING
    gid = torch.where(ktag == C_STRING, code.long() + 2, ktag.long())[passing]
    isint = (vtag == C_INT)[passing]
    vv = v[passing]
    G = vocab + 2
    cnt = torch.bincount(gid, minlength=G)
    cntv = torch.bincount(gid[isint], minlength=G)
    sm = torch.zeros(G, dtype=torch.int64, device=DEV).index_add_(0, gid[isint], vv[isint])
    nneg = torch.bincount(gid[isint & (vv < 0)], minlength=G)
    mn = torch.full((G,), 2 ** 62, dtype=torch.int64, device=DEV).scatter_reduce_(0, gid[isint], vv[isint], "amin")
    mx = torch.full((G,), -2 ** 62, dtype=torch.int64, device=DEV).scatter_reduce_(0, gid[isint], vv[isint], "amax")
    return tuple(a.cpu().numpy() for a in (cnt, cntv, sm, nneg, mn, mx))


def config5_check(res, words, ref):
    cnt, cntv, sm, nneg, mn, mx = ref
    index = {wd: i + 2 for i, wd in enumerate(words)}
    rows = res.rows()
    assert len(rows) == int((cnt > 0).sum()), (len(rows), int((cnt > 0).sum()))
    for keys, a in rows:
        k = keys[0]
        i = 0 if k is q.MISSING else (1 if k is None else index[k])
        assert a[0] == cnt[i] and a[1] == cntv[i], (k, a, cnt[i], cntv[i])
        if cntv[i] == 0:
            assert a[2] is None and a[3] is None and a[4] is None, (k, a)
            continue
        # intValue.Add: a sum over ints of both signs is carried in float64 (value/integer.go:266-277)
        if nneg[i] == 0 or nneg[i] == cntv[i]:
            assert type(a[2]) is int and a[2] == sm[i], (k, a, sm[i])
        else:
            assert float(a[2]) == float(sm[i]), (k, a, sm[i])
        assert a[3] == mn[i] and a[4] == mx[i], (k, a, mn[i], mx[i])


def config5():
    n = int(1_000_000_000 * SCALE)
    t, words, tensors = config5_table(n)
    qq = q.Query(t, "d", "((`d`.`v`) is not missing)", ["(`d`.`k`)"], CONFIG5_AGGS)
    res, scan_ns, wall, ng = timed(qq)
    ref = config5_reference(words, tensors)
    config5_check(res, words, ref)
    cnt, cntv = ref[0], ref[1]
    line("config5", n, qq, res, scan_ns, wall,
         {"accumulator_updates_per_s": float(cnt.sum() + 4 * cntv.sum()) / (scan_ns * 1e-9),
          "check": "every group: COUNT(*), COUNT(v), SUM(v), MIN(v), MAX(v) exact vs torch bincount/index_add_/scatter_reduce_ on the same tensors"})


if __name__ == "__main__":
    q.init(0)
    todo = sys.argv[1:] or ["config3", "config4", "config5"]
    for name in todo:
        {"config3": config3, "config4": config4, "config5": config5}[name]()
        torch.cuda.empty_cache()
