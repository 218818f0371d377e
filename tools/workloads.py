"""Device-generated keyspaces of BASELINE.json configs 2-5, their queries, and size-independent checks.

Every config is ONE global keyspace of `rows` documents whose shredded columns are generated on the device, chunk by
chunk (each chunk has its own seed), so rank r of w generates exactly rows [r*rows/w, (r+1)*rows/w) of the SAME
keyspace whatever w is: strong scaling keeps the data fixed.  A table is built with n1gpu_table_set_column_device (no
host staging of the 14 GB of config 5).

The oracle cannot run these sizes; a result is checked against torch reductions (bincount / index_add_ /
scatter_reduce_ / unique) over the regenerated chunks - exact for counts, integer sums, min / max, group sets and
DISTINCT sets; float64 sums of config 3 against EXACT integer-scaled totals (prices are whole cents, discounts and
taxes whole percents) within the north-star tolerance 1e-12.  With several ranks every rank reduces its own rows and the
partial references are all-reduced, so each rank can check the groups it finalised.

Used by bench.py, tools/full_size.py and tests/test_gpu_full_size.py.  Nothing here is on the product path."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import query_b200 as q  # noqa: E402

C_MISSING, C_NULL, C_INT, C_FLOAT, C_STRING = 0, 1, 4, 5, 6
CHUNK = 12_500_000  # rows per generation chunk: 1 B rows = 80 chunks, a multiple of 1, 2, 4, 8 ranks
REL_TOL = 1e-12     # BASELINE.json north_star tolerance for float64 SUM / AVG


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def world():
    d = _dist()
    return d.get_world_size() if d else 1


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def row_range(rows, r=None, w=None):
    r = rank() if r is None else r
    w = world() if w is None else w
    return rows * r // w, rows * (r + 1) // w


def _chunks(lo, hi):
    """(chunk index, first row of the chunk, slice inside the chunk) covering global rows [lo, hi)"""
    c = lo // CHUNK
    while c * CHUNK < hi:
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        yield c, slice(a - c * CHUNK, b - c * CHUNK)
        c += 1


def _gen(dev, seed, chunk):
    return torch.Generator(device=dev).manual_seed(seed * 1_000_003 + chunk)


def _tags_of(r, present):
    """class bytes: 10 % MISSING, 10 % NULL, else `present` (r uniform over 0..9)"""
    tg = torch.full(r.shape, present, dtype=torch.uint8, device=r.device)
    tg[r == 0] = C_MISSING
    tg[r == 1] = C_NULL
    return tg


def _allreduce(parts):
    """parts: [(tensor, 'sum' | 'min' | 'max')] reduced over the ranks in place"""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return
    ops = {"sum": d.ReduceOp.SUM, "min": d.ReduceOp.MIN, "max": d.ReduceOp.MAX}
    for t, op in parts:
        d.all_reduce(t, op=ops[op])


class Workload:
    name = ""
    rows = 0
    alias = "d"
    where = None
    keys = ()
    aggs = ()
    survey_bytes_per_row = 0  # SURVEY.md 8d accounting figure (always one class byte per referenced column)

    def __init__(self, rows=None, dev=None, scale=1.0):
        self.rows = int((rows if rows is not None else self.rows) * scale)
        self.dev = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
        self.lo, self.hi = row_range(self.rows)
        self.n = self.hi - self.lo

    def chunk(self, c):
        """the columns of generation chunk c (all CHUNK rows of it, or fewer at the end of the keyspace)"""
        raise NotImplementedError

    def chunk_rows(self, c):
        return min(CHUNK, self.rows - c * CHUNK)

    def my_chunks(self):
        for c, sl in _chunks(self.lo, self.hi):
            yield tuple(col[sl] for col in self.chunk(c))

    def columns(self):
        """this rank's rows of every column, concatenated: list of device tensors in self.chunk() order"""
        first = None
        at = 0
        for cols in self.my_chunks():
            if first is None:
                first = [torch.empty(self.n, dtype=c.dtype, device=self.dev) for c in cols]
            m = cols[0].shape[0]
            for dst, src in zip(first, cols):
                dst[at:at + m] = src
            at += m
        if first is None:
            first = [torch.empty(0, dtype=c.dtype, device=self.dev) for c in self.chunk(0)]
        return first

    def table(self):
        raise NotImplementedError

    def sealed_table(self):
        from query_b200 import dist as qd
        t = self.table()
        if world() > 1:
            qd.agree_dictionaries_and_stats(t)
        else:
            t.set_global_rows(self.rows)
        t.seal()
        return t

    def query(self, table):
        return q.Query(table, self.alias, self.where, list(self.keys), list(self.aggs))

    # reference(): dict of numpy arrays, identical on every rank; check(result, ref) asserts this rank's groups
    def reference(self):
        raise NotImplementedError

    def check(self, result, ref):
        raise NotImplementedError

    def group_ids(self, result):
        """dense group ids (into the reference arrays) of the groups this result holds"""
        raise NotImplementedError

    def check_partition(self, result, ref_live):
        """every live group finalised by exactly one rank: all-reduced per-group ownership counts == 1"""
        ids = torch.from_numpy(np.asarray(self.group_ids(result), dtype=np.int64)).to(self.dev)
        seen = torch.zeros(ref_live.shape[0], dtype=torch.int64, device=self.dev)
        if ids.numel():
            seen.index_add_(0, ids, torch.ones_like(ids))
        _allreduce([(seen, "sum")])
        live = torch.from_numpy(ref_live.astype(np.int64)).to(self.dev)
        assert torch.equal(seen, live), "%s: groups finalised %s times, expected once per live group" % (self.name, seen[seen != live][:8].tolist())


# ---- config 2: flat documents, range filter, ungrouped aggregates ----------------------------------------------------------
class Config2(Workload):
    name = "config2"
    rows = 10_000_000
    where = "((`d`.`n`) between 250000 and 749999)"
    aggs = ("count(*)", "count((`d`.`n`))", "sum((`d`.`n`))", "avg((`d`.`n`))", "min((`d`.`n`))", "max((`d`.`n`))", "sum((`d`.`f`))")
    survey_bytes_per_row = 18
    sql = "SELECT COUNT(*),COUNT(n),SUM(n),AVG(n),MIN(n),MAX(n),SUM(f) FROM d WHERE n BETWEEN 250000 AND 749999"

    def chunk(self, c):
        g = _gen(self.dev, 1, c)
        m = self.chunk_rows(c)
        n = torch.randint(0, 1_000_000, (m,), generator=g, device=self.dev, dtype=torch.int64)
        fi = torch.randint(1, 1_000_000, (m,), generator=g, device=self.dev, dtype=torch.int64)  # f = fi / 1e6 in (0, 1): never integral
        return n, fi

    def table(self):
        n, fi = self.columns()
        t = q.Table(["n", "f"])
        t.set_column_device("n", n)
        t.set_column_device("f", fi.double() / 1e6, tags=torch.full((self.n,), C_FLOAT, dtype=torch.uint8, device=self.dev))
        return t

    def reference(self):
        cnt = torch.zeros(1, dtype=torch.int64, device=self.dev)
        sm = torch.zeros(1, dtype=torch.int64, device=self.dev)
        fs = torch.zeros(1, dtype=torch.int64, device=self.dev)
        mn = torch.full((1,), 2 ** 62, dtype=torch.int64, device=self.dev)
        mx = torch.full((1,), -2 ** 62, dtype=torch.int64, device=self.dev)
        for n, fi in self.my_chunks():
            m = (n >= 250000) & (n <= 749999)
            cnt += m.sum()
            sm += n[m].sum()
            fs += fi[m].sum()
            if bool(m.any()):
                mn = torch.minimum(mn, n[m].min().reshape(1))
                mx = torch.maximum(mx, n[m].max().reshape(1))
        _allreduce([(cnt, "sum"), (sm, "sum"), (fs, "sum"), (mn, "min"), (mx, "max")])
        return {k: int(v.item()) for k, v in (("cnt", cnt), ("sum", sm), ("fsum_micro", fs), ("min", mn), ("max", mx))}

    def check(self, result, ref):
        rows = result.rows()
        assert len(rows) == 1, rows
        a = rows[0][1]
        assert a[0] == ref["cnt"] and a[1] == ref["cnt"] and a[2] == ref["sum"] and a[4] == ref["min"] and a[5] == ref["max"], (a, ref)
        avg = ref["sum"] / ref["cnt"]
        assert abs(a[3] - avg) <= REL_TOL * abs(avg), (a[3], avg)
        fs = ref["fsum_micro"] / 1e6  # exact integer total of millionths
        assert abs(a[6] - fs) <= REL_TOL * abs(fs), (a[6], fs)
        return "COUNT/SUM(int)/MIN/MAX exact; AVG and SUM(f) within 1e-12 of the exact integer-scaled totals"


# ---- config 3: TPC-H Q1 shape ---------------------------------------------------------------------------------------------
class Config3(Workload):
    name = "config3"
    rows = 60_000_000
    alias = "l"
    where = "((`l`.`l_shipdate`) <= \"1998-09-02\")"
    keys = ("(`l`.`l_returnflag`)", "(`l`.`l_linestatus`)")
    aggs = ("sum((`l`.`l_quantity`))", "sum((`l`.`l_extendedprice`))", "sum(((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))))",
            "sum((((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))) * (1 + (`l`.`l_tax`))))", "avg((`l`.`l_quantity`))",
            "avg((`l`.`l_extendedprice`))", "avg((`l`.`l_discount`))", "count(*)")
    survey_bytes_per_row = 51
    sql = ("SELECT l_returnflag,l_linestatus,SUM(l_quantity),SUM(l_extendedprice),SUM(l_extendedprice*(1-l_discount)),"
           "SUM(l_extendedprice*(1-l_discount)*(1+l_tax)),AVG(l_quantity),AVG(l_extendedprice),AVG(l_discount),COUNT(*) "
           "FROM lineitem WHERE l_shipdate <= \"1998-09-02\" GROUP BY l_returnflag,l_linestatus")
    DATES = ["%04d-%02d-%02d" % (y, m, d) for y in range(1992, 1999) for m in range(1, 13) for d in range(1, 29)]

    def chunk(self, c):
        g = _gen(self.dev, 2, c)
        m = self.chunk_rows(c)
        dev = self.dev
        ship = torch.randint(0, len(self.DATES), (m,), generator=g, device=dev, dtype=torch.int32)
        rf = torch.randint(0, 3, (m,), generator=g, device=dev, dtype=torch.int32)
        ls = torch.randint(0, 2, (m,), generator=g, device=dev, dtype=torch.int32)
        qty = torch.randint(1, 51, (m,), generator=g, device=dev, dtype=torch.int64)
        cents = torch.randint(90000, 10500000, (m,), generator=g, device=dev, dtype=torch.int64)  # whole cents: 1 % are whole dollars (INT)
        disc = torch.randint(0, 11, (m,), generator=g, device=dev, dtype=torch.int64)             # 0.00 .. 0.10: 0.00 is the INT 0
        tax = torch.randint(0, 9, (m,), generator=g, device=dev, dtype=torch.int64)               # 0.00 .. 0.08
        return ship, rf, ls, qty, cents, disc, tax

    def table(self):
        ship, rf, ls, qty, cents, disc, tax = self.columns()
        ftag = torch.full((self.n,), C_FLOAT, dtype=torch.uint8, device=self.dev)  # integral values are canonicalised to INT by the library
        t = q.Table(["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"])
        t.set_column_device("l_shipdate", ship, dictionary=self.DATES)
        t.set_column_device("l_returnflag", rf, dictionary=["A", "N", "R"])
        t.set_column_device("l_linestatus", ls, dictionary=["F", "O"])
        t.set_column_device("l_quantity", qty)
        t.set_column_device("l_extendedprice", cents.double() / 100.0, tags=ftag)
        t.set_column_device("l_discount", disc.double() / 100.0, tags=ftag)
        t.set_column_device("l_tax", tax.double() / 100.0, tags=ftag)
        return t

    def reference(self):
        dev = self.dev
        z = lambda: torch.zeros(6, dtype=torch.int64, device=dev)
        cnt, sq, sc, sdp, sch, sd = z(), z(), z(), z(), z(), z()
        cutoff = sum(1 for d in self.DATES if d <= "1998-09-02")  # ranks below this pass (sorted dictionary)
        for ship, rf, ls, qty, cents, disc, tax in self.my_chunks():
            m = ship < cutoff
            gid = (rf.long() * 2 + ls.long())[m]
            c, d, t, qv = cents[m], disc[m], tax[m], qty[m]
            cnt += torch.bincount(gid, minlength=6)
            sq.index_add_(0, gid, qv)
            sc.index_add_(0, gid, c)
            sdp.index_add_(0, gid, c * (100 - d))                 # cents * percent: exact in int64
            sch.index_add_(0, gid, c * (100 - d) * (100 + t))
            sd.index_add_(0, gid, d)
        _allreduce([(x, "sum") for x in (cnt, sq, sc, sdp, sch, sd)])
        return {k: v.cpu().numpy() for k, v in (("cnt", cnt), ("qty", sq), ("cents", sc), ("dp", sdp), ("ch", sch), ("disc", sd))}

    def _gid(self, keys):
        return "ANR".index(keys[0]) * 2 + "FO".index(keys[1])

    def group_ids(self, result):
        return [self._gid(k) for k, _a in result.rows()]

    def check(self, result, ref):
        rows = result.rows()
        seen = set()
        for keys, a in rows:
            i = self._gid(keys)
            assert i not in seen
            seen.add(i)
            n = int(ref["cnt"][i])
            assert a[7] == n and a[0] == int(ref["qty"][i]), (keys, a, n)
            want = (ref["cents"][i] / 100.0, int(ref["dp"][i]) / 1e4, int(ref["ch"][i]) / 1e6, None, int(ref["cents"][i]) / 100.0 / n,
                    int(ref["disc"][i]) / 100.0 / n)
            for got, w in ((a[1], want[0]), (a[2], want[1]), (a[3], want[2]), (a[5], want[4]), (a[6], want[5])):
                assert abs(got - w) <= REL_TOL * abs(w), (keys, got, w, abs(got - w) / abs(w))
            avgq = int(ref["qty"][i]) / n
            assert abs(a[4] - avgq) <= REL_TOL * avgq, (keys, a[4], avgq)
        if world() == 1:
            assert len(rows) == int((ref["cnt"] > 0).sum()), (len(rows), ref["cnt"])
        return "COUNT/SUM(int) exact; float SUM/AVG within 1e-12 of exact integer-scaled totals (whole cents x whole percents); 0.00 discounts / taxes and whole-dollar prices are INT rows"


# ---- config 4: 1 M groups, COUNT(DISTINCT) + SUM(DISTINCT) ------------------------------------------------------------------------
class Config4(Workload):
    name = "config4"
    rows = 200_000_000
    groups = 1_000_000
    keys = ("(`d`.`g`)",)
    aggs = ("count(distinct (`d`.`x`))", "sum(distinct (`d`.`x`))", "count(*)")
    survey_bytes_per_row = 18
    sql = "SELECT g,COUNT(DISTINCT x),SUM(DISTINCT x),COUNT(*) FROM d GROUP BY g"

    def __init__(self, rows=None, dev=None, scale=1.0):
        super().__init__(rows, dev, scale)
        self.ngroups = max(1000, int(self.groups * scale))

    def chunk(self, c):
        g = _gen(self.dev, 3, c)
        m = self.chunk_rows(c)
        gg = torch.randint(0, self.ngroups, (m,), generator=g, device=self.dev, dtype=torch.int64)
        xx = torch.randint(0, 1000, (m,), generator=g, device=self.dev, dtype=torch.int64)
        return gg, xx

    def table(self):
        gg, xx = self.columns()
        t = q.Table(["g", "x"])
        t.set_column_device("g", gg)
        t.set_column_device("x", xx)
        return t

    def reference(self):
        dev, G = self.dev, self.ngroups
        cnt = torch.zeros(G, dtype=torch.int64, device=dev)
        present = torch.zeros(G * 1000, dtype=torch.bool, device=dev)  # the DISTINCT set as a bitmap (1 GB at full size)
        for gg, xx in self.my_chunks():
            cnt += torch.bincount(gg, minlength=G)
            present[gg * 1000 + xx] = True
        if world() > 1:
            p8 = present.view(torch.uint8)
            _allreduce([(cnt, "sum"), (p8, "max")])
        pm = present.view(G, 1000)
        cd = pm.sum(1)
        sd = torch.empty(G, dtype=torch.int64, device=dev)
        ar = torch.arange(1000, device=dev, dtype=torch.int32)
        for a in range(0, G, 100_000):  # bounded temporaries
            sd[a:a + 100_000] = (pm[a:a + 100_000].to(torch.int32) * ar).sum(1)
        npairs = int(cd.sum().item())
        del present, pm
        return {"cnt": cnt.cpu().numpy(), "cd": cd.cpu().numpy(), "sd": sd.cpu().numpy(), "pairs": npairs}

    def group_ids(self, result):
        result._fetch()
        return result.key_val[:, 0]

    def check(self, result, ref):
        result._fetch()
        assert (result.key_cls[:, 0] == C_INT).all() and (result.agg_cls == C_INT).all()
        keys = result.key_val[:, 0]
        assert len(np.unique(keys)) == len(keys), "duplicate groups"
        assert (result.agg_val[:, 2] == ref["cnt"][keys]).all(), "COUNT(*) per group"
        assert (result.agg_val[:, 0] == ref["cd"][keys]).all(), "COUNT(DISTINCT x) per group"
        assert (result.agg_val[:, 1] == ref["sd"][keys]).all(), "SUM(DISTINCT x) per group"
        if world() == 1:
            assert len(keys) == int((ref["cnt"] > 0).sum()), (len(keys), int((ref["cnt"] > 0).sum()))
        return "every group: COUNT(*), COUNT(DISTINCT x), SUM(DISTINCT x) exact vs a torch bitmap of the (g, x) pairs"


# ---- config 5: Zipf string keys, 20 % MISSING / NULL -----------------------------------------------------------------------------
class Config5(Workload):
    name = "config5"
    rows = 1_000_000_000
    vocab = 100_000
    where = "((`d`.`v`) is not missing)"
    keys = ("(`d`.`k`)",)
    aggs = ("count(*)", "count((`d`.`v`))", "sum((`d`.`v`))", "min((`d`.`v`))", "max((`d`.`v`))")
    survey_bytes_per_row = 14
    sql = "SELECT k,COUNT(*),COUNT(v),SUM(v),MIN(v),MAX(v) FROM d WHERE v IS NOT MISSING GROUP BY k"

    def __init__(self, rows=None, dev=None, scale=1.0):
        super().__init__(rows, dev, scale)
        self.words = sorted("w%06d-%x" % (i, (i * 2654435761) & 0xffffff) for i in range(self.vocab))
        common = torch.Generator(device=self.dev).manual_seed(4)  # same on every rank: popularity rank -> dictionary code
        self.perm = torch.randperm(self.vocab, generator=common, device=self.dev).int()
        w = torch.arange(1, self.vocab + 1, dtype=torch.float64, device=self.dev).pow(-1.1)  # Zipf(s = 1.1) by inverse CDF
        self.cdf = torch.cumsum(w / w.sum(), 0)

    def chunk(self, c):
        g = _gen(self.dev, 4, c)
        m = self.chunk_rows(c)
        dev = self.dev
        u = torch.rand(m, generator=g, device=dev, dtype=torch.float64)
        code = self.perm[torch.searchsorted(self.cdf, u).clamp_(max=self.vocab - 1)]
        ktag = _tags_of(torch.randint(0, 10, (m,), generator=g, device=dev), C_STRING)
        v = torch.randint(-1000, 1_000_000, (m,), generator=g, device=dev, dtype=torch.int64)
        vtag = _tags_of(torch.randint(0, 10, (m,), generator=g, device=dev), C_INT)
        return code, ktag, v, vtag

    def table(self):
        code, ktag, v, vtag = self.columns()
        t = q.Table(["k", "v"])
        t.set_column_device("k", code, tags=ktag, dictionary=self.words)
        t.set_column_device("v", v, tags=vtag)
        return t

    def reference(self):
        """group id 0 = MISSING key, 1 = NULL key, 2 + code = string"""
        dev, G = self.dev, self.vocab + 2
        cnt = torch.zeros(G, dtype=torch.int64, device=dev)
        cntv, sm, nneg = torch.zeros_like(cnt), torch.zeros_like(cnt), torch.zeros_like(cnt)
        mn = torch.full((G,), 2 ** 62, dtype=torch.int64, device=dev)
        mx = torch.full((G,), -2 ** 62, dtype=torch.int64, device=dev)
        for code, ktag, v, vtag in self.my_chunks():
            passing = vtag != C_MISSING
            gid = torch.where(ktag == C_STRING, code.long() + 2, ktag.long())[passing]
            isint = (vtag == C_INT)[passing]
            vv = v[passing]
            gi, vi = gid[isint], vv[isint]
            cnt += torch.bincount(gid, minlength=G)
            cntv += torch.bincount(gi, minlength=G)
            sm.index_add_(0, gi, vi)
            nneg += torch.bincount(gi[vi < 0], minlength=G)
            mn.scatter_reduce_(0, gi, vi, "amin")
            mx.scatter_reduce_(0, gi, vi, "amax")
        _allreduce([(cnt, "sum"), (cntv, "sum"), (sm, "sum"), (nneg, "sum"), (mn, "min"), (mx, "max")])
        return {k: v.cpu().numpy() for k, v in (("cnt", cnt), ("cntv", cntv), ("sum", sm), ("nneg", nneg), ("min", mn), ("max", mx))}

    def _index(self):
        if not hasattr(self, "_idx"):
            self._idx = {wd: i + 2 for i, wd in enumerate(self.words)}
        return self._idx

    def group_ids(self, result):
        index = self._index()
        return [0 if k[0] is q.MISSING else (1 if k[0] is None else index[k[0]]) for k, _a in result.rows()]

    def check(self, result, ref):
        cnt, cntv, sm, nneg, mn, mx = (ref[k] for k in ("cnt", "cntv", "sum", "nneg", "min", "max"))
        index = self._index()
        rows = result.rows()
        seen = set()
        for keys, a in rows:
            k = keys[0]
            i = 0 if k is q.MISSING else (1 if k is None else index[k])
            assert i not in seen, "duplicate group %r" % (k,)
            seen.add(i)
            assert a[0] == cnt[i] and a[1] == cntv[i], (k, a, cnt[i], cntv[i])
            if cntv[i] == 0:
                assert a[2] is None and a[3] is None and a[4] is None, (k, a)
                continue
            # intValue.Add: a sum over ints of both signs is carried in float64 (value/integer.go:266-277)
            if nneg[i] == 0 or nneg[i] == cntv[i]:
                assert type(a[2]) is int and a[2] == sm[i], (k, a, sm[i])
            else:
                assert type(a[2]) is float and a[2] == float(sm[i]), (k, a, sm[i])
            assert a[3] == mn[i] and a[4] == mx[i], (k, a, mn[i], mx[i])
        if world() == 1:
            assert len(rows) == int((cnt > 0).sum()), (len(rows), int((cnt > 0).sum()))
        return "every group: COUNT(*), COUNT(v), SUM(v) (and its int / float class), MIN(v), MAX(v) exact vs torch bincount / index_add_ / scatter_reduce_"


CONFIGS = {"config2": Config2, "config3": Config3, "config4": Config4, "config5": Config5}
