"""Shared helpers for the parity tests: run a chain through the oracle and through libn1gpu and compare."""
from __future__ import annotations

import json
import math
import os

import query_b200 as q
from oracle import n1ql_oracle as O

REL_TOL = 1e-12  # BASELINE.json north_star: float64 SUM/AVG within 1e-12 relative; everything else bit-exact


def paths_of(where, keys, aggs):
    """Field paths (dotted) referenced by Stringer-text expressions, via the oracle's parser."""
    paths = []

    def walk(e):
        if isinstance(e, O.Field):
            p, x = [], e
            while isinstance(x, O.Field):
                p.append(x.name)
                x = x.first
            paths.append(tuple(reversed(p)))
        else:
            for c in e.children():
                walk(c)

    for text in ([where] if where else []) + list(keys) + list(aggs):
        walk(O.parse(text))
    return sorted(set(paths))


def make_table(docs, where, keys, aggs, threads=0):
    t = q.Table([list(p) for p in paths_of(where, keys, aggs)])
    t.append_json(docs, threads=threads)
    return t


def oracle_rows(docs, alias, where, keys, aggs, streams=1):
    parsed = [O.parse_document(d) for d in docs]
    rows = O.run_chain(parsed, alias, where, keys, aggs, streams=streams)
    out = {}
    for g in rows:
        k = tuple("MISSING" if v is O.MISSING else json.dumps(O.to_python(v), sort_keys=True) for v in g.keys)
        assert k not in out
        out[k] = dict(g.aggregates)  # raw oracle values: int (intValue) and float (floatValue) stay distinct
    return out


def gpu_rows(result, aggs):
    out = {}
    for ks, ag in result.rows():
        k = tuple("MISSING" if v is q.MISSING else json.dumps(_norm(v), sort_keys=True) for v in ks)
        assert k not in out, "duplicate group %r" % (k,)
        out[k] = dict(zip(aggs, ag))
    return out


def _norm(v):
    """What json.loads of the rendered value holds: an integral float renders as an integer."""
    if isinstance(v, float) and not isinstance(v, bool):
        if math.isfinite(v) and v == int(v):
            return O.to_python(v)  # the integer the shortest-digits rendering reads back as
    return v


def _is_num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def assert_same(expected, got, what=""):
    """Bit-exact for COUNT / MIN / MAX / integer SUM / strings / booleans, and the int-vs-float class of a SUM
    must agree (value/integer.go:266-277); float64 SUM and every AVG (a float64 division, then NewValue)
    within 1e-12 relative."""
    assert set(expected.keys()) == set(got.keys()), "%s group keys differ:\n only oracle: %s\n only gpu: %s" % (
        what, sorted(set(expected) - set(got))[:5], sorted(set(got) - set(expected))[:5])
    for k, aggs in expected.items():
        for a, v in aggs.items():
            w = got[k][a]
            msg = "%s group %s %s: oracle %r gpu %r" % (what, k, a, v, w)
            if a.startswith("avg(") and _is_num(v) and _is_num(w):
                assert abs(v - w) <= REL_TOL * max(abs(v), abs(w)), msg
            elif isinstance(v, float) and isinstance(w, float):
                assert v == w or abs(v - w) <= REL_TOL * max(abs(v), abs(w)), msg
            else:
                assert type(v) is type(w) and v == w, msg


def run_both(docs, alias, where, keys, aggs, what="", threads=0):
    """threads=-1 shreds on the device (shred.cu), >= 0 with host threads"""
    t = make_table(docs, where, keys, aggs, threads=threads)
    t.seal()
    qq = q.Query(t, alias, where, keys, aggs)
    res = qq.execute()
    got = gpu_rows(res, aggs)
    exp = oracle_rows(docs, alias, where, keys, aggs)
    # the planner de-duplicates aggregates; the oracle keys by text, so duplicates collapse there
    assert_same(exp, got, what)
    return qq, res


def write_keyspace(root, namespace, keyspace, docs):
    """Materialises a file-datastore keyspace <root>/<namespace>/<keyspace>/<key>.json (datastore/file)."""
    d = os.path.join(root, namespace, keyspace)
    os.makedirs(d, exist_ok=True)
    for key, text in docs:
        with open(os.path.join(d, key + ".json"), "w", encoding="utf-8") as f:
            f.write(text)
    return d
