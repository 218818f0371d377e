#!/usr/bin/env python
"""Differential fuzz of the DEVICE expression semantics (n1ql_device.cuh through the generated scan kernel) against the
oracle: seeded random Filter conditions and aggregate operands over the synthetic documents of tests/gen_n1.py -
arithmetic with overflow promotion, cross-type collation, 4-valued logic, BETWEEN / IN / IS tests.  One query per round:
WHERE e0, aggregates count(*), count(e1), min(e2), max(e3), sum(e4) (e4 over well-conditioned columns only).
Expressions outside the eligible subset (e.g. string comparisons across columns) are skipped, not counted.

Usage: python tools/where_fuzz.py [rounds] [seed]   -> one summary JSON line; exits 1 on any mismatch."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import query_b200 as q  # noqa: E402
from gen_n1 import F, make_docs  # noqa: E402
from oracle import n1ql_oracle as O  # noqa: E402
from util_n1 import assert_same, gpu_rows, make_table, oracle_rows  # noqa: E402

NUM = ["i", "p", "n", "f", "g"]
ANY = NUM + ["s", "b", "m", "t"]


def const(rng, strings=True):
    r = rng.random()
    if r < 0.4:
        return str(rng.choice([0, 1, 2, -1, 3, 10, 50, 500, 2 ** 62, -7]))
    if r < 0.6:
        return repr(rng.choice([0.5, 1.5, -0.25, 100.125, 1e-3]))
    if r < 0.8 and strings:
        return json.dumps(rng.choice(["", "a", "abc", "b", "t3", "zeta"]))
    return rng.choice(["true", "false", "null", "missing"])


def arith(rng, depth, cols):
    if depth <= 0 or rng.random() < 0.3:
        return F(rng.choice(cols)) if rng.random() < 0.75 else const(rng, strings=False)
    k = rng.choice(["+", "*", "-", "/", "%", "neg"])
    a, b = arith(rng, depth - 1, cols), arith(rng, depth - 1, cols)
    if k in ("+", "*"):
        return "(%s %s %s)" % (a, k, b)
    if k == "neg":
        return "(-%s)" % a
    return "(%s %s %s)" % (a, k, b)


def cond(rng, depth):
    if depth <= 0 or rng.random() < 0.2:
        x = arith(rng, 1, ANY)
        k = rng.choice(["=", "<", "<=", "between", "in", "is"])
        if k == "between":
            return "(%s between %s and %s)" % (x, const(rng), const(rng))
        if k == "in":
            return "(%s in [%s])" % (x, ", ".join(const(rng) for _ in range(rng.choice([1, 2, 3]))))
        if k == "is":
            return "(%s is %s%s)" % (x, rng.choice(["", "not "]), rng.choice(["null", "missing", "valued"]))
        y = const(rng) if rng.random() < 0.6 else arith(rng, 1, NUM)
        return "(%s %s %s)" % ((x, k, y) if rng.random() < 0.5 else (y, k, x))
    k = rng.choice(["and", "or", "not"])
    if k == "not":
        return "(not %s)" % cond(rng, depth - 1)
    return "(" + (" %s " % k).join(cond(rng, depth - 1) for _ in range(rng.choice([2, 3]))) + ")"


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    q.init(0)
    docs = make_docs(3000, seed=77)
    done = skipped = 0
    failures = []
    while done < rounds and done + skipped < rounds * 6:
        where = cond(rng, 2)
        aggs = ["count(*)", "count(%s)" % arith(rng, 2, ANY), "min(%s)" % arith(rng, 2, NUM + ["b"]), "max(%s)" % arith(rng, 2, NUM),
                "sum(%s)" % arith(rng, 1, ["i", "p", "g", "f"])]
        try:
            where = str(O.parse(where))
            aggs = sorted({str(O.parse(a)) for a in aggs})
            t = make_table(docs, where, [], aggs)
            t.seal()
            qq = q.Query(t, "d", where, [], aggs)
        except (q.Ineligible, O.ParseError):
            skipped += 1
            continue
        got = gpu_rows(qq.execute(), aggs)
        exp = oracle_rows(docs, "d", where, [], aggs)
        try:
            assert_same(exp, got, where)
        except AssertionError as e:
            failures.append({"where": where, "aggs": aggs, "error": str(e)[:400]})
        done += 1
    print(json.dumps({"rounds": done, "skipped_ineligible": skipped, "failures": failures, "seed": seed}))
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
