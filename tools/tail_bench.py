#!/usr/bin/env python
"""Host-side timing of the operator's tail (query_b200/csrc/group_tail.cpp) over a synthetic group result - no GPU
needed: N groups enter through n1gpu_operator_import_result, then HAVING + ORDER BY + LIMIT 10, and a full ORDER BY,
are timed around one n1gpu_operator_run_tail call each.  N1GPU_TRACE=1 prints the three phases.
Usage: python tools/tail_bench.py [ngroups]"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import query_b200 as q  # noqa: E402
from plans_n1 import explain_plan  # noqa: E402
from query_b200._lib import check, lib  # noqa: E402
from util_n1 import write_keyspace  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = tempfile.mkdtemp()
    write_keyspace(d, "default", "d", [("k1", '{"g":1,"x":2}')])
    keys, aggs = ["(`d`.`g`)"], sorted(["count(*)", "sum((`d`.`x`))", "max((`d`.`x`))"])
    terms = [("(`d`.`g`)", None), ("count(*)", "n"), ("(sum((`d`.`x`)) / count(*))", "mean"), ("max((`d`.`x`))", "mx")]
    order = [("`mean`", True), ("(`d`.`g`)", False)]
    rng = np.random.default_rng(1)
    kc, kv = np.full(n, 4, np.uint8), np.arange(n, dtype=np.int64)
    ac = np.full((n, 3), 4, np.uint8)
    av = np.stack([rng.integers(1, 400, n), rng.integers(0, 1000, n), rng.integers(0, 10 ** 6, n)], 1).astype(np.int64)
    for name, tail in (("HAVING + ORDER BY + LIMIT 10", dict(having="(count(*) > 100)", terms=terms, order=order, limit=10)),
                       ("ORDER BY, every row", dict(terms=terms, order=order))):
        op = q.Operator(explain_plan("default", "d", "d", None, keys, aggs, tail=tail), d, tail=True)
        res = op.import_arrays(kc, kv, ac, av, [])
        nn, nr = C.c_int64(), C.c_int64()
        t0 = time.time()
        check(lib().n1gpu_operator_run_tail(op._h, res._h, None, 0, C.byref(nn), C.byref(nr)))
        print(json.dumps({"tail": name, "groups": n, "seconds": round(time.time() - t0, 4), "rows": nr.value, "json_bytes": nn.value,
                          "host_threads": os.cpu_count()}))


if __name__ == "__main__":
    main()
