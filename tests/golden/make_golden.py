#!/usr/bin/env python
"""Extracts the reference's own golden vectors for the filter + GROUP BY path into small
fixtures that travel with the repo (the GPU box has no /root/reference).

Run here (container with /root/reference):  python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/keyspaces.json.gz  documents, per keyspace, as [key, raw JSON text] in primary-key
                                  (sorted file name) order - datastore/file/file.go:711-730
  tests/golden/cases.json         verbatim {statements, results} entries of the reference case files

Sources:
  test/filestore/json/default/{catalog,jobs,user_profile,game,orders,contacts,tags,mixed}/*.json
  test/filestore/json/default/cases/{case_group_by_having,case_where,case_func_comp,case_integer}.json
  test/multistore/test_cases/aggregate_functions/{insert,case_group_by_having,case_distinct}.json
  test/multistore/test_cases/integers/{insert,case_select}.json
  test/multistore/test_cases/where_functions/{insert,case_where}.json
"""
import gzip
import json
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def file_keyspace(path):
    out = []
    for name in sorted(os.listdir(path)):  # ioutil.ReadDir sorts by file name
        p = os.path.join(path, name)
        if os.path.isdir(p):
            continue
        key = name[: len(name) - len(os.path.splitext(name)[1])]  # documentPathToId file.go:745-749
        with open(p, "rb") as f:
            out.append([key, f.read().decode("utf-8")])
    return out


_INS = re.compile(r'^\s*INSERT\s+INTO\s+(\w+)\s*\(KEY\s*,\s*VALUE\)\s*VALUES\s*\(\s*"([^"]+)"\s*,\s*(\{.*\})\s*\)\s*$',
                  re.I | re.S)


def insert_keyspaces(path):
    """multistore areas load their data with INSERT statements (insert.json)."""
    ks = {}
    for st in json.load(open(path)):
        m = _INS.match(st["statements"])
        if not m:
            continue
        name, key, doc = m.group(1), m.group(2), m.group(3)
        json.loads(doc)  # must be valid JSON as written
        ks.setdefault(name, []).append([key, doc])
    for v in ks.values():
        v.sort(key=lambda kv: kv[0])
    return ks


def cases(path):
    out = []
    for c in json.load(open(path)):
        out.append({k: c[k] for k in ("description", "statements", "results", "error") if k in c})
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    keyspaces = {}
    fs = os.path.join(REF, "test/filestore/json/default")
    for name in ("catalog", "jobs", "user_profile", "game", "orders", "contacts", "tags", "mixed", "non-json"):
        keyspaces["filestore/" + name] = file_keyspace(os.path.join(fs, name))
    ms = os.path.join(REF, "test/multistore/test_cases")
    for area in ("aggregate_functions", "integers", "where_functions"):
        for name, docs in insert_keyspaces(os.path.join(ms, area, "insert.json")).items():
            keyspaces["multistore/%s/%s" % (area, name)] = docs
    # config 1 of BASELINE.json: data/sampledb/dimestore/product (the same 900 products, no test_id)
    keyspaces["sampledb/dimestore/product"] = file_keyspace(os.path.join(REF, "data/sampledb/dimestore/product"))
    # ... and its second query: data/sampledb/dimestore/review (10 000 reviews, int `rating`), SURVEY.md 8d.1
    keyspaces["sampledb/dimestore/review"] = file_keyspace(os.path.join(REF, "data/sampledb/dimestore/review"))

    golden = {
        "filestore/case_group_by_having": cases(os.path.join(fs, "cases/case_group_by_having.json")),
        "filestore/case_where": cases(os.path.join(fs, "cases/case_where.json")),
        "filestore/case_func_comp": cases(os.path.join(fs, "cases/case_func_comp.json")),
        "filestore/case_integer": cases(os.path.join(fs, "cases/case_integer.json")),
        "multistore/aggregate_functions/case_group_by_having":
            cases(os.path.join(ms, "aggregate_functions/case_group_by_having.json")),
        "multistore/aggregate_functions/case_distinct":
            cases(os.path.join(ms, "aggregate_functions/case_distinct.json")),
        "multistore/integers/case_select": cases(os.path.join(ms, "integers/case_select.json")),
        "multistore/where_functions/case_where": cases(os.path.join(ms, "where_functions/case_where.json")),
    }
    with gzip.GzipFile(os.path.join(HERE, "keyspaces.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(keyspaces, sort_keys=True).encode("utf-8"))
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    for k, v in keyspaces.items():
        print("%-50s %5d docs" % (k, len(v)))
    for k, v in golden.items():
        print("%-60s %3d cases" % (k, len(v)))


if __name__ == "__main__":
    main()
