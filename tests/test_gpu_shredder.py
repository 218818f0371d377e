"""GPU tests of the device-side JSON shredder (shred.cu, n1gpu_table_append_json with threads == -1): it must
produce exactly what the host shredder and the oracle's document model produce, including the fix-up rows
(escapes, long numbers), invalid documents, duplicate names and nested paths."""
import json

import numpy as np
import pytest

import query_b200 as q
from gen_n1 import QUERIES, F, make_docs
from golden_plans import CASES, keyspaces
from oracle import cref
from util_n1 import assert_same, gpu_rows, oracle_rows, run_both

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    q.init(0)


@pytest.mark.parametrize("name,where,keys,aggs", QUERIES, ids=[x[0] for x in QUERIES])
def test_matrix_with_device_shredder(name, where, keys, aggs):
    docs = make_docs(3000, seed=23)
    run_both(docs, "d", where, keys, aggs, name + " (device shredder)", threads=-1)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.id)
def test_goldens_with_device_shredder(case):
    docs = [t for _k, t in case.docs()]
    run_both(docs, case.alias, case.where, case.keys, sorted(set(case.aggs)), case.id, threads=-1)


EDGE_DOCS = [
    '{"a": 1, "a": 2, "s": "x"}',                 # duplicate name: first wins
    '  \n\t{"a": 1.0, "s": "y"}',                  # leading whitespace, integral float -> int
    '"hello"', '[1,2]', '17', '',                  # non-object documents
    '{"a": 1} trailing', '{"a": tru}', '{"a": 01}', '{"a": 1,}', '{"a" 1}', '{"a": [1,2}', '{"a": "unterminated}',
    '{"a": 1e2, "s": "z"}', '{"a": -0.0}', '{"a": 9223372036854775807}', '{"a": 9223372036854775808}',
    '{"a": -9223372036854775808}', '{"a": 0.1}', '{"a": 123456789.123456789}', '{"a": 1e-7}', '{"a": 2.5e10}',
    '{"a": 1.7976931348623157e308}', '{"a": 4.9e-324}', '{"a": 12345678901234567890}', '{"a": 0.30000000000000004}',
    '{"s": "\\u00e9\\ud83d\\ude00", "a": 3}', '{"\\u0061": 5}', '{"s": "tab\\there"}', '{"s": ""}',
    '{"c": {"b": 1}, "s": "obj"}', '{"c": [1, {"a": 5}], "s": "arr"}', '{"n": {"a": 7, "deep": {"a": 8}}, "a": 9}',
    '{"x": {"y": [1, 2, {"z": "q"}]}, "a": true, "s": null}', '{"a": false}', '{"a": null}',
    '{"s": "a"}', '{"s": "b"}', '{"s": "a"}', '{"s": "ab"}', '{"s": "B"}',
    '{ "a" : 4 , "s" : "spaced" }', '{"a":5,"s":"\\/"}',
]


def test_edge_documents_device_vs_host_vs_oracle():
    aggs = ["count(*)", "count((`d`.`a`))", "countn((`d`.`a`))", "sum((`d`.`a`))", "min((`d`.`a`))", "max((`d`.`a`))",
            "count((`d`.`s`))", "min((`d`.`s`))", "max((`d`.`s`))", "count(distinct (`d`.`s`))", "count(((`d`.`n`).`a`))",
            "sum((((`d`.`n`).`deep`).`a`))"]
    for keys in ([], ["(`d`.`s`)"], ["(`d`.`a`)"]):
        qh, rh = run_both(EDGE_DOCS, "d", None, keys, aggs, "edge host", threads=1)
        qd, rd = run_both(EDGE_DOCS, "d", None, keys, aggs, "edge device", threads=-1)
        assert_same(gpu_rows(rh, aggs), gpu_rows(rd, aggs), "host vs device shredder")
        assert qh.kernel_source == qd.kernel_source, "same statistics and dictionaries -> same kernel"


def test_device_shredder_config2_documents_1m():
    """BASELINE config-2 documents (the bench's e2e input) at 1 M rows: device shredder + scan vs the C oracle."""
    n = 1_000_000
    buf, offs = cref.gen_docs(2, 42, 0, n)
    where = "((`d`.`n`) between 250000 and 749999)"
    aggs = ["count(*)", "count((`d`.`n`))", "sum((`d`.`n`))", "avg((`d`.`n`))", "min((`d`.`n`))", "max((`d`.`n`))", "sum((`d`.`f`))",
            "count(distinct (`d`.`type`))", "max((`d`.`type`))"]
    t = q.Table(["n", "f", "type"])
    t.append_json((buf, offs), threads=-1)
    t.seal()
    got = gpu_rows(q.Query(t, "d", where, [], aggs).execute(), aggs)
    exp = cref.rows((buf, offs), "d", where, [], aggs, threads=8)
    assert_same(exp, got, "config2 1M")
    got = gpu_rows(q.Query(t, "d", where, ["(`d`.`type`)"], aggs[:7]).execute(), aggs[:7])
    exp = cref.rows((buf, offs), "d", where, ["(`d`.`type`)"], aggs[:7], threads=8)
    assert_same(exp, got, "config2 1M grouped by type")


@pytest.mark.parametrize("config,where,keys,aggs", [
    (3, "((`l`.`l_shipdate`) <= \"1998-09-02\")", ["(`l`.`l_returnflag`)", "(`l`.`l_linestatus`)"],
     ["sum((`l`.`l_quantity`))", "sum((`l`.`l_extendedprice`))", "sum(((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))))",
      "sum((((`l`.`l_extendedprice`) * (1 - (`l`.`l_discount`))) * (1 + (`l`.`l_tax`))))", "avg((`l`.`l_quantity`))",
      "avg((`l`.`l_extendedprice`))", "avg((`l`.`l_discount`))", "count(*)"]),
    (4, None, ["(`l`.`g`)"], ["count(distinct (`l`.`x`))", "sum(distinct (`l`.`x`))", "count(*)"]),
    (5, "((`l`.`v`) is not missing)", ["(`l`.`k`)"], ["count(*)", "count((`l`.`v`))", "sum((`l`.`v`))", "min((`l`.`v`))", "max((`l`.`v`))"]),
])
def test_baseline_config_shapes_300k(config, where, keys, aggs):
    """BASELINE.json configs 3-5 (Q1 shape, 1M-group DISTINCT shape, Zipf string keys with MISSING/NULL) at 300 k
    documents, device shredder, against the C oracle."""
    n = 300_000
    buf, offs = cref.gen_docs(config, 7, 0, n)
    import re
    from util_n1 import paths_of
    t = q.Table([list(p) for p in paths_of(where, keys, aggs)])
    t.append_json((buf, offs), threads=-1)
    t.seal()
    got = gpu_rows(q.Query(t, "l", where, keys, aggs).execute(), aggs)
    exp = cref.rows((buf, offs), "l", where, keys, aggs, threads=8)
    assert_same(exp, got, "config %d" % config)
