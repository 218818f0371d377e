"""Pins the oracle (oracle/n1ql_oracle.py) against the reference's own golden vectors
(SURVEY.md 8c).  CPU only."""
import json

import pytest

from golden_plans import CASES, WHERE_CASES, _MISSING, normalise
from oracle import n1ql_oracle as O


def oracle_groups(case, streams=1):
    docs = [O.parse_document(text) for _key, text in case.docs()]
    rows = O.run_chain(docs, case.alias, case.where, case.keys, case.aggs, streams=streams)
    return [([_MISSING if k is O.MISSING else O.to_python(k) for k in g.keys],
             {a: O.to_python(v) for a, v in g.aggregates.items()}) for g in rows]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.id)
def test_oracle_matches_reference_golden(case):
    got = case.project(oracle_groups(case))
    assert normalise(got) == normalise(case.golden["results"]), case.golden["statements"]


@pytest.mark.parametrize("case", WHERE_CASES, ids=lambda c: c.id)
def test_oracle_filter_matches_reference_golden(case):
    got = case.project(oracle_groups(case))
    assert got == [{"n": len(case.golden["results"])}], case.golden["statements"]


@pytest.mark.parametrize("case", [c for c in CASES if "distinct" not in " ".join(c.aggs)], ids=lambda c: c.id)
def test_oracle_three_phase_merge_is_stream_count_invariant(case):
    """IntermediateGroup merge (group_intermediate.go:56-104): N initial streams == 1 stream, exactly
    for everything but float sums (tolerance 1e-12, the north_star's)."""
    one = {json.dumps(k, default=str): a for k, a in oracle_groups(case, 1)}
    three = {json.dumps(k, default=str): a for k, a in oracle_groups(case, 3)}
    assert one.keys() == three.keys()
    for k in one:
        for agg, v in one[k].items():
            w = three[k][agg]
            if isinstance(v, float) or isinstance(w, float):
                assert abs(v - w) <= 1e-12 * max(abs(v), abs(w))
            else:
                assert v == w


def test_case_integer_literals_and_arithmetic():
    """test/filestore/json/default/cases/case_integer.json:4-22 (int64 vs float64 literals, + - * neg)."""
    ev = lambda s: O.to_python(O.parse(s).evaluate({}))
    assert ev("9007199254740993.0") == 9007199254740992
    assert ev("9007199254740993") == 9007199254740993
    assert ev("(9007199254740993 + 0)") == 9007199254740993
    assert ev("(9007199254740993 * 1)") == 9007199254740993
    assert ev("(-9007199254740993)") == -9007199254740993
    assert ev("(9007199254740993 - 0)") == 9007199254740993
    assert ev("(5 / 2)") == 2.5
    assert ev("(5 / 0)") is None


def test_cross_type_collation_golden():
    """case_func_comp.json:123-137: LEAST("Yes",99)=99, GREATEST("Yes",99)="Yes" (string > number)."""
    assert O.collate("Yes", 99) > 0
    assert O.parse('(99 < "Yes")').evaluate({}) is True


def test_mixed_keyspace_collation_order():
    """test/filestore/json/default/mixed (21 docs): MIN/MAX over a column of every JSON type."""
    from golden_plans import keyspaces
    docs = [O.parse_document(t) for _k, t in keyspaces()["filestore/mixed"]]
    vals = [O.field(d, "field") for d in docs]
    present = [v for v in vals if O.vtype(v) > O.T_NULL]
    rows = O.run_chain(docs, "m", None, [], ["min((`m`.`field`))", "max((`m`.`field`))", "count((`m`.`field`))",
                                              "countn((`m`.`field`))", "count(*)"])
    a = rows[0].aggregates
    assert a["count(*)"] == 21
    assert a["count((`m`.`field`))"] == len(present)
    assert a["countn((`m`.`field`))"] == sum(1 for v in present if O.vtype(v) == O.T_NUMBER)
    lo = a["min((`m`.`field`))"]
    hi = a["max((`m`.`field`))"]
    assert all(O.collate(lo, v) <= 0 for v in present)
    assert all(O.collate(hi, v) >= 0 for v in present)
    assert O.vtype(lo) == O.T_BOOLEAN and O.vtype(hi) == O.T_OBJECT


def test_missing_and_null_are_distinct_groups():
    """group_util.go:28-30 (A.6 / A.8)."""
    docs = [O.parse_document(t) for t in ['{"k":null}', '{}', '{"k":1}', '{"k":1.0}', '{"k":null}', '{"x":2}']]
    rows = O.run_chain(docs, "d", None, ["(`d`.`k`)"], ["count(*)"])
    got = {("M" if r.keys[0] is O.MISSING else json.dumps(r.keys[0])): r.aggregates["count(*)"] for r in rows}
    assert got == {"null": 2, "M": 2, "1": 2}


def test_empty_input_defaults():
    """group_final.go:108-117."""
    rows = O.run_chain([], "d", None, [], ["count(*)", "sum((`d`.`n`))", "min((`d`.`n`))", "avg((`d`.`n`))"])
    assert len(rows) == 1
    assert rows[0].aggregates == {"count(*)": 0, "sum((`d`.`n`))": None, "min((`d`.`n`))": None, "avg((`d`.`n`))": None}
    assert O.run_chain([], "d", None, ["(`d`.`k`)"], ["count(*)"]) == []


def test_int_sum_semantics():
    """value/integer.go:266-277: same-sign sums stay int64 until overflow, mixed-sign adds go float."""
    s = lambda xs: O.run_chain([{"n": x} for x in xs], "d", None, [], ["sum((`d`.`n`))"])[0].aggregates["sum((`d`.`n`))"]
    assert s([1, 2, 3]) == 6 and isinstance(s([1, 2, 3]), int)
    assert isinstance(s([5, -3]), float) and s([5, -3]) == 2.0
    big = s([O.INT64_MAX, 1])
    assert isinstance(big, float) and big == 9.223372036854775808e18
    assert s([-5, -6]) == -11 and isinstance(s([-5, -6]), int)


def test_stringer_round_trip():
    for text in ["((`d`.`n`) between 10 and 20)", "(((`d`.`a`) + (`d`.`b`) + 1) < 5)", "(not ((`d`.`s`) = \"x\"))",
                 "((`d`.`n`) in [1, 2, 3])", "(-(`d`.`n`))", "count(distinct (`d`.`x`))", "count(*)",
                 "(((`d`.`a`).`b`) is not valued)", "sum(((`d`.`p`) * (1 - (`d`.`q`))))"]:
        assert str(O.parse(text)) == text.replace("[1, 2, 3]", "[1,2,3]")


@pytest.mark.parametrize("case", [c for c in CASES if c.tail is not None], ids=lambda c: c.id)
def test_oracle_tail_against_golden_rows(case):
    """the oracle's restatement of Let / Filter / InitialProject / Order / Offset / Limit /
    FinalProject (O.run_tail), fed with the operators' Stringer text, reproduces the reference's golden rows - order
    included.  This is what pins the checker of query_b200/csrc/group_tail.cpp."""
    docs = [O.parse_document(t) for _k, t in case.docs()]
    groups = O.run_chain(docs, case.alias, case.where, case.keys, sorted(set(case.aggs)))
    rows = O.run_tail(groups, having=case.tail.get("having"), terms=case.tail["terms"], order=case.tail.get("order", ()),
                      limit=case.tail.get("limit"))
    assert normalise(rows) == normalise(case.golden["results"]), case.golden["statements"]


def test_oracle_tail_scope_rules():
    """project_initial.go:98-117 + let.go:53-60: LETTING bindings see the item, not each other; an explicit alias is
    visible to ORDER BY; a MISSING projection value leaves its field out (object.go:246-255)."""
    docs = [{"k": "a", "x": 1}, {"k": "a", "x": 5}, {"k": "b", "x": 2}, {"x": 9}]
    groups = O.run_chain(docs, "d", None, ["(`d`.`k`)"], ["count(*)", "sum((`d`.`x`))"])
    rows = O.run_tail(groups, letting=[("n", "count(*)"), ("m", "(`n` + 1)")], terms=[("(`d`.`k`)", None), ("`n`", None), ("`m`", "mm"),
                                                                                     ("sum((`d`.`x`))", "s")], order=[("`s`", True)])
    assert rows == [{"n": 1, "s": 9}, {"k": "a", "n": 2, "s": 6}, {"k": "b", "n": 1, "s": 2}]  # `m` saw no `n`: MISSING + 1
