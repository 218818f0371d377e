"""the operators behind FinalGroup (LETTING, HAVING, projection, ORDER BY, OFFSET, LIMIT) taken
over by the substituted operator (n1gpu_plan_build_tail / n1gpu_operator_run_tail, query_b200/csrc/group_tail.cpp).

No GPU needed: the groups come from the oracle's chain and enter through n1gpu_operator_import_result, so this file checks
the host evaluator against (a) the reference's golden results and (b) the oracle's restatement of the same operators.
tests/test_gpu_parity.py runs the same plans end to end on the device."""
import json
import os

import numpy as np
import pytest

import query_b200 as q
from gen_n1 import F, make_docs
from golden_plans import CASES, normalise
from oracle import n1ql_oracle as O
from plans_n1 import distinct_plan, explain_plan
from util_n1 import write_keyspace

TAILED = [c for c in CASES if c.tail is not None]


def oracle_groups(docs, alias, where, keys, aggs):
    parsed = [O.parse_document(d) for d in docs]
    return O.run_chain(parsed, alias, where, keys, aggs)


def as_rows(groups, aggs):
    """oracle GroupRows -> the (keys, aggregates) rows n1gpu_operator_import_result takes"""
    conv = lambda v: q.MISSING if v is O.MISSING else v
    return [([conv(k) for k in g.keys], [conv(g.aggregates[a]) for a in aggs]) for g in groups]


def tie_groups(rows, keys):
    """rows in sorted order -> list of sets of rows with equal sort keys (the reference leaves ties unordered)"""
    out, last = [], object()
    for r, k in zip(rows, keys):
        k = json.dumps(k, sort_keys=True)
        if k != last:
            out.append([])
            last = k
        out[-1].append(json.dumps(normalise(r), sort_keys=True))
    return [sorted(g) for g in out]


@pytest.mark.parametrize("case", TAILED, ids=lambda c: c.id)
def test_golden_tails_through_the_operator(case, tmp_path):
    ns, ks = "default", case.keyspace.split("/")[-1]
    write_keyspace(str(tmp_path), ns, ks, case.docs())
    aggs = sorted(set(case.aggs))
    plan = explain_plan(ns, ks, case.alias, case.where, case.keys, aggs, tail=case.tail)
    op = q.Operator(plan, str(tmp_path), tail=True)
    want_ops = (["Filter"] if case.tail.get("having") else []) + ["InitialProject"]
    if case.tail.get("order"):
        want_ops += ["Order"] + (["Limit"] if case.tail.get("limit") is not None else []) + ["FinalProject"]
        assert op.rest_index == 6 and op.outer_rest_index == 1 + len(want_ops) - (2 if case.tail.get("having") else 1)
    else:
        want_ops += ["FinalProject"]
        assert op.rest_index == 6 and op.outer_rest_index == 0
    assert op.tail_operators == want_ops
    groups = oracle_groups([t for _k, t in case.docs()], case.alias, case.where, case.keys, aggs)
    got = op.run_tail(op.import_result(as_rows(groups, aggs)))
    assert normalise(got) == normalise(case.golden["results"]), case.golden["statements"]
    want = O.run_tail(groups, having=case.tail.get("having"), terms=case.tail["terms"], order=case.tail.get("order", ()),
                      limit=case.tail.get("limit"))
    assert normalise(got) == normalise(want)


def test_other_scalar_functions_keep_the_tail_with_the_caller(tmp_path):
    """Only ROUND (pinned by the reference's aggregate goldens) is evaluated in the tail; any other scalar function keeps
    the operators behind FinalGroup with the caller - the chain itself is still substituted."""
    case = CASES[0]
    ns, ks = "default", case.keyspace.split("/")[-1]
    write_keyspace(str(tmp_path), ns, ks, case.docs()[:20])
    aggs = sorted(set(case.aggs))
    for fn in ("abs(%s)", "ceil(%s)", "lower(%s)", "to_string(%s)"):
        tail = dict(terms=[(fn % aggs[0], "r")], order=[("`r`", False)])
        op = q.Operator(explain_plan(ns, ks, case.alias, case.where, case.keys, aggs, tail=tail), str(tmp_path), tail=True)
        assert op.tail_operators == [] and op.rest_index == 5 and op.outer_rest_index == 0


def test_round_half_to_even_and_propagation(tmp_path):
    """expression/func_num.go:1304-1336,1715-1736 through the tail: ROUND over group values, against the oracle."""
    docs = ['{"g": %d, "x": %s}' % (i % 7, v) for i, v in enumerate(
        ["0.5", "1.5", "2.5", "-0.5", "-1.5", "2.675", "1e300", "12345.678", "-7", "null", '"s"', "0.125", "0.375", "15"])]
    keys, aggs = ["(`d`.`g`)"], sorted({"sum((`d`.`x`))", "max((`d`.`x`))", "count(*)"})
    tail = dict(terms=[("(`d`.`g`)", None), ("round(sum((`d`.`x`)))", "r0"), ("round(sum((`d`.`x`)), 2)", "r2"),
                       ("round(sum((`d`.`x`)), -1)", "rm1"), ("round(max((`d`.`x`)), 1)", "rmax"),
                       ("round(sum((`d`.`x`)), 0.5)", "bad_digits"), ("round(sum((`d`.`x`)), `nothing`)", "unbound")],
                order=[("(`d`.`g`)", False)])
    tail_ok = dict(tail, terms=tail["terms"][:-1])  # `nothing` is no LETTING variable: that tail would be ineligible
    write_keyspace(str(tmp_path), "default", "d", [("k%03d" % i, t) for i, t in enumerate(docs)])
    assert q.Operator(explain_plan("default", "d", "d", None, keys, aggs, tail=tail), str(tmp_path), tail=True).tail_operators == []
    op = q.Operator(explain_plan("default", "d", "d", None, keys, aggs, tail=tail_ok), str(tmp_path), tail=True)
    groups = oracle_groups(docs, "d", None, keys, aggs)
    got = op.run_tail(op.import_result(as_rows(groups, aggs)))
    want = O.run_tail(groups, terms=tail_ok["terms"], order=tail_ok["order"])
    assert normalise(got) == normalise(want)
    consts = dict(terms=[("round(0.5)", "a"), ("round(1.5)", "b"), ("round(2.5)", "c"), ("round(-2.5)", "dd"), ("round(2.675, 2)", "e"),
                         ("round(1250, -2)", "f"), ("round(null)", "g"), ("round(\"x\")", "h"), ("round(missing)", "i")], limit=1)
    op = q.Operator(explain_plan("default", "d", "d", None, keys, aggs, tail=consts), str(tmp_path), tail=True)
    got = op.run_tail(op.import_result(as_rows(groups, aggs)))
    assert got == O.run_tail(groups, terms=consts["terms"], limit=1)
    assert got == [{"a": 0, "b": 2, "c": 2, "dd": -2, "e": 2.68, "f": 1200, "g": None, "h": None}]  # half to even; `i` MISSING: left out


def _mk(tmp_path, docs, where, keys, aggs, tail):
    write_keyspace(str(tmp_path), "default", "d", [("k%06d" % i, t) for i, t in enumerate(docs)])
    op = q.Operator(explain_plan("default", "d", "d", where, keys, aggs, tail=tail), str(tmp_path), tail=True)
    groups = oracle_groups(docs, "d", where, keys, aggs)
    return op, groups


TAILS = [
    ("having_arith_order_desc_limit",
     dict(having="((count(*) + 1) > 3)", terms=[(F("t"), "t"), ("count(*)", None), ("(sum(%s) / count(*))" % F("p"), "mean")],
          order=[("`mean`", True), (F("t"), False)], offset=1, limit=4)),
    ("letting_and_alias_order",
     dict(letting=[("n", "count(*)"), ("big", "(max(%s) * 2)" % F("p"))], having="((`n` >= 2) and (`big` is not null))",
          terms=[(F("t"), None), ("`n`", None), ("`big`", "twice"), ("(`big` - `n`)", None)], order=[("`twice`", False), (F("t"), True)])),
    ("missing_and_null_keys_collate",
     dict(terms=[(F("h"), "h"), ("count(*)", "c"), ("min(%s)" % F("s"), None)], order=[("`h`", False)])),
    ("between_in_not_on_groups",
     dict(having="((count(*) between 2 and 400) and (not (%s in [\"zzz\", null])))" % F("t"),
          terms=[(F("t"), None), ("(count(*) % 7)", "m"), ("(-sum(%s))" % F("p"), "neg")], order=[("`m`", False), ("`neg`", True)], limit=6)),
    ("order_only_by_aggregate_no_alias",
     dict(terms=[(F("t"), None), ("avg(%s)" % F("f"), None)], order=[("avg(%s)" % F("f"), True)], limit=3)),
    ("offset_beyond_and_zero_limit", dict(terms=[(F("t"), None)], order=[(F("t"), False)], offset=1000, limit=0)),
    ("limit_without_order", dict(terms=[("count(*)", "c")], limit=1)),
]


@pytest.mark.parametrize("name,tail", TAILS, ids=[t[0] for t in TAILS])
def test_synthetic_tails_against_the_oracle(name, tail, tmp_path):
    docs = make_docs(600, seed=31)
    keys = [F("h")] if name.startswith("missing") else ([] if name == "limit_without_order" else [F("t")])
    aggs = sorted({"count(*)", "sum(%s)" % F("p"), "max(%s)" % F("p"), "min(%s)" % F("s"), "avg(%s)" % F("f")})
    op, groups = _mk(tmp_path, docs, "(%s is not missing)" % F("p"), keys, aggs, tail)
    assert "InitialProject" in op.tail_operators and "FinalProject" in op.tail_operators, op.tail_operators
    got = op.run_tail(op.import_result(as_rows(groups, aggs)))
    want, sort_keys = O.run_tail(groups, letting=tail.get("letting", ()), having=tail.get("having"), terms=tail["terms"],
                                 order=tail.get("order", ()), offset=tail.get("offset"), limit=tail.get("limit"), with_sort_keys=True)
    assert len(got) == len(want)
    if tail.get("order") and tail.get("offset") is None and tail.get("limit") is None:
        got_keys = [k for _r, k in zip(got, sort_keys)]
        assert tie_groups(got, got_keys) == tie_groups(want, sort_keys)
    else:
        assert normalise(got) == normalise(want)


def test_ineligible_tail_shapes(tmp_path):
    docs = make_docs(50, seed=2)
    keys, aggs = [F("t")], ["count(*)"]
    write_keyspace(str(tmp_path), "default", "d", [("k%03d" % i, t) for i, t in enumerate(docs)])

    def ops(tail, mutate=None):
        plan = explain_plan("default", "d", "d", None, keys, aggs, tail=tail)
        if mutate:
            mutate(plan)
        return q.Operator(plan, str(tmp_path), tail=True).tail_operators

    assert ops(dict(terms=[(F("t"), None)])) == ["InitialProject", "FinalProject"]
    assert ops(dict(terms=[(F("p"), None)])) == []                       # not a group key
    assert ops(dict(terms=[("max(%s)" % F("p"), None)])) == []           # aggregate the group operators do not compute
    assert ops(dict(terms=[("`d`", None)])) == []                        # the whole document
    assert ops(dict(terms=[("(count(*) in %s)" % F("t"), "x")])) == []   # IN over a non-constructed array: decided at build time
    assert ops(dict(terms=[("[1, 2]", "x")])) == []                      # an array value of its own
    assert ops(dict(terms=[(F("t"), None)], limit="1.5")) == ["InitialProject", "FinalProject"]  # not integral: Limit stays with the caller
    proj = lambda p: p["~children"][0]["~children"][5]["~child"]["~children"]
    assert ops(dict(terms=[(F("t"), None)]), lambda p: proj(p)[0].update(distinct=True)) == []
    assert ops(dict(terms=[(F("t"), None)]), lambda p: proj(p)[0].update(raw=True)) == []
    assert ops(dict(terms=[(F("t"), None)]), lambda p: proj(p)[0]["result_terms"].append({"star": True})) == []
    # an alias that shadows the keyspace alias would change what a MISSING value resolves to in ORDER BY
    assert ops(dict(terms=[(F("t"), "d")], order=[("`d`", False)])) == []


DISTINCTS = [
    ("one_string", None, [(F("t"), None)], None, None),
    ("two_terms_where", "(%s < 500)" % F("p"), [(F("t"), "kind"), (F("h"), None)], None, None),
    ("computed_term_ordered", None, [("(%s %% 5)" % F("p"), "m"), (F("t"), None)], [("`m`", True), (F("t"), False)], 7),
    ("mixed_types_ordered", "(%s is not missing)" % F("p"), [(F("h"), "h")], [("`h`", False)], None),
]


@pytest.mark.parametrize("name,where,terms,order,limit", DISTINCTS, ids=[d[0] for d in DISTINCTS])
def test_select_distinct_is_a_group_by_over_the_projected_terms(name, where, terms, order, limit, tmp_path):
    """SELECT DISTINCT substituted as GROUP BY <terms> with no aggregates + the projection tail."""
    terms = [(e.replace("%%", "%"), a) for e, a in terms]
    docs = make_docs(700, seed=51)
    write_keyspace(str(tmp_path), "default", "d", [("k%06d" % i, t) for i, t in enumerate(docs)])
    plan = distinct_plan("default", "d", "d", where, terms, order=order, limit=limit)
    with pytest.raises(q.Ineligible):
        q.Operator(plan, str(tmp_path))  # without a tail there is nothing this operator could hand downstream
    op = q.Operator(plan, str(tmp_path), tail=True)
    assert op.tail_operators[0] == "InitialProject" and op.tail_operators.count("Distinct") == 2 and "FinalProject" in op.tail_operators
    assert op.rest_index == 4  # the parallel projection and the serial Distinct are both absorbed
    keys = sorted({e for e, _a in terms})
    groups = oracle_groups(docs, "d", where, keys, [])
    got = op.run_tail(op.import_result(as_rows(groups, [])))
    want = O.run_distinct([O.parse_document(d) for d in docs], "d", where, terms)
    canon = lambda rows: sorted(json.dumps(normalise(r), sort_keys=True) for r in rows)
    if limit is None:
        assert canon(got) == canon(want)
    else:
        assert len(got) == min(limit, len(want)) and set(canon(got)) <= set(canon(want))
    if order:  # sorted by the sort terms (all of them projected under explicit aliases or their own names here)
        names = [e.strip("`") if e.startswith("`") else [a or e for e2, a in terms if e2 == e][0] for e, _d in order]
        import functools
        from golden_plans import _cmp, _MISSING
        def cmp(a, b):
            for n_, (_e, desc) in zip(names, order):
                c = _cmp(a.get(n_, _MISSING), b.get(n_, _MISSING))
                if c:
                    return -c if desc else c
            return 0
        assert all(cmp(got[i], got[i + 1]) <= 0 for i in range(len(got) - 1))


def test_tail_over_empty_input(tmp_path):
    """group_final.go:108-117 + the tail: no documents and no GROUP BY still give one row of Default() values, which
    HAVING may drop; with GROUP BY there is no row at all."""
    os.makedirs(str(tmp_path / "default" / "d"))
    aggs = sorted({"count(*)", "sum((`d`.`p`))", "min((`d`.`p`))"})
    terms = [("count(*)", "n"), ("sum((`d`.`p`))", "s"), ("(min((`d`.`p`)) is null)", "nomin"), ("(count(*) + 1)", None)]
    for keys, having, want in (([], None, [{"n": 0, "s": None, "nomin": True, "$1": 1}]), ([], "(count(*) > 0)", []),
                               (["(`d`.`t`)"], None, [])):
        tail = dict(having=having, terms=terms, order=[("`n`", False)])
        op = q.Operator(explain_plan("default", "d", "d", None, keys, aggs, tail=tail), str(tmp_path), tail=True)
        groups = oracle_groups([], "d", None, keys, aggs)
        got = op.run_tail(op.import_result(as_rows(groups, aggs)))
        assert got == want == O.run_tail(groups, having=having, terms=terms, order=tail["order"])
