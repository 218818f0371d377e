"""Seeded synthetic documents + a query matrix covering the eligible subset."""
from __future__ import annotations

import json
import random

VOCAB = ["", "a", "ab", "abc", "b", "ba", "zeta", "Zeta", "é", "日本", "x y", "q\"uote", "tab\there"]


def make_docs(n, seed=1, missing_rate=0.15):
    rnd = random.Random(seed)
    docs = []
    for i in range(n):
        d = {}

        def maybe(name, fn):
            r = rnd.random()
            if r < missing_rate * 0.5:
                return  # MISSING
            if r < missing_rate:
                d[name] = None
                return
            d[name] = fn()

        maybe("i", lambda: rnd.randint(-50, 50))
        maybe("p", lambda: rnd.randint(0, 1000))
        maybe("n", lambda: rnd.choice([rnd.randint(-2 ** 62, 2 ** 62), rnd.randint(-10 ** 6, 10 ** 6), 2 ** 53 + rnd.randint(0, 9)]))
        maybe("f", lambda: rnd.choice([rnd.uniform(-100, 100), float(rnd.randint(-5, 5)), rnd.random() * 1e-3, 0.5, -0.0]))
        maybe("s", lambda: rnd.choice(VOCAB))
        maybe("b", lambda: rnd.random() < 0.5)
        maybe("m", lambda: rnd.choice([rnd.randint(-3, 3), rnd.uniform(-3, 3), rnd.choice(VOCAB), True, False, 2.0, 1e300, -1e300]))
        maybe("g", lambda: rnd.randint(0, 9))
        maybe("h", lambda: rnd.randint(0, 10 ** 9))
        maybe("t", lambda: rnd.choice(["t%d" % k for k in range(16)]))
        if rnd.random() > missing_rate:
            d["nest"] = {"k": rnd.choice(VOCAB[:6]), "v": rnd.randint(0, 3), "deep": {"x": rnd.uniform(0, 1)}}
        elif rnd.random() < 0.3:
            d["nest"] = rnd.choice([5, "str", None])  # field access on a non-object -> MISSING
        docs.append(json.dumps(d, ensure_ascii=rnd.random() < 0.5))
    return docs


def F(name, alias="d"):
    parts = name.split(".")
    e = "`%s`" % alias
    for p in parts:
        e = "(%s.`%s`)" % (e, p)
    return e


ALL_AGGS_ON = lambda x: ["count(*)", "count(%s)" % x, "countn(%s)" % x, "sum(%s)" % x, "avg(%s)" % x, "min(%s)" % x, "max(%s)" % x]

# (name, where, group keys, aggregates)
QUERIES = [
    ("ungrouped_all_i", None, [], ALL_AGGS_ON(F("i"))),
    ("ungrouped_all_n_bigint", None, [], ALL_AGGS_ON(F("n"))),
    ("ungrouped_all_f", None, [], ALL_AGGS_ON(F("f"))),
    ("ungrouped_all_mixed", None, [], ALL_AGGS_ON(F("m"))),
    ("ungrouped_all_string", None, [], ALL_AGGS_ON(F("s"))),
    ("ungrouped_all_bool", None, [], ALL_AGGS_ON(F("b"))),
    ("between_ints", "(%s between -10 and 20)" % F("i"), [], ALL_AGGS_ON(F("i")) + ["sum(%s)" % F("f")]),
    ("between_mixed_bounds", "(%s between 0.5 and \"b\")" % F("m"), [], ["count(*)", "min(%s)" % F("m"), "max(%s)" % F("m")]),
    ("lt_gt", "((%s < 10) and (-5 < %s))" % (F("i"), F("i")), [], ["count(*)", "sum(%s)" % F("i")]),
    ("le_float_const", "(%s <= 2.5)" % F("f"), [], ["count(*)", "sum(%s)" % F("f"), "avg(%s)" % F("f")]),
    ("eq_string", "(%s = \"ab\")" % F("s"), [], ["count(*)"]),
    ("eq_string_absent", "(%s = \"nope\")" % F("s"), [], ["count(*)", "sum(%s)" % F("i")]),
    ("lt_string_absent", "(%s < \"aba\")" % F("s"), [], ["count(*)", "max(%s)" % F("s")]),
    ("ne_string", "(not (%s = \"ab\"))" % F("s"), [], ["count(*)"]),
    ("in_ints", "(%s in [1, 2, 3, 50])" % F("i"), [], ["count(*)", "sum(%s)" % F("i")]),
    ("in_mixed", "(%s in [1, \"a\", true, 2.0, null])" % F("m"), [], ["count(*)"]),
    ("not_in", "(not (%s in [1, 2, 3]))" % F("i"), [], ["count(*)"]),
    ("or_and_not", "(((%s < 0) or (%s = \"a\")) and (not (%s is null)))" % (F("i"), F("s"), F("b")), [], ["count(*)"]),
    ("is_tests", "((%s is not missing) and (%s is valued))" % (F("i"), F("s")), [], ["count(*)"]),
    ("is_null_or_missing", "((%s is null) or (%s is missing))" % (F("i"), F("s")), [], ["count(*)", "count(%s)" % F("i")]),
    ("is_not_valued", "(%s is not valued)" % F("m"), [], ["count(*)"]),
    ("truth_of_column", F("b"), [], ["count(*)"]),
    ("truth_of_number", F("i"), [], ["count(*)"]),
    ("truth_of_string", F("s"), [], ["count(*)"]),
    ("truth_of_mixed", F("m"), [], ["count(*)"]),
    ("arith_add_filter", "((%s + %s) < 30)" % (F("i"), F("p")), [], ["count(*)", "sum((%s + %s))" % (F("i"), F("p"))]),
    ("arith_sub_mult", "(((%s - %s) * 2) <= 100)" % (F("p"), F("i")), [], ["count(*)", "sum(((%s - %s) * 2))" % (F("p"), F("i")), "min((%s * %s))" % (F("i"), F("i")), "max((-%s))" % F("i")]),
    ("arith_div_mod", "((%s / 4) < 100)" % F("p"), [], ["count(*)", "sum((%s / 4))" % F("p"), "sum((%s %% 7))" % F("p"), "avg((%s / %s))" % (F("p"), F("i"))]),
    ("arith_float", None, [], ["sum((%s * (1 - %s)))" % (F("f"), F("f")), "sum(((%s * (1 - %s)) * (1 + %s)))" % (F("p"), F("f"), F("f")), "avg((%s + 0.25))" % F("f")]),
    ("arith_overflow", None, [], ["sum((%s * %s))" % (F("n"), F("n")), "sum((%s + %s))" % (F("n"), F("n")), "max((%s * 4))" % F("n"), "min((-%s))" % F("n")]),
    ("arith_on_strings_null", "((%s + 1) is null)" % F("s"), [], ["count(*)"]),
    ("cross_column_numbers", "(%s < %s)" % (F("i"), F("p")), [], ["count(*)"]),
    ("nested_paths", "(%s < 2)" % F("nest.v"), [F("nest.k")], ["count(*)", "sum(%s)" % F("nest.deep.x"), "max(%s)" % F("nest.v")]),
    ("group_small_int", None, [F("g")], ALL_AGGS_ON(F("i"))),
    ("group_small_int_filter", "(%s between 100 and 900)" % F("p"), [F("g")], ALL_AGGS_ON(F("f"))),
    ("group_string", None, [F("t")], ALL_AGGS_ON(F("p"))),
    ("group_string_with_empty", None, [F("s")], ["count(*)", "sum(%s)" % F("i")]),
    ("group_bool", None, [F("b")], ["count(*)", "avg(%s)" % F("f")]),
    ("group_mixed_key", None, [F("m")], ["count(*)", "sum(%s)" % F("p")]),
    ("group_two_keys", None, [F("g"), F("t")], ["count(*)", "sum(%s)" % F("i"), "min(%s)" % F("s")]),
    ("group_three_keys", "(%s is valued)" % F("g"), [F("g"), F("b"), F("nest.k")], ["count(*)", "max(%s)" % F("f")]),
    ("group_high_card", None, [F("h")], ["count(*)", "sum(%s)" % F("p"), "min(%s)" % F("i")]),
    ("group_high_card_two", None, [F("h"), F("i")], ["count(*)", "max(%s)" % F("p")]),
    ("group_bigint_key", None, [F("n")], ["count(*)", "sum(%s)" % F("i")]),
    ("group_float_key", None, [F("f")], ["count(*)", "sum(%s)" % F("i")]),
    ("group_wide_keys_128", None, [F("n"), F("h")], ["count(*)", "sum(%s)" % F("p")]),
    ("group_float_and_int_128", None, [F("f"), F("g")], ["count(*)", "sum(%s)" % F("p")]),
    ("group_expr_key", None, ["(%s %% 3)" % F("p")], ["count(*)", "sum(%s)" % F("p")]),
    ("group_expr_key_add", None, ["(%s + %s)" % (F("g"), F("g"))], ["count(*)"]),
    ("distinct_ungrouped", None, [], ["count(distinct %s)" % F("i"), "countn(distinct %s)" % F("m"), "sum(distinct %s)" % F("p"),
                                      "avg(distinct %s)" % F("i"), "count(distinct %s)" % F("s"), "count(distinct %s)" % F("m")]),
    ("distinct_grouped", None, [F("g")], ["count(distinct %s)" % F("i"), "sum(distinct %s)" % F("i"), "count(*)", "avg(distinct %s)" % F("f"),
                                          "count(distinct %s)" % F("b")]),
    ("distinct_grouped_filter", "(%s < 500)" % F("p"), [F("t")], ["count(distinct %s)" % F("p"), "sum(distinct %s)" % F("p"), "countn(distinct %s)" % F("s")]),
    ("distinct_high_card_group", None, [F("h")], ["count(distinct %s)" % F("i"), "count(*)"]),
    ("distinct_bigint", None, [F("g")], ["count(distinct %s)" % F("n"), "sum(distinct %s)" % F("n")]),
    ("distinct_expr", None, [], ["count(distinct (%s %% 5))" % F("p"), "sum(distinct (%s / 2))" % F("i")]),
    ("where_false_everywhere", "(%s = \"never\")" % F("t"), [F("g")], ["count(*)"]),
    ("where_false_ungrouped", "(%s = \"never\")" % F("t"), [], ["count(*)", "sum(%s)" % F("i"), "min(%s)" % F("i"), "avg(%s)" % F("i"), "count(distinct %s)" % F("i")]),
    ("count_constant", "(90 < %s)" % F("p"), [], ["count(1)", "sum(2)", "max((\"z\" < %s))" % F("s")]),
    ("const_const_compare", "((1 < 2) and (\"a\" < \"b\"))", [], ["count(*)"]),
]
# max("z" < s) applies MAX to a boolean expression; keep it: MIN/MAX collate any type.


def config5_docs(n, vocab, seed):
    """BASELINE config 5 shape at oracle size: Zipf(1.1) string keys, 10 % MISSING + 10 % null on k and on v"""
    import numpy as np
    rng = np.random.default_rng(seed)
    w = np.arange(1, vocab + 1, dtype=np.float64) ** -1.1
    ranks = rng.choice(vocab, size=n, p=w / w.sum())
    perm = rng.permutation(vocab)
    docs = []
    for i in range(n):
        parts = []
        r = rng.integers(0, 10)
        if r == 1:
            parts.append('"k": null')
        elif r > 1:
            parts.append('"k": "w%05d"' % perm[ranks[i]])
        r = rng.integers(0, 10)
        if r == 1:
            parts.append('"v": null')
        elif r > 1:
            parts.append('"v": %d' % rng.integers(-1000, 1000000))
        docs.append("{" + ", ".join(parts) + "}")
    return docs


def config4_docs(n, groups, seed):
    """BASELINE config 4 shape at oracle size: many groups, few distinct values per group"""
    import numpy as np
    rng = np.random.default_rng(seed)
    g = rng.integers(0, groups, n)
    x = rng.integers(0, 1000, n)
    return ['{"g": %d, "x": %d, "y": %.3f}' % (g[i], x[i], rng.random() * 1000) for i in range(n)]
