"""Golden plan cases: each reference golden statement (tests/golden/cases.json, extracted verbatim
from /root/reference by tests/golden/make_golden.py) restated as the operator chain the reference
planner builds for it (planner/build_select_sub.go:23-296): Filter condition, group keys and
aggregates in Stringer text (what plan/filter.go / plan/group.go marshal), plus the projection /
HAVING / ORDER BY / LIMIT that the operators *after* FinalGroup apply, done here in Python.

Used by tests/test_oracle_golden.py (oracle vs golden) and tests/test_gpu_golden.py (CUDA path vs
golden and vs oracle).
"""
from __future__ import annotations

import functools
import gzip
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))


@functools.lru_cache(maxsize=None)
def keyspaces():
    with gzip.open(os.path.join(HERE, "golden", "keyspaces.json.gz"), "rb") as f:
        return json.loads(f.read().decode("utf-8"))


@functools.lru_cache(maxsize=None)
def golden_cases():
    with open(os.path.join(HERE, "golden", "cases.json")) as f:
        return json.load(f)


def round_float(x, prec):
    """expression/func_num.go:1715-1736 roundFloat, then value.NewValue (integral -> int)."""
    if x is None:
        return None
    x = float(x)
    sign = 1.0
    if x < 0:
        sign, x = -1.0, -x
    p = math.pow(10, float(prec))
    inter = x * p + 0.5
    r = math.floor(inter)
    if r == inter and math.fmod(r, 2) != 0:
        r -= 1
    v = sign * r / p
    return int(v) if v == int(v) else v


class Case:
    def __init__(self, cid, ref, casefile, index, keyspace, alias, where, keys, aggs, select,
                 having=None, order=None, limit=None, key_names=None, tail=None):
        self.id, self.ref, self.casefile, self.index = cid, ref, casefile, index
        self.keyspace, self.alias = keyspace, alias
        self.where, self.keys, self.aggs = where, keys, aggs
        self.select, self.having, self.order, self.limit = select, having, order, limit
        # the same HAVING / projection / ORDER BY / LIMIT as the reference planner formalises them (Stringer text of the
        # operators behind FinalGroup): dict(having=, terms=[(expr, as)], order=[(expr, desc)], limit=).  None: the
        # statement is checked through a stand-in aggregate (ms_totcolors), so its own projection does not apply.
        self.tail = tail

    @property
    def golden(self):
        return golden_cases()[self.casefile][self.index]

    def docs(self):
        return keyspaces()[self.keyspace]

    def project(self, groups):
        """groups: list of (keys list [python values or MISSING marker], {agg text: python value}).
        Applies HAVING, the SELECT list, ORDER BY (N1QL collation on the python values) and LIMIT."""
        rows = []
        for keys, aggs in groups:
            if self.having is not None and not self.having(keys, aggs):
                continue
            row = {}
            for name, fn in self.select:
                v = fn(keys, aggs)
                if v is not _MISSING:
                    row[name] = v
            rows.append(row)
        if self.order:
            rows.sort(key=functools.cmp_to_key(lambda a, b: _cmp_rows(a, b, self.order)))
        if self.limit is not None:
            rows = rows[: self.limit]
        return rows


_MISSING = object()


def _rank(v):
    if v is _MISSING:
        return 0
    if v is None:
        return 1
    if isinstance(v, bool):
        return 2
    if isinstance(v, (int, float)):
        return 3
    if isinstance(v, str):
        return 4
    if isinstance(v, list):
        return 5
    return 6


def _cmp(a, b):
    ra, rb = _rank(a), _rank(b)
    if ra != rb:
        return ra - rb
    if ra <= 1:
        return 0
    if ra == 4:
        a, b = a.encode(), b.encode()
    if ra == 5:
        for x, y in zip(a, b):
            c = _cmp(x, y)
            if c:
                return c
        return len(a) - len(b)
    return -1 if a < b else (1 if a > b else 0)


def _cmp_rows(a, b, order):
    for name in order:
        c = _cmp(a.get(name, _MISSING), b.get(name, _MISSING))
        if c:
            return c
    return 0


def K(i):
    return lambda keys, aggs: keys[i]


def A(text):
    return lambda keys, aggs: aggs[text]


def RND(text, prec):
    return lambda keys, aggs: round_float(aggs[text], prec)


PL = "(`catalog`.`pricing`).`list`"
PLIST = "((`catalog`.`pricing`).`list`)"
UP = "(`product`.`unitPrice`)"
TID = "((`product`.`test_id`) = \"agg_func\")"

CASES = [
    Case("fs_groupby_count", "test/filestore/json/default/cases/case_group_by_having.json:3-15",
         "filestore/case_group_by_having", 0, "filestore/catalog", "catalog",
         None, ["(`catalog`.`type`)"], ["count(*)"],
         [("type", K(0)), ("count", A("count(*)"))], order=["type"],
         tail=dict(terms=[("(`catalog`.`type`)", None), ("count(*)", "count")], order=[("(`catalog`.`type`)", False)])),
    Case("fs_all_aggs_nogroup", "test/filestore/json/default/cases/case_group_by_having.json:17-29",
         "filestore/case_group_by_having", 1, "filestore/catalog", "catalog",
         None, [], ["min(%s)" % PLIST, "max(%s)" % PLIST, "avg(%s)" % PLIST, "sum(%s)" % PLIST, "count(%s)" % PLIST],
         [("min", A("min(%s)" % PLIST)), ("max", A("max(%s)" % PLIST)), ("avg", A("avg(%s)" % PLIST)),
          ("sum", A("sum(%s)" % PLIST)), ("count", A("count(%s)" % PLIST))],
         tail=dict(terms=[("min(%s)" % PLIST, "min"), ("max(%s)" % PLIST, "max"), ("avg(%s)" % PLIST, "avg"), ("sum(%s)" % PLIST, "sum"), ("count(%s)" % PLIST, "count")])),
    Case("fs_all_aggs_group", "test/filestore/json/default/cases/case_group_by_having.json:31-52",
         "filestore/case_group_by_having", 2, "filestore/catalog", "catalog",
         None, ["(`catalog`.`type`)"],
         ["min(%s)" % PLIST, "max(%s)" % PLIST, "avg(%s)" % PLIST, "sum(%s)" % PLIST, "count(%s)" % PLIST],
         [("type", K(0)), ("min", A("min(%s)" % PLIST)), ("max", A("max(%s)" % PLIST)), ("avg", A("avg(%s)" % PLIST)),
          ("sum", A("sum(%s)" % PLIST)), ("count", A("count(%s)" % PLIST))], order=["type"],
         tail=dict(terms=[("(`catalog`.`type`)", None)] + [("min(%s)" % PLIST, "min"), ("max(%s)" % PLIST, "max"), ("avg(%s)" % PLIST, "avg"), ("sum(%s)" % PLIST, "sum"), ("count(%s)" % PLIST, "count")], order=[("(`catalog`.`type`)", False)])),
    Case("fs_all_aggs_having", "test/filestore/json/default/cases/case_group_by_having.json:54-67",
         "filestore/case_group_by_having", 3, "filestore/catalog", "catalog",
         None, ["(`catalog`.`type`)"],
         ["min(%s)" % PLIST, "max(%s)" % PLIST, "avg(%s)" % PLIST, "sum(%s)" % PLIST, "count(%s)" % PLIST],
         [("type", K(0)), ("min", A("min(%s)" % PLIST)), ("max", A("max(%s)" % PLIST)), ("avg", A("avg(%s)" % PLIST)),
          ("sum", A("sum(%s)" % PLIST)), ("count", A("count(%s)" % PLIST))],
         having=lambda k, a: a["count(%s)" % PLIST] > 1, order=["type"],
         tail=dict(having="(1 < count(%s))" % PLIST, terms=[("(`catalog`.`type`)", None)] + [("min(%s)" % PLIST, "min"), ("max(%s)" % PLIST, "max"), ("avg(%s)" % PLIST, "avg"), ("sum(%s)" % PLIST, "sum"), ("count(%s)" % PLIST, "count")], order=[("(`catalog`.`type`)", False)])),
    # case :161-180 has an ANY ... SATISFIES term (ineligible); all 3 catalog docs satisfy it, so the
    # eligible remainder of the predicate is checked instead.
    Case("fs_float_filter_sum", "test/filestore/json/default/cases/case_group_by_having.json:161-180",
         "filestore/case_group_by_having", 8, "filestore/catalog", "catalog",
         "(0.5 < ((`catalog`.`dimensions`).`height`))", ["(`catalog`.`title`)"],
         ["sum(((`catalog`.`dimensions`).`length`))", "sum(((`catalog`.`dimensions`).`width`))"],
         [("title", K(0)), ("$1", A("sum(((`catalog`.`dimensions`).`length`))")),
          ("$2", A("sum(((`catalog`.`dimensions`).`width`))"))],
         having=lambda k, a: a["sum(((`catalog`.`dimensions`).`width`))"] > 1 and a["sum(((`catalog`.`dimensions`).`length`))"] > 1,
         order=["title"],
         tail=dict(having="((1 < %s) and (1 < %s))" % ("sum(((`catalog`.`dimensions`).`width`))", "sum(((`catalog`.`dimensions`).`length`))"), terms=[("(`catalog`.`title`)", None), ("sum(((`catalog`.`dimensions`).`length`))", None), ("sum(((`catalog`.`dimensions`).`width`))", None)], order=[("(`catalog`.`title`)", False)])),
    Case("fs_two_keys", "test/filestore/json/default/cases/case_group_by_having.json:182-208",
         "filestore/case_group_by_having", 9, "filestore/user_profile", "user_profile",
         None, ["((`user_profile`.`personal_details`).`state`)",
                "(((`user_profile`.`profile_details`).`loyalty`).`membership_type`)"], ["count(*)"],
         [("state", K(0)), ("membership_type", K(1)), ("gold_members", A("count(*)"))],
         having=lambda k, a: k[1] == "Gold", order=["state"],
         tail=dict(having="(%s = \"Gold\")" % "(((`user_profile`.`profile_details`).`loyalty`).`membership_type`)", terms=[("((`user_profile`.`personal_details`).`state`)", None), ("(((`user_profile`.`profile_details`).`loyalty`).`membership_type`)", None), ("count(*)", "gold_members")], order=[("((`user_profile`.`personal_details`).`state`)", False)])),
    Case("fs_theme", "test/filestore/json/default/cases/case_group_by_having.json:210-237",
         "filestore/case_group_by_having", 10, "filestore/user_profile", "user_profile",
         None, ["(((`user_profile`.`profile_details`).`prefs`).`ui_theme`)"], ["count(*)"],
         [("ui_theme", K(0)), ("theme_usage", A("count(*)"))], order=["ui_theme"],
         tail=dict(terms=[("(((`user_profile`.`profile_details`).`prefs`).`ui_theme`)", None), ("count(*)", "theme_usage")], order=[("(((`user_profile`.`profile_details`).`prefs`).`ui_theme`)", False)])),
    Case("fs_count_distinct", "test/filestore/json/default/cases/case_group_by_having.json:239-252",
         "filestore/case_group_by_having", 11, "filestore/jobs", "jobs",
         None, ["(`jobs`.`join_yr`)"], ["count(distinct (`jobs`.`job_title`))"],
         [("distinct_title_count", A("count(distinct (`jobs`.`job_title`))")), ("join_yr", K(0))], order=["join_yr"],
         tail=dict(terms=[("count(distinct (`jobs`.`job_title`))", "distinct_title_count"), ("(`jobs`.`join_yr`)", None)], order=[("(`jobs`.`join_yr`)", False)])),
    Case("fs_count_distinct_and_count", "test/filestore/json/default/cases/case_group_by_having.json:277-292",
         "filestore/case_group_by_having", 13, "filestore/jobs", "jobs",
         None, ["(`jobs`.`join_yr`)"], ["count(distinct (`jobs`.`job_title`))", "count((`jobs`.`job_title`))"],
         [("distinct_title_count", A("count(distinct (`jobs`.`job_title`))")),
          ("title_count", A("count((`jobs`.`job_title`))")), ("join_yr", K(0))], order=["join_yr"],
         tail=dict(terms=[("count(distinct (`jobs`.`job_title`))", "distinct_title_count"), ("count((`jobs`.`job_title`))", "title_count"), ("(`jobs`.`join_yr`)", None)], order=[("(`jobs`.`join_yr`)", False)])),
    Case("fs_theme_order2", "test/filestore/json/default/cases/case_group_by_having.json:328-356",
         "filestore/case_group_by_having", 15, "filestore/user_profile", "user_profile",
         None, ["(((`user_profile`.`profile_details`).`prefs`).`ui_theme`)"], ["count(*)"],
         [("ui_theme", K(0)), ("theme_usage", A("count(*)"))], order=["theme_usage", "ui_theme"],
         tail=dict(terms=[("(((`user_profile`.`profile_details`).`prefs`).`ui_theme`)", None), ("count(*)", "theme_usage")], order=[("`theme_usage`", False), ("(((`user_profile`.`profile_details`).`prefs`).`ui_theme`)", False)])),
    # multistore aggregate goldens == BASELINE.json config 1 known answers
    Case("ms_product_all", "test/multistore/test_cases/aggregate_functions/case_group_by_having.json:24-36",
         "multistore/aggregate_functions/case_group_by_having", 1, "multistore/aggregate_functions/product", "product",
         TID, [], ["min(%s)" % UP, "max(%s)" % UP, "avg(%s)" % UP, "sum(%s)" % UP, "count(%s)" % UP],
         [("min", A("min(%s)" % UP)), ("max", A("max(%s)" % UP)), ("avg", RND("avg(%s)" % UP, 5)),
          ("sum", RND("sum(%s)" % UP, 5)), ("count", A("count(%s)" % UP))],
         tail=dict(terms=[("min(%s)" % UP, "min"), ("max(%s)" % UP, "max"), ("round(avg(%s), 5)" % UP, "avg"), ("round(sum(%s), 5)" % UP, "sum"), ("count(%s)" % UP, "count")], order=[("`min`", False)])),
    Case("ms_product_by_color", "test/multistore/test_cases/aggregate_functions/case_group_by_having.json:38-82",
         "multistore/aggregate_functions/case_group_by_having", 2, "multistore/aggregate_functions/product", "product",
         TID, ["(`product`.`color`)"], ["min(%s)" % UP, "max(%s)" % UP, "avg(%s)" % UP, "sum(%s)" % UP, "count(%s)" % UP],
         [("product_color", K(0)), ("min", A("min(%s)" % UP)), ("max", A("max(%s)" % UP)), ("avg", RND("avg(%s)" % UP, 5)),
          ("sum", RND("sum(%s)" % UP, 5)), ("count", A("count(%s)" % UP))], order=["min", "avg"], limit=5,
         tail=dict(terms=[("(`product`.`color`)", "product_color"), ("min(%s)" % UP, "min"), ("max(%s)" % UP, "max"), ("round(avg(%s), 5)" % UP, "avg"), ("round(sum(%s), 5)" % UP, "sum"), ("count(%s)" % UP, "count")], order=[("`min`", False), ("`avg`", False)], limit=5)),
    Case("ms_product_by_color_having", "test/multistore/test_cases/aggregate_functions/case_group_by_having.json:84-122",
         "multistore/aggregate_functions/case_group_by_having", 3, "multistore/aggregate_functions/product", "product",
         TID, ["(`product`.`color`)"], ["min(%s)" % UP, "max(%s)" % UP, "avg(%s)" % UP, "sum(%s)" % UP, "count(%s)" % UP],
         [("product_colori", K(0)), ("min", A("min(%s)" % UP)), ("max", A("max(%s)" % UP)), ("avg", RND("avg(%s)" % UP, 5)),
          ("sum", RND("sum(%s)" % UP, 5)), ("count", A("count(%s)" % UP))],
         having=lambda k, a: a["count(%s)" % UP] > 34, order=["min", "avg"],
         tail=dict(having="(34 < count(%s))" % UP, terms=[("(`product`.`color`)", "product_colori"), ("min(%s)" % UP, "min"), ("max(%s)" % UP, "max"), ("round(avg(%s), 5)" % UP, "avg"), ("round(sum(%s), 5)" % UP, "sum"), ("count(%s)" % UP, "count")], order=[("`min`", False), ("`avg`", False)])),
    Case("ms_custid_count", "test/multistore/test_cases/aggregate_functions/case_group_by_having.json:2-22",
         "multistore/aggregate_functions/case_group_by_having", 0, "multistore/aggregate_functions/orders", "orders",
         "((`orders`.`test_id`) = \"agg_func\")", ["(`orders`.`custId`)"], ["count(*)"],
         [("custId", K(0)), ("c", A("count(*)"))], order=["c", "custId"],
         tail=dict(terms=[("(`orders`.`custId`)", None), ("count(*)", "c")], order=[("`c`", False), ("(`orders`.`custId`)", False)])),
    Case("ms_totcolors", "test/multistore/test_cases/aggregate_functions/case_distinct.json:115-124",
         "multistore/aggregate_functions/case_distinct", 3, "multistore/aggregate_functions/product", "product",
         TID, [], ["count(distinct (`product`.`color`))", "count((`product`.`test_id`))"],
         # COUNT(product.categories) counts an ARRAY column (ineligible on the GPU path); every product
         # has categories, so COUNT(test_id) is the eligible stand-in with the same golden value 900.
         [("totcolors", A("count(distinct (`product`.`color`))")), ("totcategories", A("count((`product`.`test_id`))"))]),
    Case("ms_countn", "test/multistore/test_cases/aggregate_functions/case_distinct.json:231-240",
         "multistore/aggregate_functions/case_distinct", 8, "multistore/aggregate_functions/orders", "orders",
         "((`orders`.`test_id`) = \"cntn_agg_func\")", [],
         ["countn((`orders`.`cntn`))", "countn(distinct (`orders`.`cntn`))", "count((`orders`.`cntn`))",
          "count(distinct (`orders`.`cntn`))"],
         [("cntn", A("countn((`orders`.`cntn`))")), ("dcntn", A("countn(distinct (`orders`.`cntn`))")),
          ("cnt", A("count((`orders`.`cntn`))")), ("dcnt", A("count(distinct (`orders`.`cntn`))"))],
         tail=dict(terms=[("countn((`orders`.`cntn`))", "cntn"), ("countn(distinct (`orders`.`cntn`))", "dcntn"), ("count((`orders`.`cntn`))", "cnt"), ("count(distinct (`orders`.`cntn`))", "dcnt")])),
    Case("ms_bigint_sum", "test/multistore/test_cases/integers/case_select.json:33-40",
         "multistore/integers/case_select", 4, "multistore/integers/orders", "orders",
         "(((`orders`.`test_id`) = \"select_big_int\") and ((`orders`.`type`) = \"aggr\"))", ["(`orders`.`type`)"],
         ["sum((`orders`.`num`))"],
         [("total", A("sum((`orders`.`num`))")), ("type", K(0))],
         tail=dict(terms=[("sum((`orders`.`num`))", "total"), ("(`orders`.`type`)", None)])),
    Case("ms_bigint_count", "test/multistore/test_cases/integers/case_select.json:42-49",
         "multistore/integers/case_select", 5, "multistore/integers/orders", "orders",
         "(((`orders`.`test_id`) = \"select_big_int\") and (90 < (`orders`.`num`)))", [], ["count(1)"],
         [("total", A("count(1)"))],
         tail=dict(terms=[("count(1)", "total")])),
]


def _where_case(cid, ref, casefile, index, keyspace, alias, where):
    """WHERE-only goldens (no GROUP BY): the Filter is checked through COUNT(*) == number of golden rows."""
    c = Case(cid, ref, casefile, index, keyspace, alias, where, [], ["count(*)"], [("n", A("count(*)"))])
    c.count_only = True
    return c


FW = "test/filestore/json/default/cases/case_where.json"
WHERE_CASES = [
    _where_case("w_not_missing", FW + ":4-14", "filestore/case_where", 0, "filestore/tags", "tags",
                "((`tags`.`banned-on`) is not missing)"),
    _where_case("w_not_null", FW + ":16-26", "filestore/case_where", 1, "filestore/tags", "tags",
                "((`tags`.`banned-on`) is not null)"),
    _where_case("w_is_null", FW + ":28-36", "filestore/case_where", 2, "filestore/tags", "tags",
                "((`tags`.`banned-on`) is null)"),
    _where_case("w_alias_eq", FW, "filestore/case_where", 8, "filestore/contacts", "contact",
                "((`contact`.`name`) = \"dave\")"),
    _where_case("w_nested_eq_int", FW, "filestore/case_where", 9, "filestore/catalog", "catalog",
                "(((`catalog`.`pricing`).`list`) = 799)"),
    _where_case("w_not_valued", FW + ":157-170", "filestore/case_where", 14, "filestore/orders", "orders",
                "((`orders`.`shipped-on`) is not valued)"),
    _where_case("w_valued", FW, "filestore/case_where", 15, "filestore/orders", "orders",
                "((`orders`.`shipped-on`) is valued)"),
    _where_case("w_o_not_null", FW, "filestore/case_where", 16, "filestore/orders", "orders",
                "((`orders`.`shipped-on`) is not null)"),
    _where_case("w_o_null", FW, "filestore/case_where", 17, "filestore/orders", "orders",
                "((`orders`.`shipped-on`) is null)"),
    _where_case("w_o_not_missing", FW, "filestore/case_where", 18, "filestore/orders", "orders",
                "((`orders`.`shipped-on`) is not missing)"),
    _where_case("w_o_missing", FW + ":211-221", "filestore/case_where", 19, "filestore/orders", "orders",
                "((`orders`.`shipped-on`) is missing)"),
    _where_case("w_ne", FW + ":223-242", "filestore/case_where", 20, "filestore/contacts", "contacts",
                "(not ((`contacts`.`name`) = \"dave\"))"),
    _where_case("w_le", FW + ":244-254", "filestore/case_where", 21, "filestore/game", "game",
                "((`game`.`score`) <= 8)"),
    _where_case("w_ge", FW + ":256-268", "filestore/case_where", 22, "filestore/game", "game",
                "(10 <= (`game`.`score`))"),
    _where_case("w_or", FW + ":270-282", "filestore/case_where", 23, "filestore/contacts", "contacts",
                "(((`contacts`.`name`) = \"dave\") or ((`contacts`.`name`) = \"earl\"))"),
]
MW = "test/multistore/test_cases/where_functions/case_where.json"
_WT = "((`orders`.`test_id`) = \"where_func\")"
WHERE_CASES += [
    _where_case("mw_is_null", MW, "multistore/where_functions/case_where", 2, "multistore/where_functions/product",
                "product", "(((`product`.`unitPrice`) is null) and ((`product`.`test_id`) = \"where_func\"))"),
    _where_case("mw_valued", MW, "multistore/where_functions/case_where", 14, "multistore/where_functions/orders",
                "orders", "(((`orders`.`shipped-on`) is valued) and %s)" % _WT),
    _where_case("mw_not_valued", MW, "multistore/where_functions/case_where", 16, "multistore/where_functions/orders",
                "orders", "(((`orders`.`shipped-on`) is not valued) and %s)" % _WT),
    _where_case("mw_ne", MW, "multistore/where_functions/case_where", 20, "multistore/where_functions/orders",
                "orders", "((not ((`orders`.`id`) = \"1234\")) and %s)" % _WT),
    _where_case("mw_le_str", MW, "multistore/where_functions/case_where", 21, "multistore/where_functions/orders",
                "orders", "(((`orders`.`id`) <= \"1234\") and %s)" % _WT),
    _where_case("mw_gt_str", MW, "multistore/where_functions/case_where", 22, "multistore/where_functions/orders",
                "orders", "((\"1234\" < (`orders`.`id`)) and %s)" % _WT),
    _where_case("mw_or", MW, "multistore/where_functions/case_where", 23, "multistore/where_functions/orders",
                "orders", "((((`orders`.`id`) = \"1200\") or ((`orders`.`id`) = \"1236\")) and %s)" % _WT),
]


def normalise(v):
    """python value as read from a golden JSON file / produced by the engines, made comparable
    (JSON has one number type: 566 == 566.0)."""
    if isinstance(v, bool) or v is None:
        return v
    if isinstance(v, float) and v == int(v) and abs(v) < 2 ** 63:
        return int(v)
    if isinstance(v, list):
        return [normalise(x) for x in v]
    if isinstance(v, dict):
        return {k: normalise(x) for k, x in v.items()}
    return v
