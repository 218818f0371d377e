"""Plan JSON in the reference's EXPLAIN shape (golden: test/filestore/json/default/cases/case_by_id.json:369-456;
planner/build_select_sub.go:217-296, planner/build_select.go:75-110) for a
SELECT ... FROM ks [WHERE] GROUP BY ... [LETTING] [HAVING] [ORDER BY] [OFFSET] [LIMIT] statement."""


def explain_plan(namespace, keyspace, alias, where, keys, aggs, parallel=True, tail=None):
    """tail (optional): dict(letting=[(var, expr)], having=expr, terms=[(expr, as or None)], order=[(expr, desc)],
    offset=expr, limit=expr) in Stringer text - the operators the planner puts behind FinalGroup.  Without it the
    projection lists the aggregates (what every earlier test used)."""
    term = {"keyspace": keyspace, "namespace": namespace}
    if alias and alias != keyspace:
        term["as"] = alias
    sub = []
    if where:
        sub.append({"#operator": "Filter", "condition": where})
    sub.append({"#operator": "InitialGroup", "aggregates": sorted(set(aggs)), "group_keys": list(keys)})
    mid = {"#operator": "Parallel", "~child": {"#operator": "Sequence", "~children": sub}} if parallel else \
        {"#operator": "Sequence", "~children": sub}
    tail = tail or {}
    proj = []
    if tail.get("letting"):
        proj.append({"#operator": "Let", "bindings": [{"var": v, "expr": e} for v, e in tail["letting"]]})
    if tail.get("having"):
        proj.append({"#operator": "Filter", "condition": tail["having"]})
    terms = tail.get("terms") or [(a, None) for a in aggs]
    proj.append({"#operator": "InitialProject",
                 "result_terms": [dict({"expr": e}, **({"as": a} if a else {})) for e, a in terms]})
    delayed = bool(tail.get("order"))  # build_select_sub.go:225-233: the final projection waits for ORDER BY
    if not delayed:
        proj.append({"#operator": "FinalProject"})
    children = [
        dict({"#operator": "PrimaryScan", "index": "#primary", "using": "default"}, **term),
        dict({"#operator": "Fetch"}, **term),
        mid,
        {"#operator": "IntermediateGroup", "aggregates": sorted(set(aggs)), "group_keys": list(keys)},
        {"#operator": "FinalGroup", "aggregates": sorted(set(aggs)), "group_keys": list(keys)},
        {"#operator": "Parallel", "~child": {"#operator": "Sequence", "~children": proj}},
    ]
    outer = [{"#operator": "Sequence", "~children": children}]
    if tail.get("order"):
        o = {"#operator": "Order", "sort_terms": [dict({"expr": e}, **({"desc": True} if d else {})) for e, d in tail["order"]]}
        if tail.get("limit") is not None:  # plan/order.go:68-73: the pushed-down hints
            o["limit"] = str(tail["limit"])
            if tail.get("offset") is not None:
                o["offset"] = str(tail["offset"])
        outer.append(o)
    if tail.get("offset") is not None:
        outer.append({"#operator": "Offset", "expr": str(tail["offset"])})
    if tail.get("limit") is not None:
        outer.append({"#operator": "Limit", "expr": str(tail["limit"])})
    if delayed:
        outer.append({"#operator": "FinalProject"})
    outer.append({"#operator": "Stream"})
    return {"#operator": "Sequence", "~children": outer}


def distinct_plan(namespace, keyspace, alias, where, terms, order=None, limit=None):
    """EXPLAIN shape of SELECT DISTINCT <terms> FROM ks [WHERE] [ORDER BY] [LIMIT] (planner/build_select_sub.go:217-243:
    the parallel part projects and de-duplicates per stream, a serial Distinct follows; planner/build_select.go:75-110)."""
    term = {"keyspace": keyspace, "namespace": namespace}
    if alias and alias != keyspace:
        term["as"] = alias
    sub = []
    if where:
        sub.append({"#operator": "Filter", "condition": where})
    sub.append({"#operator": "InitialProject", "distinct": True,
                "result_terms": [dict({"expr": e}, **({"as": a} if a else {})) for e, a in terms]})
    sub.append({"#operator": "Distinct"})
    if not order:
        sub.append({"#operator": "FinalProject"})
    children = [
        dict({"#operator": "PrimaryScan", "index": "#primary", "using": "default"}, **term),
        dict({"#operator": "Fetch"}, **term),
        {"#operator": "Parallel", "~child": {"#operator": "Sequence", "~children": sub}},
        {"#operator": "Distinct"},
    ]
    outer = [{"#operator": "Sequence", "~children": children}]
    if order:
        outer.append({"#operator": "Order", "sort_terms": [dict({"expr": e}, **({"desc": True} if d else {})) for e, d in order]})
    if limit is not None:
        outer.append({"#operator": "Limit", "expr": str(limit)})
    if order:
        outer.append({"#operator": "FinalProject"})
    outer.append({"#operator": "Stream"})
    return {"#operator": "Sequence", "~children": outer}
