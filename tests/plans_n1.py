"""Plan JSON in the reference's EXPLAIN shape (golden: test/filestore/json/default/cases/case_by_id.json:369-456;
planner/build_select_sub.go:217-296) for a SELECT ... FROM ks [WHERE] GROUP BY ... statement."""


def explain_plan(namespace, keyspace, alias, where, keys, aggs, parallel=True):
    term = {"keyspace": keyspace, "namespace": namespace}
    if alias and alias != keyspace:
        term["as"] = alias
    sub = []
    if where:
        sub.append({"#operator": "Filter", "condition": where})
    sub.append({"#operator": "InitialGroup", "aggregates": sorted(set(aggs)), "group_keys": list(keys)})
    mid = {"#operator": "Parallel", "~child": {"#operator": "Sequence", "~children": sub}} if parallel else \
        {"#operator": "Sequence", "~children": sub}
    children = [
        dict({"#operator": "PrimaryScan", "index": "#primary", "using": "default"}, **term),
        dict({"#operator": "Fetch"}, **term),
        mid,
        {"#operator": "IntermediateGroup", "aggregates": sorted(set(aggs)), "group_keys": list(keys)},
        {"#operator": "FinalGroup", "aggregates": sorted(set(aggs)), "group_keys": list(keys)},
        {"#operator": "Parallel", "~child": {"#operator": "Sequence", "~children": [
            {"#operator": "InitialProject", "result_terms": [{"expr": a} for a in aggs]},
            {"#operator": "FinalProject"}]}},
    ]
    return {"#operator": "Sequence", "~children": [{"#operator": "Sequence", "~children": children}, {"#operator": "Stream"}]}
