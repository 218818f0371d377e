"""CPU-side checks of libn1gpu.so (no GPU needed): the C ABI loads and exports every declared symbol,
the shredder matches the oracle's document model, every query of the matrix compiles to an sm_100a
cubin through NVRTC, and the error contract (parse / ineligible / no-device) holds."""
import json
import os

import numpy as np
import pytest

import query_b200 as q
from gen_n1 import QUERIES, F, make_docs
from golden_plans import CASES, WHERE_CASES, keyspaces
from oracle import n1ql_oracle as O
from util_n1 import make_table, oracle_rows, paths_of


def test_library_exports_every_declared_symbol():
    L = q.lib()
    names = q.declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), n
    assert b"sm_100a" in L.n1gpu_version()


def _shred_check(docs, paths):
    t = q.Table([list(p) for p in paths])
    t.append_json(docs, threads=3)
    parsed = [O.parse_document(d) for d in docs]
    for p in paths:
        pay, tags = t.peek(list(p))
        d = t.dictionary(list(p))
        for r, doc in enumerate(parsed):
            v = doc
            for name in p:
                v = O.field(v, name)
            ty = O.vtype(v)
            if ty == O.T_MISSING:
                assert tags[r] == 0, (p, r)
            elif ty == O.T_NULL:
                assert tags[r] == 1
            elif ty == O.T_BOOLEAN:
                assert tags[r] == (3 if v else 2)
            elif ty == O.T_NUMBER and isinstance(v, int):
                assert tags[r] == 4 and int(pay[r]) == v, (p, r, v, int(pay[r]))
            elif ty == O.T_NUMBER:
                assert tags[r] == 5 and O.float_bits(v) == int(pay[r])
            elif ty == O.T_STRING:
                assert tags[r] == 6 and d[int(pay[r])] == v.encode("utf-8")
            else:
                assert tags[r] == 7
        assert d == sorted(d), "dictionary must be bytewise sorted"


def test_shredder_matches_oracle_on_synthetic_docs():
    docs = make_docs(3000, seed=11)
    _shred_check(docs, [("i",), ("n",), ("f",), ("s",), ("b",), ("m",), ("nest", "k"), ("nest", "deep", "x"), ("nest",), ("absent",)])


@pytest.mark.parametrize("name", ["filestore/mixed", "filestore/catalog", "filestore/non-json", "filestore/orders",
                                  "multistore/integers/orders", "filestore/user_profile"])
def test_shredder_matches_oracle_on_reference_fixtures(name):
    docs = [t for _k, t in keyspaces()[name]]
    paths = set()
    for d in docs:
        v = O.parse_document(d)
        if isinstance(v, dict):
            for k, x in v.items():
                paths.add((k,))
                if isinstance(x, dict):
                    for k2 in x:
                        paths.add((k, k2))
    _shred_check(docs, sorted(paths)[:16])


def test_shredder_edge_documents():
    docs = ['{"a": 1, "a": 2}',            # duplicate name: first wins (FirstFind)
            '  \n\t{"a": 1.0}',            # leading whitespace, integral float -> int
            '"hello"', '[1,2]', '17',      # non-object documents: all MISSING
            '{"a": 1} trailing',           # invalid JSON -> BINARY -> MISSING
            '{"a": 1e2}', '{"a": -0.0}', '{"a": 9223372036854775807}', '{"a": 9223372036854775808}',
            '{"a": "\\u00e9\\ud83d\\ude00"}', '{"\\u0061": 5}', '{"a": {"b": 1}}', '{"a": tru}', '']
    t = q.Table(["a"])
    t.append_json(docs, threads=1)
    pay, tags = t.peek("a")
    d = t.dictionary("a")
    assert list(tags[:2]) == [4, 4] and list(pay[:2]) == [1, 1]
    assert list(tags[2:6]) == [0, 0, 0, 0]
    assert tags[6] == 4 and pay[6] == 100
    assert tags[7] == 4 and pay[7] == 0
    assert tags[8] == 4 and pay[8] == 9223372036854775807
    assert tags[9] == 5 and O.float_bits(9223372036854775808.0) == int(pay[9])
    assert tags[10] == 6 and d[int(pay[10])] == "é\U0001F600".encode("utf-8")
    assert tags[11] == 4 and pay[11] == 5
    assert tags[12] == 7
    assert tags[13] == 0 and tags[14] == 0
    # and the oracle's document model agrees
    for r, doc in enumerate(docs):
        v = O.field(O.parse_document(doc), "a")
        assert (O.vtype(v) == O.T_MISSING) == (tags[r] == 0), (r, doc)


def test_numbers_nobody_reads_are_only_checked_for_syntax():
    """json.Validate (value/parsed.go:38-67) looks at syntax: an unreferenced 1e999 keeps its document valid, a malformed
    number makes every field of the document MISSING.  (The device shredder treats skipped numbers the same way.)"""
    docs = ['{"a": 1, "q": 1e999}', '{"a": 2, "q": -1E+400, "z": [1e999, {"y": 2e-999}]}', '{"a": 3, "q": 1e}', '{"a": 4, "q": 01}',
            '{"a": 5, "q": 1.}', '{"a": 6, "q": -}', '{"a": 7, "q": 1.5e-3}', '{"a": 8, "q": .5}', '{"a": 9, "q": +1}']
    t = q.Table(["a"])
    t.append_json(docs)
    pay, tags = t.peek(0)
    assert list(tags) == [4, 4, 0, 0, 0, 0, 4, 0, 0] and [int(pay[i]) for i in (0, 1, 6)] == [1, 2, 7]
    for r, doc in enumerate(docs):  # and the oracle's document model agrees on which documents are valid
        assert (O.vtype(O.field(O.parse_document(doc), "a")) == O.T_MISSING) == (tags[r] == 0), doc


@pytest.mark.parametrize("name,where,keys,aggs", QUERIES, ids=[x[0] for x in QUERIES])
def test_every_matrix_query_compiles_for_sm100a_and_oracle_runs(name, where, keys, aggs):
    docs = make_docs(400, seed=5)
    t = make_table(docs, where, keys, aggs)
    t.seal()
    qq = q.Query(t, "d", where, keys, aggs)
    src = qq.kernel_source
    assert "nq_scan" in src and "__ballot_sync" in src
    assert qq.info["words"] >= 1
    assert oracle_rows(docs, "d", where, keys, aggs) is not None


def test_row_counters_share_a_table_word_only_under_a_row_bound():
    """Behind the front cache two row counters are packed into one 64-bit table word (32-bit fields) when the declared
    row count of the whole keyspace proves that neither field can overflow; 2^32 rows or more keep one word each."""
    import numpy as np
    n = 5000
    rng = np.random.default_rng(3)
    words = ["w%05d" % i for i in range(3000)]
    vt = np.full(n, 4, dtype=np.uint8)
    vt[::7] = 1
    aggs = ["count(*)", "count((`d`.`v`))", "sum((`d`.`v`))", "min((`d`.`v`))", "max((`d`.`v`))"]
    seen = {}
    for rows in (None, n, 1 << 33):
        t = q.Table(["k", "v"])
        t.set_column("k", rng.integers(0, len(words), n).astype(np.uint32), dictionary=words)
        t.set_column("v", rng.integers(-5, 1000, n, dtype=np.int64), tags=vt)
        if rows is not None:
            t.set_global_rows(rows)
        t.seal()
        qq = q.Query(t, "d", None, ["(`d`.`k`)"], aggs)
        assert qq.info["mode"] == "hbm-direct"
        seen[rows] = (qq.info["words"], "pk0" in qq.kernel_source)
    # logical words: rows, count(v), sum(v), min, max (the sign mix of the sum is read off min / max); two counters share a word
    assert seen[None] == (4, True) and seen[n] == (4, True)
    assert seen[1 << 33][1] is False and seen[1 << 33][0] >= 5


@pytest.mark.parametrize("case", CASES + WHERE_CASES, ids=lambda c: c.id)
def test_every_golden_plan_compiles(case):
    docs = [t for _k, t in case.docs()]
    t = make_table(docs, case.where, case.keys, case.aggs)
    t.seal()
    qq = q.Query(t, case.alias, case.where, case.keys, case.aggs)
    assert qq.info["mode"] in ("ungrouped", "dense-shared-memory", "hbm-direct", "hbm-hash-64", "hbm-hash-128")


def _expect(code, fn):
    with pytest.raises(q.N1GpuError) as ei:
        fn()
    assert ei.value.code == code, ei.value
    return ei.value


def test_error_contract():
    docs = ['{"a": 1, "arr": [1,2], "s": "x", "s2": "y"}', '{"a": 2, "arr": [3], "s": "z", "s2": "w"}']
    t = q.Table(["a", "arr", "s", "s2"])
    t.append_json(docs)
    t.seal()
    E = q._lib
    _expect(E.E_PARSE, lambda: q.Query(t, "d", "((`d`.`a`) <", [], ["count(*)"]))
    # outside the subset -> INELIGIBLE (the caller keeps its Go operators); never a silent fallback
    for where in ["((`d`.`s`) like \"x%\")", "(length((`d`.`s`)) < 3)", "any `x` in (`d`.`arr`) satisfies (`x` = 1) end",
                  "((`d`.`arr`) = [1, 2])", "((`d`.`s`) < (`d`.`s2`))", "(`d` is valued)", "((`e`.`a`) = 1)"]:
        e = _expect(E.E_INELIGIBLE, lambda: q.Query(t, "d", where, [], ["count(*)"]))
        assert isinstance(e, q.Ineligible)
    _expect(E.E_INELIGIBLE, lambda: q.Query(t, "d", None, ["(`d`.`arr`)"], ["count(*)"]))
    _expect(E.E_INELIGIBLE, lambda: q.Query(t, "d", None, [], ["array_agg((`d`.`a`))"]))
    _expect(E.E_INELIGIBLE, lambda: q.Query(t, "d", None, [], ["min(distinct (`d`.`a`))"]))
    _expect(E.E_INVALID, lambda: q.Query(t, "d", "((`d`.`unknown`) = 1)", [], ["count(*)"]))
    t2 = q.Table(["a"])
    t2.append_json(docs)
    _expect(E.E_INVALID, lambda: q.Query(t2, "d", None, [], ["count(*)"]))  # not sealed


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    t = q.Table(["a"])
    t.append_json(['{"a": 1}'])
    t.seal()
    qq = q.Query(t, "d", None, [], ["count(*)"])
    e = _expect(q._lib.E_CUDA, qq.execute)
    assert "no CPU fallback" in str(e)
    _expect(q._lib.E_CUDA, lambda: q.init(0))


def test_plan_build_substitution_rules(tmp_path):
    from util_n1 import write_keyspace
    write_keyspace(str(tmp_path), "default", "game", keyspaces()["filestore/game"])
    from plans_n1 import explain_plan
    plan = explain_plan("default", "game", None, None, [], ["count(*)", "min((`game`.`score`))"])
    import torch
    if torch.cuda.is_available():
        op = q.Operator(plan, str(tmp_path))
        assert op.rest_index == 5
    # ineligible shapes are refused with INELIGIBLE whatever the device situation
    bad = json.loads(json.dumps(plan))
    bad["~children"][0]["~children"][0]["limit"] = "10"
    _expect(q._lib.E_INELIGIBLE, lambda: q.Operator(bad, str(tmp_path)))
    nogroup = {"#operator": "Sequence", "~children": [plan["~children"][0]["~children"][0], plan["~children"][0]["~children"][1]]}
    _expect(q._lib.E_INELIGIBLE, lambda: q.Operator(nogroup, str(tmp_path)))
    _expect(q._lib.E_PARSE, lambda: q.Operator("{not json", str(tmp_path)))


@pytest.mark.parametrize("ks,where,keys,aggs,ref", [
    ("catalog", None, [], ["array_agg((`catalog`.`asin`))"], "case_group_by_having.json:69-81 ARRAY_AGG"),
    ("orders", None, ["((`orders`.`orderlines`)[1])"], ["count(*)"], "case_group_by_having.json:83-110 array element key"),
    ("orders", None, ["(`orders`.`orderlines`)"], ["count(*)"], "case_group_by_having.json:112-146 array-valued key"),
    ("catalog", "any `director` in ((`catalog`.`details`).`director`) satisfies `director` end", ["((`catalog`.`details`).`director`)"],
     ["count(*)"], "case_group_by_having.json:148-159 ANY ... SATISFIES"),
    ("jobs", None, ["(`jobs`.`join_yr`)"], ["array_agg(distinct (`jobs`.`job_title`))"], "case_group_by_having.json:254-275 ARRAY_AGG DISTINCT"),
], ids=["array_agg", "element_key", "array_key", "any_satisfies", "array_agg_distinct"])
def test_golden_statements_outside_the_subset_are_not_substituted(ks, where, keys, aggs, ref, tmp_path):
    """SURVEY.md 8c lists these goldens as "must stay on the Go path": the plan builder answers INELIGIBLE (the caller
    keeps its operators) - at parse time for constructs outside the subset, at bind time for a column holding arrays."""
    from plans_n1 import explain_plan
    from util_n1 import write_keyspace
    write_keyspace(str(tmp_path), "default", ks, keyspaces()["filestore/" + ks])
    plan = explain_plan("default", ks, None, where, keys, aggs)
    for tail in (False, True):
        e = _expect(q._lib.E_INELIGIBLE, lambda: q.Operator(plan, str(tmp_path), tail=tail))
        assert isinstance(e, q.Ineligible), ref


def test_on_disk_kernel_cache(tmp_path):
    """N1GPU_KERNEL_CACHE_DIR: a query shape compiled once (NVRTC + ptxas, 0.25-0.45 s) is an ELF cubin on disk that a
    restarted process loads instead of compiling; the key covers source, device library, NVRTC version and target."""
    import subprocess
    import sys
    prog = (
        "import sys, time; sys.path.insert(0, %r)\n"
        "import numpy as np, query_b200 as q\n"
        "t = q.Table(['n']); t.set_column('n', np.arange(1000, dtype=np.int64)); t.seal()\n"
        "t0 = time.time(); qq = q.Query(t, 'd', '((`d`.`n`) between 31 and 415)', [], ['count(*)', 'sum((`d`.`n`))'])\n"
        "print(time.time() - t0)\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, N1GPU_KERNEL_CACHE_DIR=str(tmp_path))
    cold = float(subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, check=True).stdout.split()[-1])
    files = [f for f in os.listdir(str(tmp_path)) if f.endswith(".cubin")]
    assert len(files) == 1 and open(os.path.join(str(tmp_path), files[0]), "rb").read(4) == b"\x7fELF"
    warm = float(subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, check=True).stdout.split()[-1])
    assert len(os.listdir(str(tmp_path))) == 1 and warm < cold / 3, (cold, warm)
    # a damaged file is ignored and replaced - also one that is still an ELF but fails the entry's checksum
    good = open(os.path.join(str(tmp_path), files[0]), "rb").read()
    assert good[-16:-8] == b"N1CUBIN1"
    for bad in (b"garbage", good[:len(good) // 2] + bytes([good[len(good) // 2] ^ 1]) + good[len(good) // 2 + 1:], good[:-16]):
        open(os.path.join(str(tmp_path), files[0]), "wb").write(bad)
        subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, check=True)
        assert open(os.path.join(str(tmp_path), files[0]), "rb").read() == good


def test_random_expression_trees_compile_for_sm100a():
    """The generator of tools/where_fuzz.py (the device differential fuzz, run under -m gpu): every random Filter /
    operand tree it produces is either outside the subset (INELIGIBLE) or turns into a kernel NVRTC accepts."""
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import where_fuzz as wf
    rng = random.Random(7)
    docs = make_docs(300, seed=77)
    compiled = 0
    for _ in range(24):
        where = str(O.parse(wf.cond(rng, 2)))
        aggs = sorted({str(O.parse(a)) for a in ["count(*)", "count(%s)" % wf.arith(rng, 2, wf.ANY), "min(%s)" % wf.arith(rng, 2, wf.NUM + ["b"]),
                                                 "max(%s)" % wf.arith(rng, 2, wf.NUM), "sum(%s)" % wf.arith(rng, 1, ["i", "p", "g", "f"])]})
        keys = [wf.arith(rng, 1, ["g", "t", "b", "i"])] if rng.random() < 0.5 else []
        try:
            t = make_table(docs, where, keys, aggs)
            t.seal()
            q.Query(t, "d", where, keys, aggs)
            compiled += 1
        except q.Ineligible:
            pass
    assert compiled >= 12


def test_fast_count_plan_is_left_alone(tmp_path):
    """SELECT COUNT(*) FROM ks without WHERE is planned as CountScan + the group operators (golden EXPLAIN:
    test/filestore/json/default/cases/case_by_id.json:297-366) and must not be touched - there is no PrimaryScan / Fetch
    to replace."""
    from util_n1 import write_keyspace
    write_keyspace(str(tmp_path), "default", "game", keyspaces()["filestore/game"])
    grp = lambda name: {"#operator": name, "aggregates": ["count(*)"], "group_keys": []}
    plan = {"#operator": "Sequence", "~children": [
        {"#operator": "CountScan", "keyspace": "game", "namespace": "default"},
        {"#operator": "Parallel", "maxParallelism": 1, "~child": {"#operator": "Sequence", "~children": [grp("InitialGroup")]}},
        grp("IntermediateGroup"), grp("FinalGroup"),
        {"#operator": "Parallel", "maxParallelism": 1, "~child": {"#operator": "Sequence", "~children": [
            {"#operator": "InitialProject", "result_terms": [{"as": "c", "expr": "count(*)"}]}, {"#operator": "FinalProject"}]}}]}
    for tail in (False, True):
        _expect(q._lib.E_INELIGIBLE, lambda: q.Operator(plan, str(tmp_path), tail=tail))


def test_keyspace_directory_follows_the_reference_id_rule(tmp_path):
    """datastore/file/file.go: the primary index scans ids = file name minus its LAST extension (:711-730, :757-761) in name
    order and Fetch reads <id>.json (:346); an id without such a file is silently no document (:319-322).  So a.txt next to
    a.json yields a.json twice, foo.txt / .DS_Store / editor backups alone yield nothing, b.tar.json is the id b.tar."""
    d = tmp_path / "default" / "ks"
    d.mkdir(parents=True)
    (d / "a.json").write_text('{"n": 1}')
    (d / "a.txt").write_text("not json")            # id "a" again -> a.json a second time
    (d / "foo.txt").write_text("no foo.json here")   # id "foo" -> foo.json missing -> nothing
    (d / ".DS_Store").write_text("x")                # id "" -> ".json" missing -> nothing
    (d / "b.tar.json").write_text('{"n": 10}')       # id "b.tar" -> b.tar.json
    (d / "c.json~").write_text('{"n": 100}')         # id "c" -> c.json missing -> nothing
    (d / "e.json").write_text("")                    # an empty file IS a document (not JSON: every field MISSING)
    (d / "sub").mkdir()                              # directories are not entries
    t = q.Table(["n"])
    t.load_dir(str(d))
    assert t.num_rows == 4
    pay, tags = t.peek("n")
    assert sorted(zip(tags.tolist(), pay.tolist())) == [(0, 0), (4, 1), (4, 1), (4, 10)]


def test_parameters_and_constants_are_kernel_arguments():
    """algebra/param_named.go:63, param_positional.go (Stringer text $name / $1, expression/stringer.go:611-620): a prepared
    statement's parameters take the request's values when the chain is built; constants and parameter values reach the
    kernel as arguments, so every binding - and the same statement with literals - shares ONE compiled kernel."""
    docs = make_docs(1500, seed=5)
    where = "((%s between $lo and $2) and (%s = $t))" % (F("p"), F("t"))
    keys, aggs = [F("t")], ["count(*)", "sum(%s)" % F("p"), "max(%s)" % F("s")]
    t = make_table(docs, "((%s between 1 and 2) and (%s = \"x\"))" % (F("p"), F("t")), keys, aggs)
    t.seal()
    c0, _r0 = q.jit_stats()
    a = q.Query(t, "d", where, keys, aggs, params={"lo": 10, "2": 500, "t": "t1"})
    c1, r1 = q.jit_stats()
    b = q.Query(t, "d", where, keys, aggs, params={"$lo": -7, "2": 90000, "t": "t3"})
    lit = q.Query(t, "d", "((%s between 11 and 499) and (%s = \"t2\"))" % (F("p"), F("t")), keys, aggs)
    c2, r2 = q.jit_stats()
    assert a.kernel_source == b.kernel_source == lit.kernel_source and "p.cst[" in a.kernel_source
    assert c1 - c0 <= 1 and c2 == c1 and r2 >= r1 + 2, "re-binding must not compile"
    with pytest.raises(q.N1GpuError, match="No value for named parameter"):
        q.Query(t, "d", where, keys, aggs, params={"lo": 10, "2": 5})
    with pytest.raises(q.N1GpuError, match="No value for positional parameter"):
        q.Query(t, "d", where, keys, aggs, params={"lo": 10, "t": "x"})
    with pytest.raises(q.Ineligible):  # non-scalar values stay with the caller's operators
        q.Query(t, "d", where, keys, aggs, params={"lo": [1, 2], "2": 5, "t": "x"})
    # a float where an int stood changes the class: another kernel, same statement
    f = q.Query(t, "d", where, keys, aggs, params={"lo": 10.5, "2": 500, "t": "t1"})
    assert f.kernel_source != a.kernel_source


def test_dictionaries_of_many_ranks_merge_in_parallel_cuts():
    """n1gpu_table_dict_merge: the column's own sorted dictionary and those of 7 other ranks (60 000-string vocabulary, each
    rank saw ~70 % of it) merge into the bytewise sorted union - the range is cut across threads - and the column's ranks
    are remapped; unsorted or repeated input is refused."""
    import numpy as np
    rng = np.random.default_rng(9)
    vocab = sorted({("w%06d" % i).encode() for i in range(60000)} | {b"", "é".encode(), b"z" * 40, b"w000010\x00x"})

    def some(frac):
        return [s for s, keep in zip(vocab, rng.random(len(vocab)) < frac) if keep]

    def raw(strings):
        offs = np.zeros(len(strings) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in strings], out=offs[1:])
        return np.frombuffer(b"".join(strings) or b"\0", dtype=np.uint8), offs

    own = some(0.7)
    codes = rng.integers(0, len(own), 5000).astype(np.uint32)
    t = q.Table(["k"])
    t.set_column("k", codes, dictionary=[s.decode() for s in own])
    others = [some(0.7) for _ in range(6)] + [[]]
    t.merge_dictionaries("k", [raw(o) for o in others])
    union = sorted(set(own).union(*others))
    assert t.dictionary("k") == union
    pay, tags = t.peek("k")
    assert [union[int(r)] for r in pay[:500]] == [own[int(c)] for c in codes[:500]]
    t2 = q.Table(["k"])
    t2.set_column("k", codes, dictionary=[s.decode() for s in own])
    bad = list(others[0])
    bad[30000], bad[30001] = bad[30001], bad[30000]
    with pytest.raises(q.N1GpuError):
        t2.merge_dictionaries("k", [raw(bad)])
    with pytest.raises(q.N1GpuError):
        t2.merge_dictionaries("k", [raw(others[1][:100] + others[1][99:])])


def test_plain_filter_project_is_not_substituted(tmp_path):
    """SURVEY 8 f4 decision: a Filter -> Project scan without GROUP BY / DISTINCT returns every selected
    row - nothing to aggregate, output as large as the input - so the plan builder leaves it to the caller's operators
    (the shape is the reference's EXPLAIN of `SELECT name FROM game WHERE score > 5`, plan/build_select_sub.go order)."""
    from util_n1 import write_keyspace
    write_keyspace(str(tmp_path), "default", "game", keyspaces()["filestore/game"])
    term = {"keyspace": "game", "namespace": "default"}
    plan = {"#operator": "Sequence", "~children": [{"#operator": "Sequence", "~children": [
        dict({"#operator": "PrimaryScan", "index": "#primary", "using": "default"}, **term), dict({"#operator": "Fetch"}, **term),
        {"#operator": "Parallel", "~child": {"#operator": "Sequence", "~children": [
            {"#operator": "Filter", "condition": "((`game`.`score`) > 5)"},
            {"#operator": "InitialProject", "result_terms": [{"expr": "(`game`.`name`)"}]}, {"#operator": "FinalProject"}]}}]},
        {"#operator": "Stream"}]}
    for tail in (False, True):
        e = _expect(q._lib.E_INELIGIBLE, lambda: q.Operator(plan, str(tmp_path), tail=tail))
        assert isinstance(e, q.Ineligible)


def test_kernel_shapes_follow_the_table_statistics():
    """Layout decisions that the B200 measurements settled, pinned on the generated text (no GPU needed): the skewed string
    GROUP BY of config 5 gets one 1024-thread block per SM capped at 56 registers (room for the merge / finalisation kernels
    of the neighbouring steps), a direct-mapped u32 front cache, MIN / MAX cells interleaved for one 64-bit load and a
    one-cell sum that sends its carries to the table; Q1's 6 groups x 7 words get thread-private tables in a 640-thread block."""
    import numpy as np
    rng = np.random.default_rng(5)
    n, vocab = 200000, 100000
    words = ["w%06d" % i for i in range(vocab)]
    t = q.Table(["k", "v"])
    kt = np.where(rng.integers(0, 10, n) == 0, 0, 6).astype(np.uint8)
    t.set_column("k", rng.integers(0, vocab, n).astype(np.uint32), tags=kt, dictionary=words)
    vt = np.where(rng.integers(0, 10, n) == 0, 1, 4).astype(np.uint8)
    t.set_column("v", rng.integers(-1000, 1000000, n, dtype=np.int64), tags=vt)
    t.set_global_rows(1_000_000_000)
    t.seal()
    qq = q.Query(t, "d", "((`d`.`v`) is not missing)", ["(`d`.`k`)"],
                 ["count(*)", "count((`d`.`v`))", "sum((`d`.`v`))", "min((`d`.`v`))", "max((`d`.`v`))"])
    src = qq.kernel_source
    assert qq.info["mode"] == "hbm-direct" and qq.info["block"] == 1024 and qq.info["words"] == 4
    assert "__maxnreg__(56)" in src and "cache_claim_1(" in src and "cache_add_carry(" in src and "mmp" in src and "volatile u64*" in src
    # TPC-H Q1 shape: two low-cardinality string keys, mixed INT / FLOAT money columns
    m = 60000
    t3 = q.Table(["f", "s", "qty", "price", "disc"])
    t3.set_column("f", rng.integers(0, 3, m).astype(np.uint32), dictionary=["A", "N", "R"])
    t3.set_column("s", rng.integers(0, 2, m).astype(np.uint32), dictionary=["F", "O"])
    t3.set_column("qty", rng.integers(1, 51, m, dtype=np.int64))
    cents = rng.integers(90000, 10500000, m)
    whole = cents % 100 == 0
    price = np.where(whole, cents // 100, 0).astype(np.int64)
    price[~whole] = np.array(cents[~whole] / 100.0).view(np.int64)
    t3.set_column("price", price, tags=np.where(whole, 4, 5).astype(np.uint8))
    pct = rng.integers(0, 11, m)
    disc = np.zeros(m, dtype=np.int64)
    disc[pct != 0] = np.array(pct[pct != 0] / 100.0).view(np.int64)
    t3.set_column("disc", disc, tags=np.where(pct == 0, 4, 5).astype(np.uint8))
    t3.set_global_rows(60_000_000)
    t3.seal()
    q3 = q.Query(t3, "l", None, ["(`l`.`f`)", "(`l`.`s`)"],
                 ["sum((`l`.`qty`))", "sum((`l`.`price`))", "sum(((`l`.`price`) * (1 - (`l`.`disc`))))", "avg((`l`.`disc`))", "count(*)"])
    assert q3.info["mode"] == "dense-shared-memory" and q3.info["slots"] == 6
    assert "s_priv" in q3.kernel_source and q3.info["block"] >= 512 and q3.info["block"] % 32 == 0
