"""Cross-checks the C restatement (oracle/oracle_ref.c: the CPU baseline and mid-size checker) against the
pure-Python oracle (pinned to the reference's golden vectors) on the whole query matrix and the goldens."""
import pytest

from gen_n1 import QUERIES, make_docs
from golden_plans import CASES, WHERE_CASES
from oracle import cref
from oracle import n1ql_oracle as O
from util_n1 import assert_same, oracle_rows


@pytest.mark.parametrize("name,where,keys,aggs", QUERIES, ids=[x[0] for x in QUERIES])
@pytest.mark.parametrize("threads", [1, 3])
def test_cref_matches_python_oracle(name, where, keys, aggs, threads):
    docs = make_docs(1500, seed=31)
    exp = oracle_rows(docs, "d", where, keys, aggs, streams=1)
    got = cref.rows(docs, "d", where, keys, aggs, threads=threads)
    assert_same(exp, got, "%s threads=%d" % (name, threads))


@pytest.mark.parametrize("case", CASES + WHERE_CASES, ids=lambda c: c.id)
def test_cref_matches_python_oracle_on_goldens(case):
    docs = [t for _k, t in case.docs()]
    aggs = sorted(set(case.aggs))
    assert_same(oracle_rows(docs, case.alias, case.where, case.keys, aggs), cref.rows(docs, case.alias, case.where, case.keys, aggs), case.id)


def test_generated_documents_are_valid_and_deterministic():
    import json
    for config in (2, 3, 4, 5):
        buf, offs = cref.gen_docs(config, 7, 100, 500)
        buf2, offs2 = cref.gen_docs(config, 7, 350, 250)
        raw = bytes(buf)
        docs = [raw[offs[i]:offs[i + 1]] for i in range(500)]
        for d in docs:
            assert isinstance(json.loads(d), dict)
        raw2 = bytes(buf2)
        assert docs[250:] == [raw2[offs2[i]:offs2[i + 1]] for i in range(250)]
    buf, offs = cref.gen_docs(5, 1, 0, 4000)
    docs = [O.parse_document(bytes(buf)[offs[i]:offs[i + 1]]) for i in range(4000)]
    miss = sum(1 for d in docs if "k" not in d)
    null = sum(1 for d in docs if "k" in d and d["k"] is None)
    assert 300 < miss < 500 and 300 < null < 500
