"""world_size-2 gloo tests (CPU) of the host-side exchange logic of query_b200.dist: owner-bucketed all-to-all,
all-gather of small states, row-range partitioning and the dictionary / statistics agreement before seal."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _exchange(rank, world):
    from query_b200 import dist as qd
    rw = 3
    # rank r sends (r+1) records to rank 0 and (r+2) records to rank 1; record = [src, dst, seq]
    counts = [rank + 1, rank + 2]
    recs = []
    for dst, c in enumerate(counts):
        for i in range(c):
            recs += [rank, dst, i]
    got, rc = qd.exchange_by_owner(torch.tensor(recs, dtype=torch.int64), counts, rw)
    return got.tolist(), rc


def test_exchange_by_owner_routes_every_record_to_its_owner():
    out = _run(_exchange)
    for owner, (flat, rc) in enumerate(out):
        recs = [tuple(flat[i:i + 3]) for i in range(0, len(flat), 3)]
        assert rc == [0 + 1 + owner, 1 + 1 + owner]
        assert all(r[1] == owner for r in recs)
        assert sorted(recs) == sorted((src, owner, i) for src in range(2) for i in range(src + 1 + owner))


def _gather(rank, world):
    from query_b200 import dist as qd
    rw = 2
    n = 3 if rank == 0 else 1
    recs = torch.tensor([v for i in range(n) for v in (rank, i)], dtype=torch.int64)
    allrec, cs = qd.gather_all(recs, n, rw)
    return allrec.tolist(), cs, qd.row_range(10)


def test_gather_all_with_ragged_counts_and_row_ranges():
    out = _run(_gather)
    for flat, cs, rr in out:
        assert cs == [3, 1]
        assert flat == [0, 0, 0, 1, 0, 2, 1, 0]
    assert out[0][2] == (0, 5) and out[1][2] == (5, 10)


def _agree(rank, world):
    import query_b200 as q
    from query_b200 import dist as qd
    docs = ['{"s":"b","n":5}', '{"s":"a","n":-2}'] if rank == 0 else ['{"s":"c","n":40}', '{"s":"a"}', '{"n":1.5}']
    t = q.Table(["s", "n"])
    t.append_json(docs)
    qd.agree_dictionaries_and_stats(t)
    pay, tags = t.peek("s")
    d = t.dictionary("s")
    st = t.stats("n")
    t.seal()
    qq = q.Query(t, "d", "((`d`.`s`) <= \"b\")", ["(`d`.`s`)"], ["count(*)", "count((`d`.`n`))", "sum((`d`.`n`))"])
    return [x.decode() for x in d], [int(p) for p, g in zip(pay, tags) if g == 6], st.tolist(), qq.kernel_source


def test_dictionary_and_statistics_agreement_gives_identical_kernels():
    out = _run(_agree)
    assert out[0][0] == out[1][0] == ["a", "b", "c"]
    assert out[0][1] == [1, 0] and out[1][1] == [2, 0]
    assert out[0][2][:5] == out[1][2][:5] == [(1 << 0) | (1 << 4) | (1 << 5), 1, -2, 40, 1]
    assert out[0][2][7] == out[1][2][7] == 1  # MISSING / NULL rows of `n` over the whole keyspace (rank 1 holds the one)
    assert out[0][3] == out[1][3], "ranks must compile the same kernel (same packing, same constants)"


def _disagree(rank, world):
    import query_b200 as q
    from query_b200 import dist as qd
    # no agreement before seal: rank 1's dictionary and integer range differ, so the ranks compile different kernels
    docs = ['{"s":"b","n":5}', '{"s":"a","n":-2}'] if rank == 0 else ['{"s":"c","n":40000000000}', '{"s":"a"}', '{"n":1.5}']
    t = q.Table(["s", "n"])
    t.append_json(docs)
    t.seal()
    qq = q.Query(t, "d", None, ["(`d`.`s`)"], ["count(*)", "sum((`d`.`n`))"])
    try:
        qd.DistributedQuery(qq)
    except RuntimeError as e:
        return str(e)
    return "accepted"


def test_ranks_with_different_kernels_are_refused_not_hung():
    """DistributedQuery compares a digest of (mode, word ops, kernel text) across ranks with one all-reduce: ranks that would
    enter different collectives or merge differently laid out words get an error on every rank instead of a hang."""
    out = _run(_disagree)
    assert all("different kernels" in o for o in out), out


def _gathered_tail(rank, world):
    """Each rank holds the groups whose key hashes to it (what the owner-bucketed merge leaves); the tail runs over the
    gathered result on every rank."""
    import sys
    import tempfile
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import query_b200 as q
    from gen_n1 import F, make_docs
    from oracle import n1ql_oracle as O
    from plans_n1 import explain_plan
    from query_b200 import dist as qd
    from util_n1 import write_keyspace
    docs = make_docs(400, seed=71)
    keys, aggs = [F("t"), F("h")], sorted({"count(*)", "sum(%s)" % F("p"), "max(%s)" % F("s")})
    tail = dict(having="(count(*) > 1)", terms=[(F("t"), None), (F("h"), "h"), ("count(*)", "c"), ("max(%s)" % F("s"), "top")],
                order=[("`c`", True), (F("t"), False), ("`h`", False)], limit=9)
    root = tempfile.mkdtemp()
    write_keyspace(root, "default", "d", [("k%05d" % i, t) for i, t in enumerate(docs)])
    op = q.Operator(explain_plan("default", "d", "d", None, keys, aggs, tail=tail), root, tail=True)
    groups = O.run_chain([O.parse_document(d) for d in docs], "d", None, keys, aggs)
    conv = lambda v: q.MISSING if v is O.MISSING else v
    rows = [([conv(k) for k in g.keys], [conv(g.aggregates[a]) for a in aggs]) for g in groups]
    mine = [r for i, r in enumerate(rows) if i % world == rank]  # this rank's share of the groups
    merged = qd.gather_results(op, op.import_result(mine))
    got = op.run_tail(merged)
    want = O.run_tail(groups, having=tail["having"], terms=tail["terms"], order=tail["order"], limit=tail["limit"])
    return merged.num_groups == len(rows), got == want or [got, want]


def test_tail_over_groups_gathered_from_all_ranks():
    for complete, same in _run(_gathered_tail):
        assert complete and same is True, same
