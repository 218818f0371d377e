"""persistent columnar segments and the packed NDJSON document source (query_b200/csrc/segment.cpp).
CPU part: the files, their validation and the statistics / dictionaries / generated kernels they lead to.  The device
part (marked gpu) checks query results through segment-loaded tables against the oracle."""
import json
import os
import time

import numpy as np
import pytest

import query_b200 as q
from gen_n1 import F, make_docs
from plans_n1 import explain_plan
from util_n1 import assert_same, gpu_rows, oracle_rows, write_keyspace

COLS = [["t"], ["p"], ["f"], ["s"], ["h"]]
WHERE = "(%s is not missing)" % F("p")
KEYS = [F("t")]
AGGS = ["count(*)", "sum(%s)" % F("p"), "min(%s)" % F("s"), "avg(%s)" % F("f"), "count(distinct %s)" % F("h")]


def shredded(docs):
    t = q.Table(COLS)
    t.append_json(docs)
    return t


def test_ndjson_source_equals_appended_documents(tmp_path):
    docs = make_docs(500, seed=41)
    path = str(tmp_path / "d.ndjson")
    with open(path, "w", encoding="utf-8") as f:
        for i, d in enumerate(docs):
            f.write(d + ("\r\n" if i % 7 == 0 else "\n"))
            if i % 50 == 0:
                f.write("   \n")  # blank lines are not documents
    a, b = shredded(docs), q.Table(COLS).load_ndjson(path)
    assert a.num_rows == b.num_rows == 500
    for c in range(len(COLS)):
        pa, ta = a.peek(c)
        pb, tb = b.peek(c)
        assert (ta == tb).all()
        assert a.dictionary(c) == b.dictionary(c) and (a.stats(c) == b.stats(c)).all()
    a.seal(), b.seal()
    assert q.Query(a, "d", WHERE, KEYS, AGGS).kernel_source == q.Query(b, "d", WHERE, KEYS, AGGS).kernel_source


def test_segment_round_trip_and_invalidation(tmp_path):
    docs = make_docs(800, seed=42)
    seg = str(tmp_path / "seg" / "d.n1seg")
    os.makedirs(os.path.dirname(seg))
    a = shredded(docs).set_segment_output(seg, "v1")
    dicts = [sorted(a.dictionary(c)) for c in range(len(COLS))]
    stats = [a.stats(c) for c in range(len(COLS))]
    a.seal()
    assert os.path.exists(seg) and not os.path.exists(seg + ".tmp")
    b = q.Table(COLS)
    assert b.load_segment(seg, "v1") and b.num_rows == 800
    for c in range(len(COLS)):
        assert b.dictionary(c) == dicts[c]  # sorted: the payloads of a segment are dictionary ranks
        sb = b.stats(c)
        assert (sb[:5] == stats[c][:5]).all() and sb[7] == stats[c][7]
    b.seal()
    for c in range(len(COLS)):
        assert a.scan_bytes(c) == b.scan_bytes(c)
    ka = q.Query(a, "d", WHERE, KEYS, AGGS).kernel_source
    kb = q.Query(b, "d", WHERE, KEYS, AGGS).kernel_source
    assert ka == kb  # same statistics, same dictionaries: the very same specialised kernel
    # stale tag, other columns, truncated or foreign files: not loaded, table untouched
    assert not q.Table(COLS).load_segment(seg, "v2")
    assert not q.Table(COLS[:-1]).load_segment(seg, "v1")
    assert not q.Table(list(reversed(COLS))).load_segment(seg, "v1")
    assert not q.Table(COLS).load_segment(str(tmp_path / "absent.n1seg"), "v1")
    blob = open(seg, "rb").read()
    for cut in (4, 20, len(blob) // 2, len(blob) - 1):
        bad = str(tmp_path / ("cut%d" % cut))
        open(bad, "wb").write(blob[:cut])
        t = q.Table(COLS)
        assert not t.load_segment(bad, "v1") and t.num_rows == 0
    # a flipped payload byte fails the checksum; a header that claims more rows / dictionary bytes than the file holds is
    # refused before anything is allocated for it; a dictionary out of order is refused (constants are binary-searched)
    flipped = bytearray(blob)
    flipped[len(blob) - 100] ^= 0x40
    open(str(tmp_path / "flipped"), "wb").write(bytes(flipped))
    assert not q.Table(COLS).load_segment(str(tmp_path / "flipped"), "v1")
    import re
    import struct
    hl = struct.unpack("<Q", blob[8:16])[0]
    header = blob[16:16 + hl].decode()
    for pat, repl in ((r'"nrows":800', '"nrows":80000000000'), (r'"dict_bytes":(\d+)', '"dict_bytes":9\\g<1>')):
        h2 = re.sub(pat, repl, header, count=1).encode()
        assert h2 != header.encode()
        open(str(tmp_path / "lying"), "wb").write(blob[:8] + struct.pack("<Q", len(h2)) + h2 + blob[16 + hl:])
        t = q.Table(COLS)
        assert not t.load_segment(str(tmp_path / "lying"), "v1") and t.num_rows == 0
    open(str(tmp_path / "junk"), "wb").write(b"not a segment at all" * 10)
    assert not q.Table(COLS).load_segment(str(tmp_path / "junk"), "v1")
    with pytest.raises(q.N1GpuError):
        shredded(docs).load_segment(seg, "v1")  # only into an empty table


def test_operator_keeps_one_segment_per_keyspace_and_column_set(tmp_path, monkeypatch):
    docs = make_docs(300, seed=43)
    root, segs = str(tmp_path / "ds"), str(tmp_path / "segments")
    os.makedirs(segs)
    write_keyspace(root, "default", "d", [("k%05d" % i, t) for i, t in enumerate(docs)])
    plan = explain_plan("default", "d", "d", WHERE, KEYS, AGGS)
    q.set_segment_dir(segs)
    try:
        q.Operator(plan, root)
        files = sorted(os.listdir(segs))
        assert len(files) == 1 and files[0].endswith(".n1seg")
        first = os.stat(os.path.join(segs, files[0])).st_mtime_ns
        # an UPDATE rewrites a document in place (file.go:375-471): directory mtime stays, the segment must not survive
        time.sleep(0.02)
        doc = os.path.join(root, "default", "d", "k00007.json")
        open(doc, "w").write('{"t": "brand-new-type", "p": 5}')
        os.utime(os.path.join(root, "default", "d"), None)  # new directory mtime: the in-process table cache misses too
        op = q.Operator(plan, root)
        assert os.stat(os.path.join(segs, files[0])).st_mtime_ns > first  # rewritten under the new source tag
        other = explain_plan("default", "d", "d", None, [F("h")], ["count(*)"])
        q.Operator(other, root)
        assert len(os.listdir(segs)) == 2  # another column set, another segment
    finally:
        q.set_segment_dir("")


@pytest.mark.gpu
def test_queries_through_segment_loaded_tables(tmp_path):
    q.init(0)
    docs = make_docs(4000, seed=44)
    seg = str(tmp_path / "d.n1seg")
    a = shredded(docs).set_segment_output(seg, "x").seal()
    b = q.Table(COLS)
    assert b.load_segment(seg, "x")
    b.seal()
    exp = oracle_rows(docs, "d", WHERE, KEYS, AGGS)
    for t in (a, b):
        assert_same(exp, gpu_rows(q.Query(t, "d", WHERE, KEYS, AGGS).execute(), AGGS), "segment")
    path = str(tmp_path / "d.ndjson")
    open(path, "w", encoding="utf-8").write("\n".join(docs) + "\n")
    for threads in (0, -1):
        c = q.Table(COLS).load_ndjson(path, threads=threads).seal()
        assert_same(exp, gpu_rows(q.Query(c, "d", WHERE, KEYS, AGGS).execute(), AGGS), "ndjson threads=%d" % threads)


@pytest.mark.gpu
def test_operator_reads_a_packed_ndjson_keyspace(tmp_path):
    """<namespace>/<keyspace>.ndjson (one document per line, primary-key order) stands in for the directory of one file per
    document: the operator shreds it on the device (pinned parallel read, chunked H2D) and returns what the oracle computes;
    a rewritten file invalidates the resident table."""
    import query_b200 as q
    from gen_n1 import F, make_docs
    from oracle import n1ql_oracle as O
    from plans_n1 import explain_plan
    from util_n1 import assert_same, gpu_rows, oracle_rows
    docs = make_docs(4000, seed=77)
    ns = tmp_path / "default"
    ns.mkdir()
    (ns / "d.ndjson").write_text("\n".join(docs[:3000]) + "\n\n  \n")  # trailing blank lines are not documents
    keys, aggs = [F("t")], sorted({"count(*)", "sum(%s)" % F("p"), "max(%s)" % F("s")})
    plan = explain_plan("default", "d", "d", "(%s is not missing)" % F("p"), keys, aggs)
    got = gpu_rows(q.Operator(plan, str(tmp_path)).run_once(), aggs)
    assert_same(oracle_rows(docs[:3000], "d", "(%s is not missing)" % F("p"), keys, aggs), got, "packed keyspace")
    import os
    import time
    time.sleep(0.01)
    (ns / "d.ndjson").write_text("\n".join(docs) + "\n")
    got = gpu_rows(q.Operator(plan, str(tmp_path)).run_once(), aggs)
    assert_same(oracle_rows(docs, "d", "(%s is not missing)" % F("p"), keys, aggs), got, "packed keyspace, rewritten")
