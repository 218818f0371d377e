"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle_ref.so (oracle_ref.c), the
reference-shaped C restatement used (a) as a faster checker for mid-size parity tests and (b) as bench.py's
cpu_baseline / --impl reference arm.  The product (query_b200/) never imports this module."""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np

from . import n1ql_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_ref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "oracle_ref.c")):
            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        L = C.CDLL(_SO)
        L.oracle_run.restype = C.c_void_p
        L.oracle_run.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_longlong, C.c_char_p, C.POINTER(C.c_char_p), C.c_int,
                                 C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_gen_docs.restype = C.c_longlong
        L.oracle_gen_docs.argtypes = [C.c_int, C.c_ulonglong, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


def _hex(s):
    return s.encode("utf-8").hex() if isinstance(s, str) else bytes(s).hex()


def sexpr(e):
    """oracle parse tree -> the S-expression oracle_ref.c interprets."""
    if isinstance(e, str):
        e = O.parse(e)
    if isinstance(e, O.Constant):
        v = e.v
        if v is O.MISSING:
            return "(cmissing)"
        if v is None:
            return "(cnull)"
        if isinstance(v, bool):
            return "(cb %d)" % int(v)
        if isinstance(v, int):
            return "(ci %d)" % v
        if isinstance(v, float):
            return "(cf %016x)" % struct.unpack("<Q", struct.pack("<d", v))[0]
        if isinstance(v, str):
            return "(cs %s)" % _hex(v)
        if isinstance(v, list):
            raise ValueError("array constant outside IN")
        raise ValueError(v)
    if isinstance(e, O.Field):
        path, x = [], e
        while isinstance(x, O.Field):
            path.append(x.name)
            x = x.first
        if not isinstance(x, O.Identifier):
            raise ValueError("field root")
        return "(field %s)" % " ".join(_hex(p) for p in reversed(path))
    if isinstance(e, O.In):
        x, arr = e.ops
        if isinstance(arr, O.Constant):
            elems = [O.Constant(O.new_value(v)) for v in arr.v]
        else:
            elems = arr.ops
        return "(in %s)" % " ".join([sexpr(x)] + [sexpr(v) for v in elems])
    names = {O.Add: "add", O.Mult: "mult", O.Sub: "sub", O.Div: "div", O.Mod: "mod", O.Eq: "eq", O.LT: "lt", O.LE: "le",
             O.And: "and", O.Or: "or"}
    for cls, n in names.items():
        if type(e) is cls:
            return "(%s %s)" % (n, " ".join(sexpr(o) for o in e.ops))
    if isinstance(e, O.Between):
        return "(between %s)" % " ".join(sexpr(o) for o in e.ops)
    unary = {O.Neg: "neg", O.Not: "not", O.IsNull: "isnull", O.IsNotNull: "isnotnull", O.IsMissing: "ismissing",
             O.IsNotMissing: "isnotmissing", O.IsValued: "isvalued", O.IsNotValued: "isnotvalued"}
    for cls, n in unary.items():
        if type(e) is cls:
            return "(%s %s)" % (n, sexpr(e.op))
    raise ValueError("no S-expression for %r" % (e,))


def agg_spec(text):
    a = O.parse(text) if isinstance(text, str) else text
    return "%s %d %s" % (a.name, int(a.distinct), "*" if a.operand is None else sexpr(a.operand))


def pack_docs(docs):
    parts = [d.encode("utf-8") if isinstance(d, str) else bytes(d) for d in docs]
    offs = np.zeros(len(parts) + 1, dtype=np.int64)
    if parts:
        np.cumsum([len(p) for p in parts], out=offs[1:])
    return np.frombuffer(b"".join(parts) + b"\0", dtype=np.uint8), offs


def _decode(v):
    if isinstance(v, dict):
        if "$missing" in v:
            return O.MISSING
        if "$f" in v:
            return struct.unpack("<d", struct.pack("<Q", int(v["$f"], 16)))[0]
        if "$s" in v:
            return bytes.fromhex(v["$s"]).decode("utf-8", "surrogateescape")
    return v


def run(buf, offs, alias, where, keys, aggs, threads=1):
    """Returns (groups, seconds, rows_passed); groups = [(keys list, aggregates list)] with oracle python values."""
    L = lib()
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    w = sexpr(where).encode() if where else None
    ks = [sexpr(k).encode() for k in keys]
    ags = [agg_spec(a).encode() for a in aggs]
    karr = (C.c_char_p * max(1, len(ks)))(*ks)
    aarr = (C.c_char_p * max(1, len(ags)))(*ags)
    el, passed = C.c_double(), C.c_longlong()
    p = L.oracle_run(buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.POINTER(C.c_longlong)), len(offs) - 1, w, karr, len(ks),
                     aarr, len(ags), threads, C.byref(el), C.byref(passed))
    if not p:
        raise RuntimeError("oracle_ref rejected the query")
    try:
        text = C.string_at(p).decode("utf-8")
    finally:
        L.oracle_free(p)
    groups = [([_decode(k) for k in g["k"]], [_decode(a) for a in g["a"]]) for g in json.loads(text)]
    return groups, el.value, passed.value


def rows(docs_or_packed, alias, where, keys, aggs, threads=1):
    """Same comparable form as tests/util_n1.oracle_rows."""
    buf, offs = docs_or_packed if isinstance(docs_or_packed, tuple) else pack_docs(docs_or_packed)
    groups, _s, _p = run(buf, offs, alias, where, keys, aggs, threads)
    out = {}
    for ks, ag in groups:
        k = tuple("MISSING" if v is O.MISSING else json.dumps(O.to_python(v), sort_keys=True) for v in ks)
        assert k not in out
        out[k] = dict(zip(aggs, ag))
    return out


def gen_docs(config, seed, first, n, newline=False):
    """Synthetic documents of BASELINE.json config 2..5 as (uint8 buffer, int64 offsets); newline=True ends every document
    with a line end (NDJSON: the buffer is then a packed keyspace file as it stands)."""
    L = lib()
    if newline:
        config |= 0x100
    cap = int(n) * 320 + 1024
    buf = np.empty(cap, dtype=np.uint8)
    offs = np.empty(int(n) + 1, dtype=np.int64)
    used = L.oracle_gen_docs(config, seed, first, n, buf.ctypes.data_as(C.c_void_p), cap, offs.ctypes.data_as(C.POINTER(C.c_longlong)))
    if used < 0:
        raise RuntimeError("generator buffer too small")
    return buf[: used + 1], offs
