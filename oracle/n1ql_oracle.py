"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A document-at-a-time CPU restatement (pure Python, small cases) of the reference's
`PrimaryScan/Fetch -> Filter -> InitialGroup -> IntermediateGroup -> FinalGroup` chain with
COUNT/COUNTN/SUM/AVG/MIN/MAX and the DISTINCT variants, in single-stream semantics
(cbq-engine's default -max-parallelism=1, server/cbq-engine/main.go:55).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker.  The product (query_b200/) never imports it.

Parity pinning: this restatement is checked in tests/test_oracle_golden.py against the
reference's own golden vectors: test/filestore/json/default/cases/
case_group_by_having.json, case_where.json, test/multistore/test_cases/aggregate_functions/
{case_group_by_having,case_distinct}.json and integers/case_select.json, with the fixture
documents extracted by tests/golden/make_golden.py.  The reference itself (Go) cannot be
built in this image (no Go toolchain; couchbase/go_json is un-vendored and un-pinned), so the
JSON-decoding boundary (duplicate object keys, exotic escapes) is "parity unpinned".

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import json
import math
import struct
from decimal import Decimal

# --------------------------------------------------------------------------------------
# Value model (value/value.go:69-79 type order)
# --------------------------------------------------------------------------------------


class _Missing:
    __slots__ = ()

    def __repr__(self):
        return "MISSING"

    def __bool__(self):
        return False


MISSING = _Missing()

T_MISSING, T_NULL, T_BOOLEAN, T_NUMBER, T_STRING, T_ARRAY, T_OBJECT, T_JSON, T_BINARY = range(9)

INT64_MIN = -(1 << 63)
INT64_MAX = (1 << 63) - 1


def vtype(v):
    """value.Value.Type()"""
    if v is MISSING:
        return T_MISSING
    if v is None:
        return T_NULL
    if isinstance(v, bool):
        return T_BOOLEAN
    if isinstance(v, (int, float)):
        return T_NUMBER
    if isinstance(v, str):
        return T_STRING
    if isinstance(v, list):
        return T_ARRAY
    if isinstance(v, dict):
        return T_OBJECT
    if isinstance(v, (bytes, bytearray)):
        return T_BINARY
    raise TypeError(type(v))


def _go_int64_of_float(x: float) -> int:
    """Go's int64(float64) on amd64: out-of-range / NaN -> MinInt64."""
    if math.isnan(x) or math.isinf(x) or x >= 9.223372036854775808e18 or x < -9.223372036854775808e18:
        return INT64_MIN
    return int(x)


def is_int(x: float) -> bool:
    """value/integer.go:354-356  IsInt: x == float64(int64(x))"""
    return x == float(_go_int64_of_float(x))


def new_value(v):
    """value.NewValue canonicalisation (value/value.go:367-430): integral float64 -> intValue."""
    if isinstance(v, bool) or v is None or v is MISSING:
        return v
    if isinstance(v, float):
        if is_int(v):
            return _go_int64_of_float(v)
        return v
    if isinstance(v, int):
        if INT64_MIN <= v <= INT64_MAX:
            return v
        return new_value(float(v))  # JSON integer literal outside int64 decodes as float64
    if isinstance(v, list):
        return [new_value(x) for x in v]
    if isinstance(v, dict):
        return {k: new_value(x) for k, x in v.items()}
    return v


def _first_wins(pairs):
    """json.FirstFind returns the FIRST occurrence of a duplicate name (value/parsed.go:193);
    parity unpinned (go_json is not in the tree) - generators never emit duplicates."""
    d = {}
    for k, v in pairs:
        if k not in d:
            d[k] = v
    return d


def parse_document(raw):
    """datastore/file fetch (file.go:732-743) + value.NewParsedValue (value/parsed.go:38-67):
    type sniff on the first significant byte (parsed.go:76-98), invalid JSON -> BINARY."""
    if isinstance(raw, str):
        raw = raw.encode("utf-8")
    sniff = T_BINARY
    for b in raw:
        c = chr(b)
        if c == "{":
            sniff = T_OBJECT
        elif c == "[":
            sniff = T_ARRAY
        elif c == '"':
            sniff = T_STRING
        elif c in "0123456789-":
            sniff = T_NUMBER
        elif c in "tf":
            sniff = T_BOOLEAN
        elif c == "n":
            sniff = T_NULL
        elif c in " \t\n":
            continue
        break
    if sniff == T_BINARY:
        return bytes(raw)
    try:
        v = json.loads(raw.decode("utf-8"), object_pairs_hook=_first_wins)
    except Exception:
        return bytes(raw)
    return new_value(v)


def field(v, name):
    """Value.Field: non-object or absent name -> MISSING (nav_field.go:134-160, parsed.go:159-207)."""
    if isinstance(v, dict):
        return v.get(name, MISSING)
    return MISSING


def truth(v) -> bool:
    """Value.Truth (missing.go:113, null.go:106, boolean.go:131, integer.go:136, float.go:190, string.go:148)."""
    t = vtype(v)
    if t <= T_NULL:
        return False
    if t == T_BOOLEAN:
        return v
    if t == T_NUMBER:
        return not (isinstance(v, float) and math.isnan(v)) and v != 0
    if t == T_STRING:
        return len(v) > 0
    if t == T_ARRAY or t == T_OBJECT:
        return len(v) > 0
    return False


def _collate_float(t: float, o: float) -> int:
    """value/float.go:123-172"""
    if math.isnan(t):
        return 0 if math.isnan(o) else -1
    if math.isnan(o):
        return 1
    if t == -math.inf:
        return 0 if o == -math.inf else -1
    if o == -math.inf:
        return 1
    if t == math.inf:
        return 0 if o == math.inf else 1
    if o == math.inf:
        return -1
    r = t - o
    return -1 if r < 0 else (1 if r > 0 else 0)


def collate(a, b) -> int:
    """Value.Collate: total order across types (integer.go:100-118, float.go:106-121,
    string.go:116-130, boolean.go:100-114, null.go:89-91, missing.go:101-103)."""
    ta, tb = vtype(a), vtype(b)
    if ta != tb:
        return ta - tb
    if ta <= T_NULL:
        return 0
    if ta == T_BOOLEAN:
        return 0 if a == b else (-1 if not a else 1)
    if ta == T_NUMBER:
        if isinstance(a, int) and isinstance(b, int):
            return -1 if a < b else (1 if a > b else 0)
        return _collate_float(float(a), float(b))
    if ta == T_STRING:
        ab, bb = a.encode("utf-8"), b.encode("utf-8")  # Go string < is bytewise
        return -1 if ab < bb else (1 if ab > bb else 0)
    if ta == T_ARRAY:
        for x, y in zip(a, b):
            c = collate(x, y)
            if c:
                return c
        return len(a) - len(b)
    if ta == T_OBJECT:
        if len(a) != len(b):
            return len(a) - len(b)
        na, nb = sorted(a), sorted(b)
        for x, y in zip(na, nb):
            if x != y:
                return -1 if x < y else 1
            c = collate(a[x], b[y])
            if c:
                return c
        return 0
    return 0


def compare(a, b):
    """Value.Compare: MISSING if either MISSING, else NULL if either NULL, else Collate sign
    (integer.go:120-130, null.go:93-101, missing.go:105-107)."""
    if a is MISSING or b is MISSING:
        return MISSING
    if a is None or b is None:
        return None
    return collate(a, b)


def equals(a, b):
    """Value.Equals (integer.go:68-86, float.go:74-92, string.go:82-96, boolean.go:72-86,
    null.go:72-80, missing.go:88-90)."""
    if a is MISSING or b is MISSING:
        return MISSING
    if a is None or b is None:
        return None
    ta, tb = vtype(a), vtype(b)
    if ta != tb:
        return False
    if ta == T_NUMBER:
        if isinstance(a, int) and isinstance(b, int):
            return a == b
        return float(a) == float(b)
    if ta in (T_ARRAY, T_OBJECT):
        return collate(a, b) == 0
    return a == b


# ---- NumberValue arithmetic (value/integer.go:266-348, value/float.go:331-381) ------------


def _wrap64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def num_add(a, b):
    if isinstance(a, int) and isinstance(b, int):
        rv = _wrap64(a + b)
        if (a >= 0 and b >= 0 and rv >= 0) or (a < 0 and b < 0 and rv < 0):
            return rv
    return float(a) + float(b)


def num_mult(a, b):
    if isinstance(a, int) and isinstance(b, int):
        rv = _wrap64(a * b)
        # integer.go:319-329: rv/this == n with Go truncated division (and MinInt64/-1 wraps)
        if a == 0:
            return rv
        if a == -1 and rv == INT64_MIN:
            q = INT64_MIN
        else:
            q = abs(rv) // abs(a)
            if (rv < 0) != (a < 0):
                q = -q
        if q == b:
            return rv
    return float(a) * float(b)


def num_neg(a):
    if isinstance(a, int):
        if a == INT64_MIN:
            return -float(a)
        return -a
    return -a


def num_sub(a, b):
    if isinstance(a, int):
        if isinstance(b, int) and b > INT64_MIN:
            return num_add(a, -b)
        return float(a) - float(b)
    return a - float(b)


# ---- rendering (value/float.go:31-48, integer MarshalJSON) -------------------------------


def format_number(v) -> str:
    if isinstance(v, int):
        return str(v)
    if math.isnan(v):
        return '"NaN"'
    if math.isinf(v):
        return '"+Infinity"' if v > 0 else '"-Infinity"'
    if v == 0:
        v = 0.0
    return format(Decimal(repr(v)), "f")  # strconv.FormatFloat(f,'f',-1,64): shortest digits, no exponent


def marshal(v) -> str:
    """Canonical JSON text as Value.MarshalJSON writes it (object names sorted, object.go:30-78)."""
    t = vtype(v)
    if t == T_MISSING:
        return "missing"
    if t == T_NULL:
        return "null"
    if t == T_BOOLEAN:
        return "true" if v else "false"
    if t == T_NUMBER:
        return format_number(v)
    if t == T_STRING:
        return json.dumps(v, ensure_ascii=False)
    if t == T_ARRAY:
        return "[" + ",".join("null" if x is MISSING else marshal(x) for x in v) + "]"
    if t == T_OBJECT:
        return "{" + ",".join(json.dumps(k, ensure_ascii=False) + ":" + marshal(v[k])
                              for k in sorted(v) if v[k] is not MISSING) + "}"
    return json.dumps(repr(v))


def to_python(v):
    """Result value -> what json.loads of the reference's result row would hold (numbers: an
    integral floatValue prints as an integer, so it reads back as int)."""
    t = vtype(v)
    if t == T_NUMBER:
        return json.loads(format_number(v))
    if t == T_ARRAY:
        return [to_python(x) for x in v]
    if t == T_OBJECT:
        return {k: to_python(x) for k, x in v.items() if x is not MISSING}
    return v


# --------------------------------------------------------------------------------------
# Expressions (expression/*.go) with the Stringer text form (expression/stringer.go)
# --------------------------------------------------------------------------------------


class Expr:
    def evaluate(self, item):
        raise NotImplementedError

    def children(self):
        return []

    def __str__(self):
        raise NotImplementedError

    def __repr__(self):
        return str(self)


class Constant(Expr):
    def __init__(self, v):
        self.v = v

    def evaluate(self, item):  # constant.go:51-53
        return self.v

    def __str__(self):  # stringer.go VisitConstant
        return marshal(self.v)


class Identifier(Expr):
    def __init__(self, name):
        self.name = name

    def evaluate(self, item):  # identifier.go:48-51 -> item.Field(name)
        return field(item, self.name)

    def __str__(self):
        return "`%s`" % self.name


class Field(Expr):
    def __init__(self, first, name):
        self.first, self.name = first, name

    def evaluate(self, item):  # nav_field.go:134-160
        return field(self.first.evaluate(item), self.name)

    def children(self):
        return [self.first]

    def __str__(self):
        return "(%s.`%s`)" % (self.first, self.name)


class ArrayConstruct(Expr):
    def __init__(self, ops):
        self.ops = ops

    def evaluate(self, item):
        return [o.evaluate(item) for o in self.ops]

    def children(self):
        return list(self.ops)

    def __str__(self):
        return "[" + ", ".join(str(o) for o in self.ops) + "]"


class _Nary(Expr):
    sym = "?"

    def __init__(self, *ops):
        self.ops = list(ops)

    def children(self):
        return list(self.ops)

    def __str__(self):
        return "(" + (" %s " % self.sym).join(str(o) for o in self.ops) + ")"


class Add(_Nary):
    sym = "+"

    def evaluate(self, item):  # arith_add.go:51-70
        null = False
        s = 0
        for op in self.ops:
            a = op.evaluate(item)
            if not null and vtype(a) == T_NUMBER:
                s = num_add(s, a)
            elif a is MISSING:
                return MISSING
            else:
                null = True
        return None if null else s


class Mult(_Nary):
    sym = "*"

    def evaluate(self, item):  # arith_mult.go:51-70
        null = False
        p = 1
        for op in self.ops:
            a = op.evaluate(item)
            if not null and vtype(a) == T_NUMBER:
                p = num_mult(p, a)
            elif a is MISSING:
                return MISSING
            else:
                null = True
        return None if null else p


class Sub(_Nary):
    sym = "-"

    def evaluate(self, item):  # arith_sub.go:53-61
        a, b = (o.evaluate(item) for o in self.ops)
        if vtype(a) == T_NUMBER and vtype(b) == T_NUMBER:
            return num_sub(a, b)
        if a is MISSING or b is MISSING:
            return MISSING
        return None


class Div(_Nary):
    sym = "/"

    def evaluate(self, item):  # arith_div.go:46-64
        a, b = (o.evaluate(item) for o in self.ops)
        if a is MISSING or b is MISSING:
            return MISSING
        if vtype(b) == T_NUMBER:
            s = float(b)
            if s == 0.0:
                return None
            if vtype(a) == T_NUMBER:
                return new_value(float(a) / s)
        return None


class Mod(_Nary):
    sym = "%"

    def evaluate(self, item):  # arith_mod.go:48-66
        a, b = (o.evaluate(item) for o in self.ops)
        if a is MISSING or b is MISSING:
            return MISSING
        if vtype(b) == T_NUMBER:
            s = float(b)
            if s == 0.0:
                return None
            if vtype(a) == T_NUMBER:
                return new_value(math.fmod(float(a), s))
        return None


class Neg(Expr):
    def __init__(self, op):
        self.op = op

    def children(self):
        return [self.op]

    def evaluate(self, item):  # arith_neg.go:51-59
        a = self.op.evaluate(item)
        if vtype(a) == T_NUMBER:
            return num_neg(a)
        if a is MISSING:
            return MISSING
        return None

    def __str__(self):
        return "(-%s)" % self.op


class Eq(_Nary):
    sym = "="

    def evaluate(self, item):  # comp_eq.go:76-78
        a, b = (o.evaluate(item) for o in self.ops)
        return equals(a, b)


class LT(_Nary):
    sym = "<"

    def evaluate(self, item):  # comp_lt.go:57-65
        a, b = (o.evaluate(item) for o in self.ops)
        c = compare(a, b)
        if isinstance(c, int):
            return c < 0
        return c


class LE(_Nary):
    sym = "<="

    def evaluate(self, item):  # comp_le.go:57-65
        a, b = (o.evaluate(item) for o in self.ops)
        c = compare(a, b)
        if isinstance(c, int):
            return c <= 0
        return c


class Between(Expr):
    def __init__(self, x, lo, hi):
        self.ops = [x, lo, hi]

    def children(self):
        return list(self.ops)

    def evaluate(self, item):  # comp_between.go:58-78
        x, lo, hi = (o.evaluate(item) for o in self.ops)
        lc = compare(x, lo)
        if lc is MISSING:
            return MISSING
        hc = compare(x, hi)
        if hc is MISSING:
            return MISSING
        if isinstance(lc, int) and isinstance(hc, int):
            return lc >= 0 and hc <= 0
        return None

    def __str__(self):
        return "(%s between %s and %s)" % tuple(self.ops)


class In(_Nary):
    sym = "in"

    def evaluate(self, item):  # coll_in.go:61-91
        a, b = (o.evaluate(item) for o in self.ops)
        if a is MISSING or b is MISSING:
            return MISSING
        if vtype(b) != T_ARRAY:
            return None
        missing = null = False
        for s in b:
            v = new_value(s)
            if vtype(a) > T_NULL and vtype(v) > T_NULL:
                if equals(a, v) is True:
                    return True
            elif v is MISSING:
                missing = True
            else:
                null = True
        if null:
            return None
        if missing:
            return MISSING
        return False


class And(_Nary):
    sym = "and"

    def evaluate(self, item):  # logic_and.go:64-88
        missing = null = False
        for op in self.ops:
            a = op.evaluate(item)
            if a is None:
                null = True
            elif a is MISSING:
                missing = True
            elif not truth(a):
                return False
        if missing:
            return MISSING
        if null:
            return None
        return True


class Or(_Nary):
    sym = "or"

    def evaluate(self, item):  # logic_or.go:98-122
        missing = null = False
        for op in self.ops:
            a = op.evaluate(item)
            if a is None:
                null = True
            elif a is MISSING:
                missing = True
            elif truth(a):
                return True
        if null:
            return None
        if missing:
            return MISSING
        return False


class Not(Expr):
    def __init__(self, op):
        self.op = op

    def children(self):
        return [self.op]

    def evaluate(self, item):  # logic_not.go:57-68
        a = self.op.evaluate(item)
        if a is MISSING or a is None:
            return a
        return not truth(a)

    def __str__(self):
        return "(not %s)" % self.op


class _IsTest(Expr):
    text = "?"

    def __init__(self, op):
        self.op = op

    def children(self):
        return [self.op]

    def __str__(self):
        return "(%s %s)" % (self.op, self.text)


class IsNull(_IsTest):
    text = "is null"

    def evaluate(self, item):  # comp_null.go:58-67
        a = self.op.evaluate(item)
        if a is None:
            return True
        if a is MISSING:
            return MISSING
        return False


class IsNotNull(_IsTest):
    text = "is not null"

    def evaluate(self, item):  # comp_null.go:116-125
        a = self.op.evaluate(item)
        if a is None:
            return False
        if a is MISSING:
            return MISSING
        return True


class IsMissing(_IsTest):
    text = "is missing"

    def evaluate(self, item):  # comp_missing.go:62-69
        return self.op.evaluate(item) is MISSING


class IsNotMissing(_IsTest):
    text = "is not missing"

    def evaluate(self, item):  # comp_missing.go:125-132
        return self.op.evaluate(item) is not MISSING


class IsValued(_IsTest):
    text = "is valued"

    def evaluate(self, item):  # comp_valued.go:61-68
        return vtype(self.op.evaluate(item)) > T_NULL


class IsNotValued(_IsTest):
    text = "is not valued"

    def evaluate(self, item):  # comp_valued.go:124-131
        return vtype(self.op.evaluate(item)) <= T_NULL


# --------------------------------------------------------------------------------------
# Aggregates (algebra/agg_*.go).  State machines: default / initial / intermediate / final.
# --------------------------------------------------------------------------------------


class ValueSet:
    """value.Set (value/set.go:22-110): numbers keyed by int64 when integral else by float64."""

    def __init__(self):
        self.booleans, self.floats, self.ints, self.strings, self.others = {}, {}, {}, {}, {}

    def add(self, v):
        t = vtype(v)
        if t == T_BOOLEAN:
            self.booleans[v] = v
        elif t == T_NUMBER:
            if isinstance(v, float):
                if is_int(v):
                    self.ints[_go_int64_of_float(v)] = v
                else:
                    self.floats[v] = v
            else:
                self.ints[v] = v
        elif t == T_STRING:
            self.strings[v] = v
        else:
            self.others[marshal(v)] = v

    def __len__(self):
        return len(self.booleans) + len(self.floats) + len(self.ints) + len(self.strings) + len(self.others)

    def values(self):  # set.go:217-265 order of maps (Go map order inside each is random)
        return (list(self.booleans.values()) + list(self.floats.values()) + list(self.ints.values())
                + list(self.strings.values()) + list(self.others.values()))

    def union(self, other):  # agg_util.go:51-81 cumulateSets
        for v in other.values():
            self.add(v)
        return self


class Aggregate(Expr):
    name = "?"
    distinct = False

    def __init__(self, operand):
        self.operand = operand

    def children(self):
        return [] if self.operand is None else [self.operand]

    def __str__(self):  # stringer.go VisitFunction :581-604
        return "%s(%s%s)" % (self.name, "distinct " if self.distinct else "",
                             "*" if self.operand is None else str(self.operand))

    def evaluate(self, item):  # algebra/aggregate.go:97-118: the value FinalGroup left under the aggregate's text
        return item[AGGREGATES][str(self)]

    def default(self):
        return None

    def initial(self, item, cum):
        raise NotImplementedError

    def intermediate(self, part, cum):
        raise NotImplementedError

    def final(self, cum):
        return cum


class Count(Aggregate):  # agg_count.go:95-149
    name = "count"

    def default(self):
        return 0

    def initial(self, item, cum):
        if self.operand is not None:
            if vtype(self.operand.evaluate(item)) <= T_NULL:
                return cum
        return num_add(cum, 1)

    def intermediate(self, part, cum):
        return num_add(cum, part)


class Countn(Aggregate):  # agg_countn.go:77-129
    name = "countn"

    def default(self):
        return 0

    def initial(self, item, cum):
        if vtype(self.operand.evaluate(item)) != T_NUMBER:
            return cum
        return num_add(cum, 1)

    def intermediate(self, part, cum):
        return num_add(cum, part)


class Sum(Aggregate):  # agg_sum.go:77-136
    name = "sum"

    def initial(self, item, cum):
        v = self.operand.evaluate(item)
        if vtype(v) != T_NUMBER:
            return cum
        return self.intermediate(v, cum)

    def intermediate(self, part, cum):
        if part is None:
            return cum
        if cum is None:
            return part
        return num_add(cum, part)


class Avg(Aggregate):  # agg_avg.go:77-157
    name = "avg"

    def initial(self, item, cum):
        v = self.operand.evaluate(item)
        if vtype(v) != T_NUMBER:
            return cum
        return self.intermediate({"sum": v, "count": 1}, cum)

    def intermediate(self, part, cum):
        if part is None:
            return cum
        if cum is None:
            return part
        return {"sum": num_add(cum["sum"], part["sum"]), "count": num_add(cum["count"], part["count"])}

    def final(self, cum):
        if cum is None:
            return None
        if float(cum["count"]) > 0.0:
            return new_value(float(cum["sum"]) / float(cum["count"]))
        return None


class Min(Aggregate):  # agg_min.go:76-127
    name = "min"

    def initial(self, item, cum):
        v = self.operand.evaluate(item)
        if vtype(v) <= T_NULL:
            return cum
        return self.intermediate(v, cum)

    def intermediate(self, part, cum):
        if part is None:
            return cum
        if cum is None:
            return part
        return part if collate(part, cum) < 0 else cum


class Max(Aggregate):  # agg_max.go:76-127
    name = "max"

    def initial(self, item, cum):
        v = self.operand.evaluate(item)
        if vtype(v) <= T_NULL:
            return cum
        return self.intermediate(v, cum)

    def intermediate(self, part, cum):
        if part is None:
            return cum
        if cum is None:
            return part
        return part if collate(part, cum) > 0 else cum


class _DistinctBase(Aggregate):
    distinct = True
    numbers_only = False

    def initial(self, item, cum):  # setAdd agg_util.go:30-46
        v = self.operand.evaluate(item)
        if self.numbers_only:
            if vtype(v) != T_NUMBER:
                return cum
        elif vtype(v) <= T_NULL:
            return cum
        if not isinstance(cum, ValueSet):
            cum = ValueSet()
        cum.add(v)
        return cum

    def intermediate(self, part, cum):
        # Single-stream oracle: guards as in agg_count_distinct.go:103-111.  (SumDistinct/
        # AvgDistinct lack the guard in the reference - a latent multi-stream bug)
        if not isinstance(part, ValueSet):
            return cum
        if not isinstance(cum, ValueSet):
            return part
        return cum.union(part)


class CountDistinct(_DistinctBase):  # agg_count_distinct.go:76-126
    name = "count"

    def default(self):
        return 0

    def final(self, cum):
        return len(cum) if isinstance(cum, ValueSet) else cum


class CountnDistinct(_DistinctBase):  # agg_countn_distinct.go:76-126
    name = "countn"
    numbers_only = True

    def default(self):
        return 0

    def final(self, cum):
        return len(cum) if isinstance(cum, ValueSet) else cum


class SumDistinct(_DistinctBase):  # agg_sum_distinct.go:78-133
    name = "sum"
    numbers_only = True

    def final(self, cum):
        if not isinstance(cum, ValueSet) or len(cum) == 0:
            return None
        s = 0  # value.ZERO_NUMBER: `0 + negative` takes the mixed-sign float branch of intValue.Add
        for v in cum.values():
            s = num_add(s, v)
        return s


class AvgDistinct(_DistinctBase):  # agg_avg_distinct.go:78-134
    name = "avg"
    numbers_only = True

    def final(self, cum):
        if not isinstance(cum, ValueSet) or len(cum) == 0:
            return None
        s = 0
        for v in cum.values():
            s = num_add(s, v)
        return new_value(float(s) / float(len(cum)))


_AGGS = {  # algebra/agg_registry.go:41-62 (array_agg is out of scope; no MIN/MAX DISTINCT: n1ql.y:2756-2764)
    ("count", False): Count, ("countn", False): Countn, ("sum", False): Sum, ("avg", False): Avg,
    ("min", False): Min, ("max", False): Max,
    ("count", True): CountDistinct, ("countn", True): CountnDistinct,
    ("sum", True): SumDistinct, ("avg", True): AvgDistinct,
}

# --------------------------------------------------------------------------------------
# Parser for the Stringer text form (what plan JSON carries: plan/filter.go, plan/group.go)
# --------------------------------------------------------------------------------------


class ParseError(ValueError):
    pass


class _P:
    def __init__(self, s):
        self.s, self.i = s, 0

    def ws(self):
        while self.i < len(self.s) and self.s[self.i] in " \t\n":
            self.i += 1

    def peek(self, k=1):
        self.ws()
        return self.s[self.i:self.i + k]

    def eat(self, tok):
        self.ws()
        if self.s.startswith(tok, self.i):
            self.i += len(tok)
            return True
        return False

    def expect(self, tok):
        if not self.eat(tok):
            raise ParseError("expected %r at %d in %r" % (tok, self.i, self.s))

    def word(self):
        self.ws()
        j = self.i
        while j < len(self.s) and (self.s[j].isalnum() or self.s[j] == "_"):
            j += 1
        return self.s[self.i:j]

    def eat_word(self, w):
        if self.word().lower() == w:
            self.i += len(w)
            return True
        return False

    def backtick(self):
        self.expect("`")
        j = self.s.index("`", self.i)
        name = self.s[self.i:j]
        self.i = j + 1
        if self.i < len(self.s) and self.s[self.i] == "i" and not (
                self.i + 1 < len(self.s) and (self.s[self.i + 1].isalnum() or self.s[self.i + 1] in "_`")):
            raise ParseError("case-insensitive identifiers are not eligible")
        return name

    def expr(self):
        self.ws()
        c = self.peek()
        if c == "(":
            self.i += 1
            e = self.inner()
            self.expect(")")
            return e
        if c == "`":
            return Identifier(self.backtick())
        if c == "[":
            self.i += 1
            ops = []
            if not self.eat("]"):
                while True:
                    ops.append(self.expr())
                    if self.eat("]"):
                        break
                    self.expect(",")
            if all(isinstance(o, Constant) for o in ops):
                return Constant([o.v for o in ops])
            return ArrayConstruct(ops)
        if c == '"':
            v, j = json.JSONDecoder().raw_decode(self.s, self.i)
            self.i = j
            return Constant(v)
        if c.isdigit() or (c == "-" and self.s[self.i + 1:self.i + 2].isdigit()):
            j = self.i + 1
            while j < len(self.s) and (self.s[j].isdigit() or self.s[j] in ".eE" or
                                       (self.s[j] in "+-" and self.s[j - 1] in "eE")):
                j += 1
            txt = self.s[self.i:j]
            self.i = j
            if any(ch in txt for ch in ".eE"):
                return Constant(new_value(float(txt)))
            return Constant(new_value(int(txt)))
        w = self.word().lower()
        if not w:
            raise ParseError("unexpected %r at %d in %r" % (c, self.i, self.s))
        self.i += len(w)
        if w == "true":
            return Constant(True)
        if w == "false":
            return Constant(False)
        if w == "null":
            return Constant(None)
        if w == "missing":
            return Constant(MISSING)
        if w == "round" and self.eat("("):
            ops = [self.expr()]
            if self.eat(","):
                ops.append(self.expr())
            self.expect(")")
            return Round(*ops)
        if self.eat("("):
            distinct = False
            if self.eat_word("distinct"):
                distinct = True
            if self.eat("*"):
                operand = None
            else:
                operand = self.expr()
            self.expect(")")
            cls = _AGGS.get((w, distinct))
            if cls is None:
                raise ParseError("function %s is not on the path" % w)
            if operand is None and cls is not Count:
                raise ParseError("%s(*)" % w)
            return cls(operand)
        raise ParseError("unexpected word %r in %r" % (w, self.s))

    def inner(self):
        if self.peek() == "-" and not self.s[self.i + 1:self.i + 2].isdigit():
            self.i += 1
            return Neg(self.expr())
        if self.word().lower() == "not":
            self.i += 3
            return Not(self.expr())
        first = self.expr()
        if self.peek() == ")":
            return first
        if self.eat("."):
            return Field(first, self.backtick())
        ops = [first]
        sym = None
        while self.peek() != ")":
            s = None
            for cand in ("<=", ">=", "!=", "<>", "==", "=", "<", ">", "+", "-", "*", "/", "%"):
                if self.eat(cand):
                    s = cand
                    break
            if s is None:
                w = self.word().lower()
                if w in ("and", "or", "in", "between", "is"):
                    self.i += len(w)
                    s = w
                elif w == "not":
                    self.i += 3
                    if self.eat_word("between"):
                        s = "not between"
                    elif self.eat_word("in"):
                        s = "not in"
                    else:
                        raise ParseError("not ? in %r" % self.s)
                else:
                    raise ParseError("operator expected at %d in %r" % (self.i, self.s))
            if s == "is":
                neg = self.eat_word("not")
                what = self.word().lower()
                self.i += len(what)
                cls = {("null", False): IsNull, ("null", True): IsNotNull, ("missing", False): IsMissing,
                       ("missing", True): IsNotMissing, ("valued", False): IsValued,
                       ("valued", True): IsNotValued}[(what, neg)]
                return cls(first)
            if s in ("between", "not between"):
                lo = self.expr()
                if not self.eat_word("and"):
                    raise ParseError("between ... and")
                hi = self.expr()
                e = Between(first, lo, hi)
                return Not(e) if s.startswith("not") else e
            if sym is not None and s != sym:
                raise ParseError("mixed operators %s %s in %r" % (sym, s, self.s))
            sym = s
            ops.append(self.expr())
        a = ops
        if sym in ("+", "*", "and", "or"):
            return {"+": Add, "*": Mult, "and": And, "or": Or}[sym](*a)
        if len(a) != 2:
            raise ParseError("binary operator %s with %d operands" % (sym, len(a)))
        if sym in ("=", "=="):
            return Eq(a[0], a[1])
        if sym in ("!=", "<>"):
            return Not(Eq(a[0], a[1]))  # comp_eq.go:92-94
        if sym == "<":
            return LT(a[0], a[1])
        if sym == "<=":
            return LE(a[0], a[1])
        if sym == ">":
            return LT(a[1], a[0])  # comp_gt.go:15-17
        if sym == ">=":
            return LE(a[1], a[0])  # comp_ge.go:15-17
        if sym == "-":
            return Sub(a[0], a[1])
        if sym == "/":
            return Div(a[0], a[1])
        if sym == "%":
            return Mod(a[0], a[1])
        if sym == "in":
            return In(a[0], a[1])
        if sym == "not in":
            return Not(In(a[0], a[1]))
        raise ParseError(sym)


def parse(text: str) -> Expr:
    p = _P(text)
    e = p.expr()
    p.ws()
    if p.i != len(p.s):
        raise ParseError("trailing text at %d in %r" % (p.i, text))
    return e


class Round(Expr):
    """round(x [, digits]) - the scalar function of the reference's aggregate goldens (ROUND(AVG(unitPrice), 5)); only the
    operators behind FinalGroup evaluate it.  expression/func_num.go:1304-1336 (Apply), :1715-1736 (roundFloat: half to
    even on the scaled value), value.NewValue (integral -> int)."""

    def __init__(self, *ops):
        self.ops = list(ops)

    def children(self):
        return list(self.ops)

    def evaluate(self, item):
        a = self.ops[0].evaluate(item)
        if a is MISSING:
            return MISSING
        if vtype(a) != T_NUMBER:
            return None
        prec = 0
        if len(self.ops) > 1:
            p = self.ops[1].evaluate(item)
            if p is MISSING:
                return MISSING
            if vtype(p) != T_NUMBER:
                return None
            pf = float(p)
            if math.isnan(pf) or (not math.isinf(pf) and pf != math.trunc(pf)):  # Go: pf != math.Trunc(pf)
                return None
            prec = _go_int64_of_float(pf)  # Go: p = int(pf)
        x = float(a)
        if math.isnan(x) or math.isinf(x):
            return new_value(x)
        sign = 1.0
        if x < 0:
            sign, x = -1.0, -x
        # IEEE arithmetic as Go's math package does it (Python raises where Go yields Inf / NaN)
        try:
            pw = math.pow(10.0, float(prec))
        except OverflowError:
            pw = math.inf
        inter = x * pw + 0.5
        r = inter if (math.isinf(inter) or math.isnan(inter)) else float(math.floor(inter))
        odd = True if (math.isinf(r) or math.isnan(r)) else math.fmod(r, 2) != 0  # Mod(Inf, 2) = NaN, and NaN != 0
        if r == inter and odd:
            r -= 1
        try:
            return new_value(sign * r / pw)
        except ZeroDivisionError:
            return new_value(math.nan if r == 0 or math.isnan(r) else math.copysign(math.inf, sign * r))

    def __str__(self):  # stringer.go:581-604
        return "round(" + ", ".join(str(o) for o in self.ops) + ")"


# --------------------------------------------------------------------------------------
# The operator chain
# --------------------------------------------------------------------------------------


def group_key(item, keys):
    """execution/group_util.go:18-35: MISSING components are omitted (so distinct from NULL);
    numbers by canonical text; any injective encoding of the marshalled object is equivalent."""
    parts = []
    for i, k in enumerate(keys):
        v = k.evaluate(item)
        if v is not MISSING:
            parts.append((i, marshal(v)))
    return tuple(parts)


AGGREGATES = "\x00aggregates"  # the "aggregates" attachment of a group item (never a valid field name of an item)


class GroupRow:
    __slots__ = ("keys", "aggregates", "item")

    def __init__(self, keys, aggregates, item=None):
        # item: the group's representative, the first item that reached InitialGroup (group_initial.go:82-88)
        self.keys, self.aggregates, self.item = keys, aggregates, item

    def __repr__(self):
        return "GroupRow(%r, %r)" % (self.keys, self.aggregates)


def run_chain(docs, alias, where, group_keys, aggregates, streams=1):
    """docs: iterable of parsed documents in primary-key order (datastore/file/file.go:711-730).
    where / group_keys / aggregates: Expr objects (or Stringer text).  Returns list[GroupRow].

    Fetch (execution/fetch.go:145): item = {alias: doc}.  Filter (execution/filter.go:49-61).
    InitialGroup (group_initial.go:56-108) per stream; IntermediateGroup (group_intermediate.go:
    56-104); FinalGroup (group_final.go:55-118)."""
    where = parse(where) if isinstance(where, str) else where
    group_keys = [parse(k) if isinstance(k, str) else k for k in group_keys]
    aggregates = [parse(a) if isinstance(a, str) else a for a in aggregates]
    # planner sorts + dedups aggregates by string form (planner/build_select_sub.go:286,551-558)
    aggs = {}
    for a in aggregates:
        aggs.setdefault(str(a), a)
    agg_list = [aggs[k] for k in sorted(aggs)]

    initial = [dict() for _ in range(streams)]
    for n, doc in enumerate(docs):
        item = {alias: doc}
        if where is not None and not truth(where.evaluate(item)):
            continue
        groups = initial[n % streams]
        gk = group_key(item, group_keys) if group_keys else ()
        g = groups.get(gk)
        if g is None:
            g = groups[gk] = GroupRow([k.evaluate(item) for k in group_keys],
                                      {str(a): a.default() for a in agg_list}, item)
        for a in agg_list:
            g.aggregates[str(a)] = a.initial(item, g.aggregates[str(a)])

    inter = {}
    for groups in initial:
        for gk, g in groups.items():
            c = inter.get(gk)
            if c is None:
                inter[gk] = g
            else:
                for a in agg_list:
                    c.aggregates[str(a)] = a.intermediate(g.aggregates[str(a)], c.aggregates[str(a)])

    out = []
    for gk, g in inter.items():
        for a in agg_list:
            g.aggregates[str(a)] = a.final(g.aggregates[str(a)])
        out.append(g)
    if not group_keys and not out:  # group_final.go:108-117
        out.append(GroupRow([], {str(a): a.default() for a in agg_list}, {}))
    return out


def run_distinct(docs, alias, where, terms):
    """SELECT DISTINCT <terms> FROM ks [WHERE]: Fetch -> Filter -> InitialProject -> Distinct -> FinalProject
    (planner/build_select_sub.go:217-243).  execution/distinct.go:60-72 keeps the first item per projection object
    (value.Set: numbers by canonical value, set.go:83-99).  terms: [(expr, as or None)].  Returns the projection rows in
    first-appearance order (the reference's order is that of its parallel streams)."""
    P = lambda e: parse(e) if isinstance(e, str) else e
    where = P(where) if where is not None else None
    terms = [(P(e), a or "") for e, a in terms]
    names, n = [], 1
    for e, a in terms:
        al = a or expr_alias(e)
        if not al:
            al, n = "$%d" % n, n + 1
        names.append(al)
    seen, out = set(), []
    for doc in docs:
        item = {alias: doc}
        if where is not None and not truth(where.evaluate(item)):
            continue
        proj = {}
        for (e, _a), al in zip(terms, names):
            v = e.evaluate(item)
            if v is MISSING:
                proj.pop(al, None)
            else:
                proj[al] = v
        k = tuple(sorted((name, marshal(v)) for name, v in proj.items()))
        if k not in seen:
            seen.add(k)
            out.append({name: to_python(v) for name, v in proj.items()})
    return out


def expr_alias(e):
    """Expression.Alias(): nav_field.go:55-57,260-262 (last field name), identifier.go:66-68, else "" (base.go:163-165)."""
    if isinstance(e, (Field, Identifier)):
        return e.name
    return ""


def run_tail(groups, letting=(), having=None, terms=None, order=(), offset=None, limit=None, with_sort_keys=False):
    """The operators behind FinalGroup (planner/build_select_sub.go:217-235,276-296; build_select.go:75-110), one item
    at a time over run_chain's GroupRows:
      Let      execution/let.go:50-62          every binding evaluated on the incoming item, set on a copy
      Filter   execution/filter.go:49-61       HAVING: forward when Truth()
      InitialProject  execution/project_initial.go:98-144   projection object by alias (algebra/result.go:358-374:
                      AS, else the expression's Alias(), else $1, $2 ...); a MISSING value leaves the field out
                      (value/object.go:246-255); explicit AS names are also set on the scope the sort sees
      Order    execution/order.go:119-166      Collate per sort term, descending flips it (sort.Sort: ties unordered)
      Offset / Limit  execution/offset.go:53-83, limit.go:53-85 (operands: numbers equal to their Trunc)
      FinalProject    execution/project_final.go:51-59      the projection object is the row
    letting: [(variable, expr)], terms: [(expr, as or None)], order: [(expr, descending)].  Expressions are Expr
    objects or Stringer text.  Returns the list of row dicts (python values); with_sort_keys=True returns
    (rows, sort-key tuples) so that a caller can compare modulo the order of ties."""
    P = lambda e: parse(e) if isinstance(e, str) else e
    letting = [(v, P(e)) for v, e in letting]
    having = P(having) if having is not None else None
    terms = [(P(e), a or "") for e, a in (terms or [])]
    order = [(P(e), bool(d)) for e, d in order]
    aliases, n = [], 1
    for e, a in terms:
        al = a or expr_alias(e)
        if not al:
            al, n = "$%d" % n, n + 1
        aliases.append(al)
    rows = []
    for g in groups:
        item = dict(g.item or {})
        item[AGGREGATES] = g.aggregates
        lv = dict(item)
        for var, e in letting:
            v = e.evaluate(item)
            if v is MISSING:
                lv.pop(var, None)
            else:
                lv[var] = v
        if having is not None and not truth(having.evaluate(lv)):
            continue
        proj, scope = {}, dict(lv)
        for (e, a), al in zip(terms, aliases):
            v = e.evaluate(lv)
            if v is MISSING:
                proj.pop(al, None)
            else:
                proj[al] = v
            if a:
                if v is MISSING:
                    scope.pop(a, None)  # ScopeValue.SetField(MISSING) unsets its own field: the parent's shows through
                    if a in lv:
                        scope[a] = lv[a]
                else:
                    scope[a] = v
        keys = tuple(e.evaluate(scope) for e, _d in order)
        rows.append((proj, keys))
    if order:
        import functools

        def cmp(x, y):
            for (e, desc), a, b in zip(order, x[1], y[1]):
                c = collate(a, b)
                if c:
                    return -c if desc else c
            return 0
        rows.sort(key=functools.cmp_to_key(cmp))

    def operand(e, what):
        v = P(e).evaluate({}) if not isinstance(e, (int, float)) else e
        if vtype(v) != T_NUMBER or math.trunc(v) != v:
            raise ValueError("Invalid %s value %r." % (what, v))
        return int(v)
    if offset is not None:
        rows = rows[max(operand(offset, "OFFSET"), 0):]
    if limit is not None:
        rows = rows[:max(operand(limit, "LIMIT"), 0)]
    out = [{k: to_python(v) for k, v in p.items()} for p, _k in rows]
    if with_sort_keys:
        return out, [tuple("\x00MISSING" if v is MISSING else to_python(v) for v in k) for _p, k in rows]
    return out


def rows_as_python(rows):
    """Canonical comparable form: {group-key tuple: {agg string: python value}}; a MISSING key
    component is the string 'missing' marker object."""
    out = {}
    for g in rows:
        k = tuple("\x00MISSING" if v is MISSING else json.dumps(to_python(v), sort_keys=True) for v in g.keys)
        out[k] = {a: to_python(v) for a, v in g.aggregates.items()}
    return out


def float_bits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", x))[0]
