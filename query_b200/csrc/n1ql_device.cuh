// n1ql_device.cuh — hand-written sm_100a device library for the N1QL filter + GROUP BY path.
//
// Every specialised scan kernel (one per compiled query, emitted by codegen.cpp and built with
// NVRTC for sm_100a) is assembled from the functions in this file: typed-value semantics
// (MISSING/NULL/BOOLEAN/NUMBER/STRING collation, 4-valued logic, int64->float64 promotion),
// 128/256-bit coalesced column loads, warp-ballot selection, block reductions, shared-memory
// dense group tables and the HBM open-addressing table (64- and 128-bit keys).
//
// The same header is compiled by nvcc into libn1gpu.so's static kernels (kernels.cu), so it is
// build-checked for sm_100a at build() time.  It must stay free of host/std includes (NVRTC).
//
// Reference semantics restated here (file:line under /root/reference):
//   value/value.go:69-79 (type order)        value/integer.go:100-130,266-348 (int collate/arith)
//   value/float.go:106-172,331-381            value/string.go:116-142  value/boolean.go:100-114
//   expression/comp_lt.go:57-65 comp_le.go:57-65 comp_eq.go:76-78 comp_between.go:58-78
//   expression/coll_in.go:61-91 logic_and.go:64-88 logic_or.go:98-122 logic_not.go:57-68
//   expression/arith_add.go:51-70 arith_mult.go:51-70 arith_sub.go:53-61 arith_div.go:46-64
//   expression/arith_mod.go:48-66 arith_neg.go:51-59 comp_null.go comp_missing.go comp_valued.go
#pragma once

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

// value classes: the per-row tag byte of a column.  Ordered so that the N1QL type order
// MISSING < NULL < BOOLEAN(false<true) < NUMBER < STRING is monotone in the class.
#define C_MISSING 0
#define C_NULL 1
#define C_FALSE 2
#define C_TRUE 3
#define C_INT 4
#define C_FLOAT 5
#define C_STRING 6
#define C_OTHER 7 /* array / object / binary: never reaches a kernel (plan is ineligible) */

#define NQ_I64_MIN ((i64)0x8000000000000000LL)
#define NQ_I64_MAX ((i64)0x7fffffffffffffffLL)
#define NQ_U64_MAX (~(u64)0)

// accumulator word operations (also the merge operation of IntermediateGroup)
#define OP_ADD_U64 0
#define OP_ADD_F64 1
#define OP_MIN_I64 2
#define OP_MAX_I64 3
#define OP_MIN_U64 4
#define OP_MAX_U64 5
#define OP_OR_U64 6

struct Val {
    int c;  // class
    i64 b;  // payload: int64 | float64 bits | 2*dictionary-rank (string; odd = absent constant)
};

#define NQ_DEV __device__ __forceinline__

NQ_DEV Val mkv(int c, i64 b) { Val v; v.c = c; v.b = b; return v; }
NQ_DEV Val mkbool(bool t) { return mkv(t ? C_TRUE : C_FALSE, 0); }
NQ_DEV Val mkint(i64 x) { return mkv(C_INT, x); }
NQ_DEV Val mkflt(double d) { return mkv(C_FLOAT, __double_as_longlong(d)); }
NQ_DEV double as_f(i64 b) { return __longlong_as_double(b); }
NQ_DEV bool is_num(int c) { return c == C_INT || c == C_FLOAT; }
NQ_DEV int type_rank(int c) { return c <= C_NULL ? c : (c <= C_TRUE ? 2 : (c <= C_FLOAT ? 3 : 4)); }
NQ_DEV double num_f(Val v) { return v.c == C_INT ? (double)v.b : as_f(v.b); }

// Go's int64(float64) on amd64 (out of range -> MinInt64) and value.IsInt (integer.go:354-356)
NQ_DEV i64 go_i64(double d) {
    if (!(d >= -9223372036854775808.0 && d < 9223372036854775808.0)) return NQ_I64_MIN;
    return (i64)d;
}
NQ_DEV bool f_is_int(double d) { return d == (double)go_i64(d); }
// value.NewValue(float64): integral -> intValue (value/value.go:377-382)
NQ_DEV Val new_num(double d) { return f_is_int(d) ? mkint(go_i64(d)) : mkflt(d); }
// canonical number identity used by group keys / DISTINCT / MIN / MAX: an integral floatValue
// (only arithmetic can make one) is the equal int (group_util.go:18-35 text form, set.go:83-99)
NQ_DEV Val canon_num(Val v) {
    if (v.c == C_FLOAT) { double d = as_f(v.b); if (f_is_int(d)) return mkint(go_i64(d)); }
    return v;
}

// ---- collation -------------------------------------------------------------------------------
NQ_DEV int collate_f(double t, double o) {  // value/float.go:123-172
    bool tn = t != t, on = o != o;
    if (tn) return on ? 0 : -1;
    if (on) return 1;
    return t < o ? -1 : (t > o ? 1 : 0);  // +-Inf order falls out of IEEE compare
}
// same-or-different type collation sign for values > NULL
NQ_DEV int collate(Val a, Val b) {
    int ra = type_rank(a.c), rb = type_rank(b.c);
    if (ra != rb) return ra - rb;
    if (ra == 3) {
        if (a.c == C_INT && b.c == C_INT) return a.b < b.b ? -1 : (a.b > b.b ? 1 : 0);
        return collate_f(num_f(a), num_f(b));
    }
    if (ra == 2) return a.c - b.c;
    if (ra == 4) return a.b < b.b ? -1 : (a.b > b.b ? 1 : 0);
    return 0;
}
#define CMP_NULL 8
#define CMP_MISSING 9
NQ_DEV int v_compare(Val a, Val b) {  // Value.Compare
    if (a.c == C_MISSING || b.c == C_MISSING) return CMP_MISSING;
    if (a.c == C_NULL || b.c == C_NULL) return CMP_NULL;
    int c = collate(a, b);
    return c < 0 ? -1 : (c > 0 ? 1 : 0);
}
NQ_DEV Val cmp_result(int cmp, bool t) { return cmp == CMP_MISSING ? mkv(C_MISSING, 0) : (cmp == CMP_NULL ? mkv(C_NULL, 0) : mkbool(t)); }
NQ_DEV Val v_lt(Val a, Val b) { int c = v_compare(a, b); return cmp_result(c, c < 0); }
NQ_DEV Val v_le(Val a, Val b) { int c = v_compare(a, b); return cmp_result(c, c <= 0); }
NQ_DEV Val v_eq(Val a, Val b) {  // Value.Equals
    if (a.c == C_MISSING || b.c == C_MISSING) return mkv(C_MISSING, 0);
    if (a.c == C_NULL || b.c == C_NULL) return mkv(C_NULL, 0);
    int ra = type_rank(a.c), rb = type_rank(b.c);
    if (ra != rb) return mkbool(false);
    if (ra == 3) {
        if (a.c == C_INT && b.c == C_INT) return mkbool(a.b == b.b);
        return mkbool(num_f(a) == num_f(b));
    }
    if (ra == 2) return mkbool(a.c == b.c);
    return mkbool(a.b == b.b);
}
NQ_DEV Val v_between(Val x, Val lo, Val hi) {  // comp_between.go:58-78
    int lc = v_compare(x, lo);
    if (lc == CMP_MISSING) return mkv(C_MISSING, 0);
    int hc = v_compare(x, hi);
    if (hc == CMP_MISSING) return mkv(C_MISSING, 0);
    if (lc == CMP_NULL || hc == CMP_NULL) return mkv(C_NULL, 0);
    return mkbool(lc >= 0 && hc <= 0);
}
// Value.Truth; empty2 = 2*rank of "" in the string operand's dictionary (or -1)
NQ_DEV bool v_truth(Val v, i64 empty2) {
    if (v.c == C_TRUE) return true;
    if (v.c == C_INT) return v.b != 0;
    if (v.c == C_FLOAT) { double d = as_f(v.b); return d == d && d != 0.0; }
    if (v.c == C_STRING) return v.b != empty2;
    return false;
}
NQ_DEV Val v_not(Val a, i64 empty2) {
    if (a.c <= C_NULL) return a;
    return mkbool(!v_truth(a, empty2));
}
// n-ary AND / OR folded one argument at a time: flags f(alse seen) / t(rue seen), m(issing), n(ull)
NQ_DEV void and_arg(Val a, i64 empty2, bool& f, bool& m, bool& n) {
    if (a.c == C_NULL) n = true; else if (a.c == C_MISSING) m = true; else if (!v_truth(a, empty2)) f = true;
}
NQ_DEV Val and_fin(bool f, bool m, bool n) { return f ? mkbool(false) : (m ? mkv(C_MISSING, 0) : (n ? mkv(C_NULL, 0) : mkbool(true))); }
NQ_DEV void or_arg(Val a, i64 empty2, bool& t, bool& m, bool& n) {
    if (a.c == C_NULL) n = true; else if (a.c == C_MISSING) m = true; else if (v_truth(a, empty2)) t = true;
}
NQ_DEV Val or_fin(bool t, bool m, bool n) { return t ? mkbool(true) : (n ? mkv(C_NULL, 0) : (m ? mkv(C_MISSING, 0) : mkbool(false))); }
NQ_DEV Val v_is_null(Val a) { return a.c == C_NULL ? mkbool(true) : (a.c == C_MISSING ? a : mkbool(false)); }
NQ_DEV Val v_is_not_null(Val a) { return a.c == C_NULL ? mkbool(false) : (a.c == C_MISSING ? a : mkbool(true)); }
NQ_DEV Val v_is_missing(Val a) { return mkbool(a.c == C_MISSING); }
NQ_DEV Val v_is_not_missing(Val a) { return mkbool(a.c != C_MISSING); }
NQ_DEV Val v_is_valued(Val a) { return mkbool(a.c > C_NULL); }
NQ_DEV Val v_is_not_valued(Val a) { return mkbool(a.c <= C_NULL); }
// one element of `x IN [..]` (coll_in.go:69-82)
NQ_DEV void in_arg(Val x, Val e, bool& hit, bool& m, bool& n) {
    if (x.c > C_NULL && e.c > C_NULL) { if (v_eq(x, e).c == C_TRUE) hit = true; }
    else if (e.c == C_MISSING) m = true;
    else n = true;
}
NQ_DEV Val in_fin(Val x, bool hit, bool m, bool n) {
    if (x.c == C_MISSING) return x;
    return hit ? mkbool(true) : (n ? mkv(C_NULL, 0) : (m ? mkv(C_MISSING, 0) : mkbool(false)));
}

// ---- NumberValue arithmetic ----------------------------------------------------------------------
NQ_DEV Val num_add(Val a, Val b) {  // integer.go:266-277, float.go:331-333
    if (a.c == C_INT && b.c == C_INT) {
        i64 rv = (i64)((u64)a.b + (u64)b.b);
        if ((a.b >= 0 && b.b >= 0 && rv >= 0) || (a.b < 0 && b.b < 0 && rv < 0)) return mkint(rv);
    }
    return mkflt(num_f(a) + num_f(b));
}
NQ_DEV Val num_mult(Val a, Val b) {  // integer.go:319-329 (rv/this == n  <=>  no overflow, bar MinInt64 * -1)
    if (a.c == C_INT && b.c == C_INT) {
        i64 lo = (i64)((u64)a.b * (u64)b.b);
        i64 hi = __mul64hi(a.b, b.b);
        bool ok = (hi == (lo >> 63)) || (a.b == -1 && b.b == NQ_I64_MIN);
        if (a.b == NQ_I64_MIN && b.b == -1) ok = false;
        if (ok) return mkint(lo);
    }
    return mkflt(num_f(a) * num_f(b));
}
NQ_DEV Val num_neg(Val a) {  // integer.go:331-337
    if (a.c == C_INT) { if (a.b == NQ_I64_MIN) return mkflt(-(double)a.b); return mkint(-a.b); }
    return mkflt(-as_f(a.b));
}
NQ_DEV Val num_sub(Val a, Val b) {  // integer.go:339-348, float.go:375-377
    if (a.c == C_INT && b.c == C_INT && b.b > NQ_I64_MIN) return num_add(a, mkint(-b.b));
    return mkflt(num_f(a) - num_f(b));
}
// n-ary + and * fold (arith_add.go:51-70, arith_mult.go:51-70)
NQ_DEV void add_arg(Val a, Val& acc, bool& m, bool& n) {
    if (!n && is_num(a.c)) acc = num_add(acc, a); else if (a.c == C_MISSING) m = true; else n = true;
}
NQ_DEV void mult_arg(Val a, Val& acc, bool& m, bool& n) {
    if (!n && is_num(a.c)) acc = num_mult(acc, a); else if (a.c == C_MISSING) m = true; else n = true;
}
NQ_DEV Val arith_fin(Val acc, bool m, bool n) { return m ? mkv(C_MISSING, 0) : (n ? mkv(C_NULL, 0) : acc); }
NQ_DEV Val v_sub(Val a, Val b) {
    if (is_num(a.c) && is_num(b.c)) return num_sub(a, b);
    if (a.c == C_MISSING || b.c == C_MISSING) return mkv(C_MISSING, 0);
    return mkv(C_NULL, 0);
}
NQ_DEV Val v_neg(Val a) { return is_num(a.c) ? num_neg(a) : (a.c == C_MISSING ? a : mkv(C_NULL, 0)); }
NQ_DEV Val v_div(Val a, Val b) {  // arith_div.go:46-64
    if (a.c == C_MISSING || b.c == C_MISSING) return mkv(C_MISSING, 0);
    if (is_num(b.c)) {
        double s = num_f(b);
        if (s == 0.0) return mkv(C_NULL, 0);
        if (is_num(a.c)) return new_num(num_f(a) / s);
    }
    return mkv(C_NULL, 0);
}
NQ_DEV Val v_mod(Val a, Val b) {  // arith_mod.go:48-66 (math.Mod == C fmod)
    if (a.c == C_MISSING || b.c == C_MISSING) return mkv(C_MISSING, 0);
    if (is_num(b.c)) {
        double s = num_f(b);
        if (s == 0.0) return mkv(C_NULL, 0);
        if (is_num(a.c)) return new_num(fmod(num_f(a), s));
    }
    return mkv(C_NULL, 0);
}

// order-preserving u64 image of a float64 (for MIN/MAX through integer atomics)
NQ_DEV u64 f64_ordered(double d) { u64 u = (u64)__double_as_longlong(d); return (u >> 63) ? ~u : (u | 0x8000000000000000ULL); }
NQ_DEV double f64_unordered(u64 k) { u64 u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k; return __longlong_as_double((i64)u); }

// ---- coalesced vector loads (4 rows per thread: 256-bit / 128-bit / 32-bit) -----------------------
// ld.global.nc + L1::no_allocate: every column byte is read exactly once per scan.
NQ_DEV void ld_rows4_b64(const i64* __restrict__ p, i64 (&v)[4]) {
#if __CUDA_ARCH__ >= 1000 && !defined(NQ_NO_LD256)
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(v[2]), "=l"(v[3]) : "l"(p + 2));
#endif
}
NQ_DEV void ld_rows4_b32(const u32* __restrict__ p, u32 (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p));
}
NQ_DEV void ld_rows4_b8(const u8* __restrict__ p, int (&v)[4]) {
    u32 w;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(w) : "l"(p));
    v[0] = w & 0xff; v[1] = (w >> 8) & 0xff; v[2] = (w >> 16) & 0xff; v[3] = w >> 24;
}

// ---- warp / block reductions ---------------------------------------------------------------------
NQ_DEV u64 shfl_xor_u64(u64 v, int m) { return (u64)__shfl_xor_sync(0xffffffffu, (i64)v, m); }
NQ_DEV u64 word_combine(int op, u64 a, u64 b) {
    switch (op) {
        case OP_ADD_U64: return a + b;
        case OP_ADD_F64: return (u64)__double_as_longlong(__longlong_as_double((i64)a) + __longlong_as_double((i64)b));
        case OP_MIN_I64: return (i64)a < (i64)b ? a : b;
        case OP_MAX_I64: return (i64)a > (i64)b ? a : b;
        case OP_MIN_U64: return a < b ? a : b;
        case OP_MAX_U64: return a > b ? a : b;
        default: return a | b;
    }
}
NQ_DEV u64 word_identity(int op) {
    switch (op) {
        case OP_MIN_I64: return (u64)NQ_I64_MAX;
        case OP_MAX_I64: return (u64)NQ_I64_MIN;
        case OP_MIN_U64: return NQ_U64_MAX;
        default: return 0;  // ADD_U64, ADD_F64 (+0.0), MAX_U64, OR
    }
}
template <int OP> NQ_DEV u64 warp_reduce_word(u64 v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = word_combine(OP, v, shfl_xor_u64(v, m));
    return v;
}
// Block reduce of one accumulator word in a fixed (deterministic) order; result valid in thread 0.
// scratch: u64[32] in shared memory.
template <int OP> NQ_DEV u64 block_reduce_word(u64 v, u64* scratch) {
    v = warp_reduce_word<OP>(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < nw ? scratch[lane] : word_identity(OP);
        v = warp_reduce_word<OP>(v);
    }
    return v;
}

// ---- atomic application of an accumulator word (shared or global memory) ---------------------------
template <int OP> NQ_DEV void atomic_word(u64* p, u64 x) {
    if (OP == OP_ADD_U64) atomicAdd(p, x);
    else if (OP == OP_ADD_F64) atomicAdd((double*)p, __longlong_as_double((i64)x));
    else if (OP == OP_MIN_I64) atomicMin((i64*)p, (i64)x);
    else if (OP == OP_MAX_I64) atomicMax((i64*)p, (i64)x);
    else if (OP == OP_MIN_U64) atomicMin(p, x);
    else if (OP == OP_MAX_U64) atomicMax(p, x);
    else atomicOr(p, x);
}
// Min / max / or into an HBM table word that is read first: an L2 load costs the SM -> L2 request path less than a
// reduction, and a settled group leaves the word unchanged for almost every row.  The load may be stale; these words
// only ever move one way, so a stale value can only cause a reduction that was not needed, never skip one that was.
template <int OP> NQ_DEV void table_word_known(u64* p, u64 x, u64 cur);
template <int OP> NQ_DEV void table_word_checked(u64* p, u64 x) { table_word_known<OP>(p, x, __ldcg(p)); }
template <int OP> NQ_DEV void table_word_known(u64* p, u64 x, u64 cur) {  // `cur`: a value the word held earlier
    if (OP == OP_MIN_I64) { if ((i64)x < (i64)cur) atomicMin((i64*)p, (i64)x); }
    else if (OP == OP_MAX_I64) { if ((i64)x > (i64)cur) atomicMax((i64*)p, (i64)x); }
    else if (OP == OP_MIN_U64) { if (x < cur) atomicMin(p, x); }
    else if (OP == OP_MAX_U64) { if (x > cur) atomicMax(p, x); }
    else if (OP == OP_OR_U64) { if ((cur & x) != x) atomicOr(p, x); }
    else atomic_word<OP>(p, x);
}
NQ_DEV void atomic_word_dyn(int op, u64* p, u64 x) {
    switch (op) {
        case OP_ADD_U64: atomic_word<OP_ADD_U64>(p, x); break;
        case OP_ADD_F64: atomic_word<OP_ADD_F64>(p, x); break;
        case OP_MIN_I64: atomic_word<OP_MIN_I64>(p, x); break;
        case OP_MAX_I64: atomic_word<OP_MAX_I64>(p, x); break;
        case OP_MIN_U64: atomic_word<OP_MIN_U64>(p, x); break;
        case OP_MAX_U64: atomic_word<OP_MAX_U64>(p, x); break;
        default: atomic_word<OP_OR_U64>(p, x); break;
    }
}

// ---- HBM open-addressing tables --------------------------------------------------------------------
// 64-bit keys: slot key array u64[cap], EMPTY = all ones (packed keys use at most 63 bits).
// 128-bit keys: ulonglong2[cap] claimed with one 128-bit CAS (ATOMG.E.CAS.128), EMPTY = all ones.
NQ_DEV u64 mix64(u64 x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31; return x;
}
// returns slot, or -1 when the probe budget is exhausted (table too small -> host grows and reruns);
// *fresh (optional) reports whether this call claimed the slot.
NQ_DEV i64 table_insert64(u64* __restrict__ keys, u64 cap_mask, u64 key, bool* fresh) {
    u64 slot = mix64(key) & cap_mask;
    for (u64 probe = 0; probe <= cap_mask; ++probe) {
        u64 cur = *(volatile u64*)&keys[slot];
        if (cur == key) { if (fresh) *fresh = false; return (i64)slot; }
        if (cur == NQ_U64_MAX) {
            u64 old = atomicCAS(&keys[slot], NQ_U64_MAX, key);
            if (old == NQ_U64_MAX) { if (fresh) *fresh = true; return (i64)slot; }
            if (old == key) { if (fresh) *fresh = false; return (i64)slot; }
        }
        slot = (slot + 1) & cap_mask;
        if (probe > 4096) break;
    }
    return -1;
}
NQ_DEV i64 table_find64(const u64* __restrict__ keys, u64 cap_mask, u64 key) {
    u64 slot = mix64(key) & cap_mask;
    for (u64 probe = 0; probe <= cap_mask; ++probe) {
        u64 cur = keys[slot];
        if (cur == key) return (i64)slot;
        if (cur == NQ_U64_MAX) return -1;
        slot = (slot + 1) & cap_mask;
    }
    return -1;
}
#if !defined(__CUDA_ARCH__) || __CUDA_ARCH__ >= 900
NQ_DEV i64 table_insert128(ulonglong2* __restrict__ keys, u64 cap_mask, u64 lo, u64 hi, bool* fresh) {
    u64 slot = mix64(lo ^ mix64(hi)) & cap_mask;
    unsigned __int128 key = ((unsigned __int128)hi << 64) | lo;
    unsigned __int128 empty = ~(unsigned __int128)0;
    for (u64 probe = 0; probe <= cap_mask; ++probe) {
        ulonglong2 cur;  // one aligned 128-bit access (LDG.E.128): never observes half a key
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(cur.x), "=l"(cur.y) : "l"(&keys[slot]) : "memory");
        if (cur.x == lo && cur.y == hi) { if (fresh) *fresh = false; return (i64)slot; }
        if (cur.x == NQ_U64_MAX && cur.y == NQ_U64_MAX) {
            unsigned __int128 old = atomicCAS((unsigned __int128*)&keys[slot], empty, key);
            if (old == empty) { if (fresh) *fresh = true; return (i64)slot; }
            if (old == key) { if (fresh) *fresh = false; return (i64)slot; }
        }
        slot = (slot + 1) & cap_mask;
        if (probe > 4096) break;
    }
    return -1;
}
NQ_DEV i64 table_find128(const ulonglong2* __restrict__ keys, u64 cap_mask, u64 lo, u64 hi) {
    u64 slot = mix64(lo ^ mix64(hi)) & cap_mask;
    for (u64 probe = 0; probe <= cap_mask; ++probe) {
        ulonglong2 cur = keys[slot];
        if (cur.x == lo && cur.y == hi) return (i64)slot;
        if (cur.x == NQ_U64_MAX && cur.y == NQ_U64_MAX) return -1;
        slot = (slot + 1) & cap_mask;
    }
    return -1;
}
#endif

// Per-block shared-memory front cache of the HBM group table (hot keys of a skewed GROUP BY are aggregated with
// shared-memory atomics and reach HBM once per block): claims or finds `key` within 4 probes, else -1 (miss).
// (any slot count, not only powers of two: the slot is the high part of hash * slots)
NQ_DEV int cache_claim_n(u64* ckeys, u32 slots, u32 hash, u64 key) {
    u32 h = __umulhi(hash, slots);
#pragma unroll
    for (int probe = 0; probe < 4; ++probe) {
        const u64 cur = ((volatile u64*)ckeys)[h];
        if (cur == key) return (int)h;
        if (cur == NQ_U64_MAX) {
            const u64 old = atomicCAS(&ckeys[h], NQ_U64_MAX, key);
            if (old == NQ_U64_MAX || old == key) return (int)h;
        }
        h = h + 1 == slots ? 0 : h + 1;
    }
    return -1;
}

// The same for packed keys of <= 31 bits: u32 keys in 16-byte buckets of four.  One 128-bit shared-memory load
// fetches the whole bucket, so a lookup costs one instruction and one latency where the linear probe above pays up to
// four dependent 64-bit loads - once the cache is full, every row of a key that is not cached walked all four.
// Slots only ever change from empty to a key and every thread tries the slots of a bucket in the same order, so a key
// settles in exactly one slot (and the block-end flush is additive, so even a duplicate would be harmless).
NQ_DEV int cache_claim_b4(u32* ckeys, u32 nbuckets, u32 hash, u32 key) {
    const u32 b = __umulhi(hash, nbuckets) * 4u;
    u32 k[4];
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(k[0]), "=r"(k[1]), "=r"(k[2]), "=r"(k[3]) : "r"((u32)__cvta_generic_to_shared(ckeys + b)) : "memory");
    int s = -1;  // select chain, not branches: the lanes of a warp match in different slots
    s = k[3] == key ? (int)b + 3 : s;
    s = k[2] == key ? (int)b + 2 : s;
    s = k[1] == key ? (int)b + 1 : s;
    s = k[0] == key ? (int)b : s;
    if (s >= 0 || k[3] != 0xffffffffu) return s;  // found, or full (slots fill in order: the last one taken = full)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (k[i] != 0xffffffffu) continue;
        const u32 old = atomicCAS(&ckeys[b + i], 0xffffffffu, key);
        if (old == 0xffffffffu || old == key) return (int)b + i;
    }
    return -1;
}

// Direct-mapped variant: one slot per key, one 32-bit load.  Measured on config 5 (tools/proto5.cu, 1 B rows): the bucket
// probe above costs a 16-byte load per lane (4 shared-memory wavefronts at best) and a 4-way select chain for every row;
// the single probe misses a few more keys and is still 6 % faster overall.
NQ_DEV int cache_claim_1(u32* ckeys, u32 nslots, u32 hash, u32 key) {
    const u32 b = __umulhi(hash, nslots);
    const u32 k = *(volatile u32*)&ckeys[b];
    if (k == key) return (int)b;
    if (k != 0xffffffffu) return -1;
    const u32 old = atomicCAS(&ckeys[b], 0xffffffffu, key);
    return (old == 0xffffffffu || old == key) ? (int)b : -1;
}

// Cells of the front cache.  Shared memory has native 32-bit atomics only (a 64-bit atomicAdd/Min/Max on shared
// memory compiles to a compare-and-swap loop, 3-9x slower and collapsing under same-key contention), so a cached
// accumulator word is kept in 32-bit cells wherever its per-block value provably fits or can be split:
//   row counters        one cell (a block scans far fewer than 2^32 rows)
//   integer sums        two cells, low word + high word with the carry propagated by the adding thread
//   ranged min / max    one cell holding (value - lo + 1); 0 (max) / all ones (min) = nothing seen
//   class-seen bits     one cell
// anything else (float64 sums, unranged or float min/max) keeps a 64-bit cell and checks before it swaps.
NQ_DEV void cache_add_wide(u32* lo, u32* hi, u64 x) {
    const u32 xl = (u32)x;
    const u32 old = atomicAdd(lo, xl);
    const u32 h = (u32)(x >> 32) + ((u32)(old + xl) < xl ? 1u : 0u);
    if (h) atomicAdd(hi, h);
}
// One-cell integer sum (addends within +-2^31): the low word is cached, whatever the add carries or borrows goes straight
// to the table word as a multiple of 2^32 (rare: once per 2^32 / |x| rows; a small negative addend wraps the cell and
// cancels its own sign extension).  Cell and table word are both plain sums, so the flush just adds the cell.
NQ_DEV void cache_add_carry(u32* lo, u64 x, u64* table_word) {
    const u32 xl = (u32)x;
    const u32 old = atomicAdd(lo, xl);
    const u32 h = (u32)(x >> 32) + ((u32)(old + xl) < xl ? 1u : 0u);
    if (h) atomicAdd(table_word, (u64)h << 32);
}
// per-thread register accumulators of the register groups (see codegen.cpp): ranged min / max keep the cache cells' image
template <int OP> NQ_DEV void reg_mm32(u32& r, u64 x, u64 bias) {
    const u32 v = (u32)(x - bias) + 1u;
    if (OP == OP_MIN_I64 || OP == OP_MIN_U64) r = v < r ? v : r; else r = v > r ? v : r;
}
// (read first: a shared-memory load costs about half an atomic, and a group's min / max settle after a few rows; a
// stale read can only cause an atomic that was not needed)
template <int OP> NQ_DEV void cache_mm32(u32* c, u64 x, u64 bias) {
    const u32 v = (u32)(x - bias) + 1u;
#ifdef NQ_NO_CELL_CHECK
    const u32 cur = (OP == OP_MIN_I64 || OP == OP_MIN_U64) ? 0xffffffffu : 0u;
#else
    const u32 cur = *(volatile u32*)c;
#endif
    if (OP == OP_MIN_I64 || OP == OP_MIN_U64) { if (v < cur) atomicMin(c, v); } else { if (v > cur) atomicMax(c, v); }
}
// the same with the cell's current value already in a register (two interleaved cells fetched by one 64-bit load)
template <int OP> NQ_DEV void cache_mm32_cur(u32* c, u64 x, u64 bias, u32 cur) {
    const u32 v = (u32)(x - bias) + 1u;
    if (OP == OP_MIN_I64 || OP == OP_MIN_U64) { if (v < cur) atomicMin(c, v); } else { if (v > cur) atomicMax(c, v); }
}
NQ_DEV void cache_or32(u32* c, u32 bits) {
    if ((*(volatile u32*)c & bits) != bits) atomicOr(c, bits);
}
template <int OP> NQ_DEV void cache_word64(u64* c, u64 x) {
    if (OP == OP_MIN_I64) { if ((i64)x < *(volatile i64*)c) atomicMin((i64*)c, (i64)x); }
    else if (OP == OP_MAX_I64) { if ((i64)x > *(volatile i64*)c) atomicMax((i64*)c, (i64)x); }
    else if (OP == OP_MIN_U64) { if (x < *(volatile u64*)c) atomicMin(c, x); }
    else if (OP == OP_MAX_U64) { if (x > *(volatile u64*)c) atomicMax(c, x); }
    else atomic_word<OP>(c, x);
}

// bit packing of group-key / DISTINCT-entry components into a 128-bit (lo,hi) key
NQ_DEV void pack_bits(u64& lo, u64& hi, int& pos, u64 v, int nbits) {
    if (nbits == 0) return;
    if (pos < 64) {
        lo |= v << pos;
        if (pos + nbits > 64) hi |= v >> (64 - pos);
    } else {
        hi |= v << (pos - 64);
    }
    pos += nbits;
}

// kernel parameter block shared by every generated scan kernel
struct NqParams {
    i64 nrows;             // rows in this partition
    const void* col[16];   // payload arrays (i64* or u32*), 32-byte aligned, padded to a multiple of 4096 rows
    const u8* tag[16];     // class byte per row
    u64* acc;              // UNGROUPED: per-block partials [grid][nwords]; DENSE/HASH: table words [nwords][cap]
    u64* keys;             // HASH: slot keys (u64[cap] or ulonglong2[cap])
    u64 cap_mask;          // HASH: capacity-1
    u64* set_keys;         // DISTINCT entry set (ulonglong2[set_cap] or u64[set_cap])
    u64 set_mask;
    int* status;           // [0] != 0: a table overflowed (host grows it and reruns)
    u64 dense_groups;      // DENSE: number of dense slots
    u64* final_dev;        // UNGROUPED: final words [nwords] in HBM (written by the last block to finish)
    u64* final_host;       // UNGROUPED/DENSE: the same words in mapped pinned host memory (zero-copy result)
    unsigned* ticket;      // blocks-done counter for the last-block pattern (self-resetting)
    u64* partials;         // UNGROUPED: per-block partials of the float64-sum words [nfloat][grid]
    // multi-GPU small-state merge fused into the scan: the last block pushes this rank's final words straight into
    // every peer's mailbox over NVLink (peer stores), then publishes a sequence flag (release, system scope)
    u64* const* peer_mail; // [nranks] mailbox base of every rank (peer-mapped; own entry = local pointer), or null
    int nranks, rank;
    u64 mail_base;         // word offset of (slot, this rank) inside a mailbox
    u64 mail_words;        // words pushed per step (accumulator words x slots of the dense table)
    u64 mail_seq;          // sequence number of this step (never 0)
    // DISTINCT bitmap larger than the L2 keeps: the scan runs in passes, pass k sets the bits of entries whose high
    // bits (entry >> set_shift) equal k - a slice of the bitmap that stays L2-resident - and only pass 0 feeds the
    // group table
    int set_pass, set_shift;
    // payloads of the chain's constants and bound parameters (int64 / float64 bits / 2 * dictionary rank): the kernel text
    // only fixes their CLASS, so statements that differ in a bound - or bindings of one prepared statement - share a cubin
    i64 cst[32];
    // execution.Operator.SendStop while the scan runs: a word in HBM that n1gpu_query_cancel sets with an asynchronous copy;
    // lane 0 of every warp polls it once per 32 tiles and the warp leaves the loop
    const int* cancel;
};
NQ_DEV bool nq_cancelled(const NqParams& p, int lane) {
    int c = 0;
    if (lane == 0) c = *(volatile const int*)p.cancel;
    return __shfl_sync(0xffffffffu, c, 0) != 0;
}

// Pushes `words` final words at src to every peer's mailbox and raises the flag word behind them.
NQ_DEV void mailbox_push(const NqParams& p, const u64* src) {
    __threadfence();
    __syncthreads();
    const u64 total = p.mail_words * (u64)p.nranks;
    for (u64 i = threadIdx.x; i < total; i += blockDim.x) {
        const u64 r = i / p.mail_words, w = i % p.mail_words;
        p.peer_mail[r][p.mail_base + w] = __ldcg(&src[w]);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p.nranks) {
        u64* flag = &p.peer_mail[threadIdx.x][p.mail_base + p.mail_words];
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(p.mail_seq) : "memory");
    }
}

// Dynamic-op warp reduction (the final, fixed-order fold of per-block partials by the last block).
NQ_DEV u64 warp_reduce_dyn(int op, u64 v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = word_combine(op, v, shfl_xor_u64(v, m));
    return v;
}
// Returns true in every thread of exactly one block per launch: the last one to arrive.  All global writes the
// other blocks made before their arrival are visible to it.
NQ_DEV bool last_block_arrives(unsigned* ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0;  // ready for the next launch on this stream
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}
