/*
 * oracle_ref.c — ORACLE / CPU BASELINE.  TEST INFRASTRUCTURE ONLY, never on the product path.
 *
 * A "reference-shaped" C restatement of the Go execution chain of pavel-paulau/query
 *   PrimaryScan/Fetch -> Filter -> InitialGroup -> IntermediateGroup -> FinalGroup
 * document-at-a-time over raw JSON text, the way the reference runs it:
 *   - every field reference re-scans the raw document for the first member of that name
 *     (value/parsed.go:159-207 -> go_json FirstFind),
 *   - the predicate and the aggregate operands are evaluated by a tree-walking interpreter over boxed
 *     values (expression/*.go Evaluate/Apply),
 *   - the group key is marshalled to a string and looked up in a hash map (execution/group_util.go:18-35,
 *     group_initial.go:56-100),
 *   - one InitialGroup stream per thread, merged by IntermediateGroup in stream order, finalised by
 *     FinalGroup (group_intermediate.go:56-104, group_final.go:55-118; algebra/agg_*.go).
 * It is cross-checked against the pure-Python oracle (oracle/n1ql_oracle.py, itself pinned to the
 * reference's golden vectors) in tests/test_oracle_cref.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Parity pinning: through n1ql_oracle.py.
 *
 * Expressions arrive as S-expressions produced by oracle/cref.py from the oracle's parse tree:
 *   (ci 5) (cf <hex bits>) (cs <hex utf8>) (cb 0|1) (cnull) (cmissing) (field <hex> <hex> ...)
 *   (add a b ..) (mult a b ..) (sub a b) (div a b) (mod a b) (neg a) (eq a b) (lt a b) (le a b)
 *   (between x lo hi) (in x e1 e2 ..) (and ..) (or ..) (not a) (isnull a) (isnotnull a) (ismissing a)
 *   (isnotmissing a) (isvalued a) (isnotvalued a)
 * aggregates: "<kind> <distinct 0|1> <operand sexpr | *>" with kind in count countn sum avg min max.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { T_MISSING = 0, T_NULL = 1, T_BOOL = 2, T_NUM = 3, T_STR = 4, T_OTHER = 5 };

typedef struct {
    int t;
    int isint;        /* T_NUM: int64 vs float64 */
    long long i;
    double f;
    int b;
    const char* s;    /* T_STR: unescaped bytes (arena) ; T_OTHER: raw JSON text */
    int slen;
} V;

/* ---- arena ------------------------------------------------------------------------------------------ */
typedef struct Chunk { struct Chunk* next; size_t used, cap; char data[]; } Chunk;
typedef struct { Chunk* head; } Arena;
static void* arena_alloc(Arena* a, size_t n) {
    n = (n + 7) & ~(size_t)7;
    if (!a->head || a->head->used + n > a->head->cap) {
        size_t cap = n > (1 << 20) ? n : (1 << 20);
        Chunk* c = (Chunk*)malloc(sizeof(Chunk) + cap);
        c->next = a->head; c->used = 0; c->cap = cap; a->head = c;
    }
    void* p = a->head->data + a->head->used;
    a->head->used += n;
    return p;
}
static void arena_reset(Arena* a) {  /* keep the first chunk */
    while (a->head && a->head->next) { Chunk* c = a->head; a->head = c->next; free(c); }
    if (a->head) a->head->used = 0;
}
static void arena_free(Arena* a) { while (a->head) { Chunk* c = a->head; a->head = c->next; free(c); } }

/* ---- JSON scanning -------------------------------------------------------------------------------------- */
static const char* skip_ws(const char* p, const char* e) { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; return p; }
static const char* skip_string(const char* p, const char* e) {  /* p at opening quote */
    ++p;
    while (p < e) { if (*p == '\\') { p += 2; continue; } if (*p == '"') return p + 1; ++p; }
    return NULL;
}
static const char* skip_value(const char* p, const char* e) {
    p = skip_ws(p, e);
    if (p >= e) return NULL;
    if (*p == '"') return skip_string(p, e);
    if (*p == '{' || *p == '[') {
        int depth = 0;
        while (p < e) {
            if (*p == '"') { p = skip_string(p, e); if (!p) return NULL; continue; }
            if (*p == '{' || *p == '[') ++depth;
            else if (*p == '}' || *p == ']') { if (--depth == 0) return p + 1; }
            ++p;
        }
        return NULL;
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') ++p;
    return p;
}
static int hexv(char c) { return c <= '9' ? c - '0' : ((c | 0x20) - 'a' + 10); }
static int unescape(const char* rb, const char* re, char* out) {
    char* o = out;
    for (const char* s = rb; s < re;) {
        if (*s != '\\') { *o++ = *s++; continue; }
        char c = s[1]; s += 2;
        switch (c) {
            case 'b': *o++ = '\b'; break; case 'f': *o++ = '\f'; break; case 'n': *o++ = '\n'; break;
            case 'r': *o++ = '\r'; break; case 't': *o++ = '\t'; break;
            case 'u': {
                unsigned cp = (hexv(s[0]) << 12) | (hexv(s[1]) << 8) | (hexv(s[2]) << 4) | hexv(s[3]); s += 4;
                if (cp >= 0xD800 && cp < 0xDC00 && re - s >= 6 && s[0] == '\\' && s[1] == 'u') {
                    unsigned lo = (hexv(s[2]) << 12) | (hexv(s[3]) << 8) | (hexv(s[4]) << 4) | hexv(s[5]);
                    if (lo >= 0xDC00 && lo < 0xE000) { cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); s += 6; } else cp = 0xFFFD;
                } else if (cp >= 0xD800 && cp < 0xE000) cp = 0xFFFD;
                if (cp < 0x80) *o++ = (char)cp;
                else if (cp < 0x800) { *o++ = (char)(0xC0 | (cp >> 6)); *o++ = (char)(0x80 | (cp & 0x3F)); }
                else if (cp < 0x10000) { *o++ = (char)(0xE0 | (cp >> 12)); *o++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *o++ = (char)(0x80 | (cp & 0x3F)); }
                else { *o++ = (char)(0xF0 | (cp >> 18)); *o++ = (char)(0x80 | ((cp >> 12) & 0x3F)); *o++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *o++ = (char)(0x80 | (cp & 0x3F)); }
                break;
            }
            default: *o++ = c;
        }
    }
    return (int)(o - out);
}
/* FirstFind: value text of the first member `name` of the object at [p,e); NULL if absent / not an object */
static const char* first_find(const char* p, const char* e, const char* name, int nlen, const char** vend, Arena* ar) {
    p = skip_ws(p, e);
    if (p >= e || *p != '{') return NULL;
    ++p;
    for (;;) {
        p = skip_ws(p, e);
        if (p >= e || *p != '"') return NULL;
        const char* kb = p + 1;
        const char* ke = skip_string(p, e);
        if (!ke) return NULL;
        int match;
        if (memchr(kb, '\\', (size_t)(ke - 1 - kb))) {
            char* tmp = (char*)arena_alloc(ar, (size_t)(ke - kb));
            int n = unescape(kb, ke - 1, tmp);
            match = n == nlen && memcmp(tmp, name, (size_t)n) == 0;
        } else match = (ke - 1 - kb) == nlen && memcmp(kb, name, (size_t)nlen) == 0;
        p = skip_ws(ke, e);
        if (p >= e || *p != ':') return NULL;
        p = skip_ws(p + 1, e);
        const char* ve = skip_value(p, e);
        if (!ve) return NULL;
        if (match) { *vend = ve; return p; }
        p = skip_ws(ve, e);
        if (p < e && *p == ',') { ++p; continue; }
        return NULL;
    }
}
static long long go_i64(double d) { if (!(d >= -9223372036854775808.0 && d < 9223372036854775808.0)) return INT64_MIN; return (long long)d; }
static int f_is_int(double d) { return d == (double)go_i64(d); }
static V mk_missing(void) { V v; memset(&v, 0, sizeof v); v.t = T_MISSING; return v; }
static V mk_null(void) { V v; memset(&v, 0, sizeof v); v.t = T_NULL; return v; }
static V mk_bool(int b) { V v; memset(&v, 0, sizeof v); v.t = T_BOOL; v.b = b; return v; }
static V mk_int(long long i) { V v; memset(&v, 0, sizeof v); v.t = T_NUM; v.isint = 1; v.i = i; return v; }
static V mk_flt(double f) { V v; memset(&v, 0, sizeof v); v.t = T_NUM; v.isint = 0; v.f = f; return v; }
static V new_num(double d) { return f_is_int(d) ? mk_int(go_i64(d)) : mk_flt(d); }  /* value.NewValue */
static double num_f(V v) { return v.isint ? (double)v.i : v.f; }

static V parse_scalar(const char* p, const char* e, Arena* ar) {  /* value.NewParsedValue on a field's text */
    V v = mk_missing();
    if (p >= e) return v;
    char c = *p;
    if (c == '"') {
        int n = (int)(e - p - 2);
        char* buf = (char*)arena_alloc(ar, (size_t)(n > 0 ? n : 1));
        v.t = T_STR; v.slen = unescape(p + 1, e - 1, buf); v.s = buf;
        return v;
    }
    if (c == '{' || c == '[') { v.t = T_OTHER; v.s = p; v.slen = (int)(e - p); return v; }
    if (c == 't') return mk_bool(1);
    if (c == 'f') return mk_bool(0);
    if (c == 'n') return mk_null();
    char tmp[64];
    int n = (int)(e - p);
    if (n > 63) n = 63;
    memcpy(tmp, p, (size_t)n); tmp[n] = 0;
    int isfrac = 0;
    for (int k = 0; k < n; ++k) if (tmp[k] == '.' || tmp[k] == 'e' || tmp[k] == 'E') isfrac = 1;
    if (!isfrac) {
        char* end;
        long long iv;
        iv = strtoll(tmp, &end, 10);
        /* overflow -> float64 */
        int neg = tmp[0] == '-';
        int digits = n - neg;
        int over = digits > 19 || (digits == 19 && ((iv == INT64_MAX && strcmp(tmp, "9223372036854775807") != 0) || (iv == INT64_MIN && strcmp(tmp, "-9223372036854775808") != 0)));
        if (!over && *end == 0) return mk_int(iv);
    }
    return new_num(strtod(tmp, NULL));
}

/* ---- expression trees ------------------------------------------------------------------------------------------ */
enum { E_CI, E_CF, E_CS, E_CB, E_CNULL, E_CMISSING, E_FIELD, E_ADD, E_MULT, E_SUB, E_DIV, E_MOD, E_NEG, E_EQ, E_LT, E_LE,
       E_BETWEEN, E_IN, E_AND, E_OR, E_NOT, E_ISNULL, E_ISNOTNULL, E_ISMISSING, E_ISNOTMISSING, E_ISVALUED, E_ISNOTVALUED };
typedef struct Expr { int k; int nops; struct Expr** ops; V c; char** path; int* plen; int npath; } Expr;

static const char* KNAMES[] = {"ci", "cf", "cs", "cb", "cnull", "cmissing", "field", "add", "mult", "sub", "div", "mod", "neg", "eq", "lt", "le",
                               "between", "in", "and", "or", "not", "isnull", "isnotnull", "ismissing", "isnotmissing", "isvalued", "isnotvalued"};
static char* unhex(const char* h, int hl, int* outlen) {
    int n = hl / 2;
    char* b = (char*)malloc((size_t)n + 1);
    for (int i = 0; i < n; ++i) b[i] = (char)((hexv(h[2 * i]) << 4) | hexv(h[2 * i + 1]));
    b[n] = 0; *outlen = n; return b;
}
static Expr* parse_sexpr(const char** pp) {
    const char* p = *pp;
    while (*p == ' ') ++p;
    if (*p != '(') return NULL;
    ++p;
    const char* w = p;
    while (*p && *p != ' ' && *p != ')') ++p;
    Expr* e = (Expr*)calloc(1, sizeof(Expr));
    e->k = -1;
    for (int k = 0; k < (int)(sizeof KNAMES / sizeof KNAMES[0]); ++k)
        if ((int)strlen(KNAMES[k]) == (int)(p - w) && memcmp(KNAMES[k], w, (size_t)(p - w)) == 0) e->k = k;
    if (e->k < 0) return NULL;
    if (e->k <= E_CB || e->k == E_FIELD) {
        while (*p == ' ') {
            while (*p == ' ') ++p;
            if (*p == ')') break;
            const char* t = p;
            while (*p && *p != ' ' && *p != ')') ++p;
            if (e->k == E_CI) e->c = mk_int(strtoll(t, NULL, 10));
            else if (e->k == E_CF) { unsigned long long b = strtoull(t, NULL, 16); double d; memcpy(&d, &b, 8); e->c = mk_flt(d); }
            else if (e->k == E_CB) e->c = mk_bool(*t == '1');
            else if (e->k == E_CS) { int n; char* s = unhex(t, (int)(p - t), &n); e->c.t = T_STR; e->c.s = s; e->c.slen = n; }
            else {
                e->path = (char**)realloc(e->path, sizeof(char*) * (size_t)(e->npath + 1));
                e->plen = (int*)realloc(e->plen, sizeof(int) * (size_t)(e->npath + 1));
                e->path[e->npath] = unhex(t, (int)(p - t), &e->plen[e->npath]);
                e->npath++;
            }
        }
        if (e->k == E_CS && e->c.t != T_STR) { e->c.t = T_STR; e->c.s = ""; e->c.slen = 0; }
    } else if (e->k == E_CNULL) e->c = mk_null();
    else if (e->k == E_CMISSING) e->c = mk_missing();
    else {
        for (;;) {
            while (*p == ' ') ++p;
            if (*p != '(') break;
            Expr* c = parse_sexpr(&p);
            if (!c) return NULL;
            e->ops = (Expr**)realloc(e->ops, sizeof(Expr*) * (size_t)(e->nops + 1));
            e->ops[e->nops++] = c;
        }
    }
    while (*p == ' ') ++p;
    if (*p != ')') return NULL;
    *pp = p + 1;
    return e;
}

/* ---- value semantics (value/*.go) ---------------------------------------------------------------------------------- */
static int truth(V v) {
    switch (v.t) {
        case T_BOOL: return v.b;
        case T_NUM: return v.isint ? v.i != 0 : (v.f == v.f && v.f != 0.0);
        case T_STR: return v.slen > 0;
        case T_OTHER: return v.slen > 2;
        default: return 0;
    }
}
static int collate_f(double t, double o) {
    if (t != t) return o != o ? 0 : -1;
    if (o != o) return 1;
    return t < o ? -1 : (t > o ? 1 : 0);
}
static int collate(V a, V b) {
    if (a.t != b.t) return a.t - b.t;
    switch (a.t) {
        case T_BOOL: return a.b - b.b;
        case T_NUM:
            if (a.isint && b.isint) return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
            return collate_f(num_f(a), num_f(b));
        case T_STR: {
            int n = a.slen < b.slen ? a.slen : b.slen;
            int c = memcmp(a.s, b.s, (size_t)n);
            if (c) return c < 0 ? -1 : 1;
            return a.slen < b.slen ? -1 : (a.slen > b.slen ? 1 : 0);
        }
        default: return 0;
    }
}
#define CMP_NULL 8
#define CMP_MISSING 9
static int compare(V a, V b) {
    if (a.t == T_MISSING || b.t == T_MISSING) return CMP_MISSING;
    if (a.t == T_NULL || b.t == T_NULL) return CMP_NULL;
    int c = collate(a, b);
    return c < 0 ? -1 : (c > 0 ? 1 : 0);
}
static V equals(V a, V b) {
    if (a.t == T_MISSING || b.t == T_MISSING) return mk_missing();
    if (a.t == T_NULL || b.t == T_NULL) return mk_null();
    if (a.t != b.t) return mk_bool(0);
    if (a.t == T_NUM) { if (a.isint && b.isint) return mk_bool(a.i == b.i); return mk_bool(num_f(a) == num_f(b)); }
    if (a.t == T_BOOL) return mk_bool(a.b == b.b);
    if (a.t == T_STR) return mk_bool(a.slen == b.slen && memcmp(a.s, b.s, (size_t)a.slen) == 0);
    return mk_bool(a.slen == b.slen && memcmp(a.s, b.s, (size_t)a.slen) == 0);
}
static V num_add(V a, V b) {
    if (a.isint && b.isint) {
        long long rv = (long long)((unsigned long long)a.i + (unsigned long long)b.i);
        if ((a.i >= 0 && b.i >= 0 && rv >= 0) || (a.i < 0 && b.i < 0 && rv < 0)) return mk_int(rv);
    }
    return mk_flt(num_f(a) + num_f(b));
}
static V num_mult(V a, V b) {
    if (a.isint && b.isint) {
        long long rv = (long long)((unsigned long long)a.i * (unsigned long long)b.i);
        if (a.i == 0) return mk_int(rv);
        long long q = (a.i == -1) ? (long long)(0ULL - (unsigned long long)rv) : rv / a.i;  /* Go: MinInt64 / -1 wraps */
        if (q == b.i) return mk_int(rv);
    }
    return mk_flt(num_f(a) * num_f(b));
}
static V num_neg(V a) { if (a.isint) { if (a.i == INT64_MIN) return mk_flt(-(double)a.i); return mk_int(-a.i); } return mk_flt(-a.f); }
static V num_sub(V a, V b) {
    if (a.isint && b.isint && b.i > INT64_MIN) return num_add(a, mk_int(-b.i));
    return mk_flt(num_f(a) - num_f(b));
}

typedef struct { const char* doc; const char* end; Arena* ar; } Item;

static V eval(const Expr* e, const Item* it) {
    switch (e->k) {
        case E_CI: case E_CF: case E_CS: case E_CB: case E_CNULL: case E_CMISSING: return e->c;
        case E_FIELD: {
            const char* p = it->doc; const char* pe = it->end;
            /* parsed.go identifyType: only an OBJECT document has fields */
            for (int k = 0; k < e->npath; ++k) {
                const char* ve;
                const char* v = first_find(p, pe, e->path[k], e->plen[k], &ve, it->ar);
                if (!v) return mk_missing();
                p = v; pe = ve;
            }
            return parse_scalar(p, pe, it->ar);
        }
        case E_ADD: case E_MULT: {
            int null = 0; V acc = mk_int(e->k == E_ADD ? 0 : 1);
            for (int k = 0; k < e->nops; ++k) {
                V a = eval(e->ops[k], it);
                if (!null && a.t == T_NUM) acc = e->k == E_ADD ? num_add(acc, a) : num_mult(acc, a);
                else if (a.t == T_MISSING) return mk_missing();
                else null = 1;
            }
            return null ? mk_null() : acc;
        }
        case E_SUB: {
            V a = eval(e->ops[0], it), b = eval(e->ops[1], it);
            if (a.t == T_NUM && b.t == T_NUM) return num_sub(a, b);
            if (a.t == T_MISSING || b.t == T_MISSING) return mk_missing();
            return mk_null();
        }
        case E_DIV: case E_MOD: {
            V a = eval(e->ops[0], it), b = eval(e->ops[1], it);
            if (a.t == T_MISSING || b.t == T_MISSING) return mk_missing();
            if (b.t == T_NUM) {
                double s = num_f(b);
                if (s == 0.0) return mk_null();
                if (a.t == T_NUM) return new_num(e->k == E_DIV ? num_f(a) / s : fmod(num_f(a), s));
            }
            return mk_null();
        }
        case E_NEG: { V a = eval(e->ops[0], it); if (a.t == T_NUM) return num_neg(a); if (a.t == T_MISSING) return a; return mk_null(); }
        case E_EQ: return equals(eval(e->ops[0], it), eval(e->ops[1], it));
        case E_LT: case E_LE: {
            int c = compare(eval(e->ops[0], it), eval(e->ops[1], it));
            if (c == CMP_MISSING) return mk_missing();
            if (c == CMP_NULL) return mk_null();
            return mk_bool(e->k == E_LT ? c < 0 : c <= 0);
        }
        case E_BETWEEN: {
            V x = eval(e->ops[0], it), lo = eval(e->ops[1], it), hi = eval(e->ops[2], it);
            int lc = compare(x, lo);
            if (lc == CMP_MISSING) return mk_missing();
            int hc = compare(x, hi);
            if (hc == CMP_MISSING) return mk_missing();
            if (lc == CMP_NULL || hc == CMP_NULL) return mk_null();
            return mk_bool(lc >= 0 && hc <= 0);
        }
        case E_IN: {
            V x = eval(e->ops[0], it);
            if (x.t == T_MISSING) return x;
            int missing = 0, null = 0;
            for (int k = 1; k < e->nops; ++k) {
                V v = eval(e->ops[k], it);
                if (x.t > T_NULL && v.t > T_NULL) { V r = equals(x, v); if (r.t == T_BOOL && r.b) return mk_bool(1); }
                else if (v.t == T_MISSING) missing = 1;
                else null = 1;
            }
            return null ? mk_null() : (missing ? mk_missing() : mk_bool(0));
        }
        case E_AND: {
            int missing = 0, null = 0, isfalse = 0;
            for (int k = 0; k < e->nops; ++k) {
                V a = eval(e->ops[k], it);
                if (a.t == T_NULL) null = 1; else if (a.t == T_MISSING) missing = 1; else if (!truth(a)) { isfalse = 1; break; }
            }
            if (isfalse) return mk_bool(0);
            return missing ? mk_missing() : (null ? mk_null() : mk_bool(1));
        }
        case E_OR: {
            int missing = 0, null = 0, istrue = 0;
            for (int k = 0; k < e->nops; ++k) {
                V a = eval(e->ops[k], it);
                if (a.t == T_NULL) null = 1; else if (a.t == T_MISSING) missing = 1; else if (truth(a)) { istrue = 1; break; }
            }
            if (istrue) return mk_bool(1);
            return null ? mk_null() : (missing ? mk_missing() : mk_bool(0));
        }
        case E_NOT: { V a = eval(e->ops[0], it); if (a.t <= T_NULL) return a; return mk_bool(!truth(a)); }
        case E_ISNULL: { V a = eval(e->ops[0], it); return a.t == T_NULL ? mk_bool(1) : (a.t == T_MISSING ? a : mk_bool(0)); }
        case E_ISNOTNULL: { V a = eval(e->ops[0], it); return a.t == T_NULL ? mk_bool(0) : (a.t == T_MISSING ? a : mk_bool(1)); }
        case E_ISMISSING: return mk_bool(eval(e->ops[0], it).t == T_MISSING);
        case E_ISNOTMISSING: return mk_bool(eval(e->ops[0], it).t != T_MISSING);
        case E_ISVALUED: return mk_bool(eval(e->ops[0], it).t > T_NULL);
        case E_ISNOTVALUED: return mk_bool(eval(e->ops[0], it).t <= T_NULL);
    }
    return mk_missing();
}

/* ---- canonical encodings ------------------------------------------------------------------------------------------- */
typedef struct { char* p; size_t n, cap; } Buf;
static void buf_put(Buf* b, const void* s, size_t n) {
    if (b->n + n + 1 > b->cap) { b->cap = (b->n + n + 1) * 2; b->p = (char*)realloc(b->p, b->cap); }
    memcpy(b->p + b->n, s, n); b->n += n; b->p[b->n] = 0;
}
static void buf_printf_ll(Buf* b, long long v) { char t[32]; int n = snprintf(t, sizeof t, "%lld", v); buf_put(b, t, (size_t)n); }
/* identity of a value inside a group key / DISTINCT set: numbers by canonical text (1 == 1.0) */
static void canon(Buf* b, V v) {
    char t = (char)('0' + v.t);
    buf_put(b, &t, 1);
    if (v.t == T_BOOL) buf_put(b, v.b ? "1" : "0", 1);
    else if (v.t == T_NUM) {
        if (!v.isint && f_is_int(v.f)) { v.isint = 1; v.i = go_i64(v.f); }
        if (v.isint) { buf_put(b, "i", 1); buf_printf_ll(b, v.i); }
        else { unsigned long long u; memcpy(&u, &v.f, 8); char h[24]; int n = snprintf(h, sizeof h, "f%016llx", u); buf_put(b, h, (size_t)n); }
    } else if (v.t == T_STR || v.t == T_OTHER) { buf_printf_ll(b, v.slen); buf_put(b, ":", 1); buf_put(b, v.s, (size_t)v.slen); }
}

/* ---- string-keyed hash map ------------------------------------------------------------------------------------------- */
typedef struct { char* key; int klen; int idx; } Slot;
typedef struct { Slot* slots; int cap, n; } Map;
static unsigned long long fnv(const char* s, int n) { unsigned long long h = 1469598103934665603ULL; for (int i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; } return h; }
static void map_init(Map* m, int cap) { m->cap = cap; m->n = 0; m->slots = (Slot*)calloc((size_t)cap, sizeof(Slot)); }
static int map_get(Map* m, const char* key, int klen, int create_idx) {
    if (m->n * 2 >= m->cap) {
        Map big; map_init(&big, m->cap * 2);
        for (int i = 0; i < m->cap; ++i) if (m->slots[i].key) {
            unsigned long long h = fnv(m->slots[i].key, m->slots[i].klen) & (unsigned long long)(big.cap - 1);
            while (big.slots[h].key) h = (h + 1) & (unsigned long long)(big.cap - 1);
            big.slots[h] = m->slots[i]; big.n++;
        }
        free(m->slots); *m = big;
    }
    unsigned long long h = fnv(key, klen) & (unsigned long long)(m->cap - 1);
    while (m->slots[h].key) {
        if (m->slots[h].klen == klen && memcmp(m->slots[h].key, key, (size_t)klen) == 0) return m->slots[h].idx;
        h = (h + 1) & (unsigned long long)(m->cap - 1);
    }
    if (create_idx < 0) return -1;
    m->slots[h].key = (char*)malloc((size_t)klen + 1);
    memcpy(m->slots[h].key, key, (size_t)klen); m->slots[h].key[klen] = 0;
    m->slots[h].klen = klen; m->slots[h].idx = create_idx; m->n++;
    return create_idx;
}
static void map_free(Map* m) { for (int i = 0; i < m->cap; ++i) free(m->slots[i].key); free(m->slots); }

/* ---- aggregates (algebra/agg_*.go) ------------------------------------------------------------------------------------- */
enum { A_COUNT, A_COUNTN, A_SUM, A_AVG, A_MIN, A_MAX };
typedef struct { int kind, distinct, star; Expr* operand; } Agg;
typedef struct {
    int has;            /* SUM/AVG/MIN/MAX: cumulative is not NULL */
    V val;              /* COUNT: count ; SUM/AVG: sum ; MIN/MAX: winner (strings owned) */
    long long count;    /* AVG */
    Map set; int set_init; V* setvals; int nset, setcap;  /* DISTINCT */
} AggState;
typedef struct { V* keys; AggState* st; } Group;

static V own(V v) { if (v.t == T_STR || v.t == T_OTHER) { char* s = (char*)malloc((size_t)v.slen + 1); memcpy(s, v.s, (size_t)v.slen); s[v.slen] = 0; v.s = s; } return v; }

static void agg_default(const Agg* a, AggState* s) {
    memset(s, 0, sizeof *s);
    if (a->kind == A_COUNT || a->kind == A_COUNTN) s->val = mk_int(0);
}
static void set_add(AggState* s, V v, Buf* tmp) {
    if (!s->set_init) { map_init(&s->set, 16); s->set_init = 1; }
    tmp->n = 0; canon(tmp, v);
    if (map_get(&s->set, tmp->p, (int)tmp->n, -1) >= 0) return;
    map_get(&s->set, tmp->p, (int)tmp->n, s->nset);
    if (s->nset == s->setcap) { s->setcap = s->setcap ? s->setcap * 2 : 8; s->setvals = (V*)realloc(s->setvals, sizeof(V) * (size_t)s->setcap); }
    s->setvals[s->nset++] = own(v);
}
static void cumulate_part(const Agg* a, AggState* s, V part, long long pcount) {  /* Sum/Avg/Min/Max cumulatePart */
    if (!s->has) { s->has = 1; s->val = own(part); s->count = pcount; return; }
    if (a->kind == A_SUM || a->kind == A_AVG) { s->val = num_add(s->val, part); s->count += pcount; }
    else if (a->kind == A_MIN) { if (collate(part, s->val) < 0) s->val = own(part); }
    else if (a->kind == A_MAX) { if (collate(part, s->val) > 0) s->val = own(part); }
}
static void cumulate_initial(const Agg* a, AggState* s, const Item* it, Buf* tmp) {
    V o = mk_missing();
    if (!a->star) o = eval(a->operand, it);
    if (a->distinct) {
        if (a->kind == A_COUNT) { if (o.t <= T_NULL) return; }
        else if (o.t != T_NUM) return;
        set_add(s, o, tmp);
        return;
    }
    switch (a->kind) {
        case A_COUNT: if (!a->star && o.t <= T_NULL) return; s->val = num_add(s->val, mk_int(1)); break;
        case A_COUNTN: if (o.t != T_NUM) return; s->val = num_add(s->val, mk_int(1)); break;
        case A_SUM: case A_AVG: if (o.t != T_NUM) return; cumulate_part(a, s, o, 1); break;
        default: if (o.t <= T_NULL) return; cumulate_part(a, s, o, 0); break;
    }
}
static void cumulate_intermediate(const Agg* a, AggState* c, AggState* part, Buf* tmp) {
    if (a->distinct) { for (int i = 0; i < part->nset; ++i) set_add(c, part->setvals[i], tmp); return; }
    if (a->kind == A_COUNT || a->kind == A_COUNTN) { c->val = num_add(c->val, part->val); return; }
    if (!part->has) return;
    cumulate_part(a, c, part->val, part->count);
}
static V compute_final(const Agg* a, AggState* s) {
    if (a->distinct) {
        if (a->kind == A_COUNT || a->kind == A_COUNTN) return mk_int(s->nset);
        if (s->nset == 0) return mk_null();
        V sum = mk_int(0);  /* value.ZERO_NUMBER */
        for (int i = 0; i < s->nset; ++i) if (!s->setvals[i].isint) sum = num_add(sum, s->setvals[i]);  /* floats map first (set.go:217-265) */
        for (int i = 0; i < s->nset; ++i) if (s->setvals[i].isint) sum = num_add(sum, s->setvals[i]);
        if (a->kind == A_SUM) return sum;
        return new_num(num_f(sum) / (double)s->nset);
    }
    if (a->kind == A_COUNT || a->kind == A_COUNTN) return s->val;
    if (!s->has) return mk_null();
    if (a->kind == A_AVG) { if (s->count > 0) return new_num(num_f(s->val) / (double)s->count); return mk_null(); }
    return s->val;
}

/* ---- the chain ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const char* buf; const long long* offs; long long lo, hi;
    Expr* where; Expr** keys; int nkeys; Agg* aggs; int naggs;
    Map map; Group* groups; int ngroups, gcap;
    long long passed;
} Stream;

static int find_or_seed(Stream* s, const char* key, int klen, const V* kv) {
    int g = map_get(&s->map, key, klen, -1);
    if (g >= 0) return g;
    if (s->ngroups == s->gcap) { s->gcap = s->gcap ? s->gcap * 2 : 64; s->groups = (Group*)realloc(s->groups, sizeof(Group) * (size_t)s->gcap); }
    g = s->ngroups++;
    map_get(&s->map, key, klen, g);
    s->groups[g].keys = (V*)malloc(sizeof(V) * (size_t)(s->nkeys ? s->nkeys : 1));
    for (int k = 0; k < s->nkeys; ++k) s->groups[g].keys[k] = own(kv[k]);
    s->groups[g].st = (AggState*)malloc(sizeof(AggState) * (size_t)(s->naggs ? s->naggs : 1));
    for (int a = 0; a < s->naggs; ++a) agg_default(&s->aggs[a], &s->groups[g].st[a]);
    return g;
}

static void* initial_stream(void* arg) {  /* Filter + InitialGroup over one row range */
    Stream* s = (Stream*)arg;
    Arena ar = {0};
    Buf key = {0}, tmp = {0};
    V kv[16];
    map_init(&s->map, 1024);
    for (long long d = s->lo; d < s->hi; ++d) {
        arena_reset(&ar);
        Item it = {s->buf + s->offs[d], s->buf + s->offs[d + 1], &ar};
        /* parsed.go:76-98: sniff; a non-object document has no fields (eval handles it: first_find needs '{') */
        const char* p = it.doc;
        while (p < it.end && (*p == ' ' || *p == '\t' || *p == '\n')) ++p;
        it.doc = p;
        if (s->where && !truth(eval(s->where, &it))) continue;
        s->passed++;
        key.n = 0;
        for (int k = 0; k < s->nkeys; ++k) {
            kv[k] = eval(s->keys[k], &it);
            if (kv[k].t != T_MISSING) { char idx = (char)('A' + k); buf_put(&key, &idx, 1); canon(&key, kv[k]); }
        }
        if (!key.p) buf_put(&key, "", 0);
        int g = find_or_seed(s, key.p, (int)key.n, kv);
        for (int a = 0; a < s->naggs; ++a) cumulate_initial(&s->aggs[a], &s->groups[g].st[a], &it, &tmp);
    }
    free(key.p); free(tmp.p); arena_free(&ar);
    return NULL;
}

static void render(Buf* out, V v) {
    char t[64];
    switch (v.t) {
        case T_MISSING: buf_put(out, "{\"$missing\":1}", 14); break;
        case T_NULL: buf_put(out, "null", 4); break;
        case T_BOOL: buf_put(out, v.b ? "true" : "false", v.b ? 4 : 5); break;
        case T_NUM:
            if (v.isint) buf_printf_ll(out, v.i);
            else { unsigned long long u; memcpy(&u, &v.f, 8); int n = snprintf(t, sizeof t, "{\"$f\":\"%016llx\"}", u); buf_put(out, t, (size_t)n); }
            break;
        default: {
            buf_put(out, "{\"$s\":\"", 7);
            for (int i = 0; i < v.slen; ++i) { int n = snprintf(t, sizeof t, "%02x", (unsigned char)v.s[i]); buf_put(out, t, (size_t)n); }
            buf_put(out, "\"}", 2);
        }
    }
}

/* Public entry.  Returns a malloc'd JSON text: [{"k":[..],"a":[..]},..]; *elapsed = seconds in the chain. */
char* oracle_run(const char* buf, const long long* offs, long long ndocs, const char* where, const char* const* keys, int nkeys,
                 const char* const* aggs, int naggs, int threads, double* elapsed, long long* rows_passed) {
    struct timespec t0, t1;
    Expr* w = NULL;
    if (where && *where) { const char* p = where; w = parse_sexpr(&p); if (!w) return NULL; }
    Expr** ks = (Expr**)calloc((size_t)(nkeys ? nkeys : 1), sizeof(Expr*));
    for (int k = 0; k < nkeys; ++k) { const char* p = keys[k]; ks[k] = parse_sexpr(&p); if (!ks[k]) return NULL; }
    Agg* as = (Agg*)calloc((size_t)(naggs ? naggs : 1), sizeof(Agg));
    for (int a = 0; a < naggs; ++a) {
        char kind[16]; int dist = 0, off = 0;
        if (sscanf(aggs[a], "%15s %d %n", kind, &dist, &off) < 2) return NULL;
        static const char* kn[] = {"count", "countn", "sum", "avg", "min", "max"};
        as[a].kind = -1;
        for (int k = 0; k < 6; ++k) if (strcmp(kind, kn[k]) == 0) as[a].kind = k;
        if (as[a].kind < 0) return NULL;
        as[a].distinct = dist;
        const char* p = aggs[a] + off;
        if (*p == '*') as[a].star = 1; else { as[a].operand = parse_sexpr(&p); if (!as[a].operand) return NULL; }
    }
    if (threads < 1) threads = 1;
    if ((long long)threads > ndocs) threads = ndocs > 0 ? (int)ndocs : 1;
    Stream* st = (Stream*)calloc((size_t)threads, sizeof(Stream));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; ++t) {
        st[t].buf = buf; st[t].offs = offs; st[t].lo = ndocs * t / threads; st[t].hi = ndocs * (t + 1) / threads;
        st[t].where = w; st[t].keys = ks; st[t].nkeys = nkeys; st[t].aggs = as; st[t].naggs = naggs;
        if (threads == 1) initial_stream(&st[t]); else pthread_create(&th[t], NULL, initial_stream, &st[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    /* IntermediateGroup: first arrival stored as-is, later ones merged (stream order) */
    Stream* fin = &st[0];
    Buf key = {0}, tmp = {0};
    long long passed = st[0].passed;
    for (int t = 1; t < threads; ++t) {
        passed += st[t].passed;
        for (int g = 0; g < st[t].ngroups; ++g) {
            key.n = 0;
            for (int k = 0; k < nkeys; ++k) if (st[t].groups[g].keys[k].t != T_MISSING) { char idx = (char)('A' + k); buf_put(&key, &idx, 1); canon(&key, st[t].groups[g].keys[k]); }
            if (!key.p) buf_put(&key, "", 0);
            int had = fin->ngroups;
            int gi = find_or_seed(fin, key.p, (int)key.n, st[t].groups[g].keys);
            (void)had;
            for (int a = 0; a < naggs; ++a) cumulate_intermediate(&as[a], &fin->groups[gi].st[a], &st[t].groups[g].st[a], &tmp);
        }
    }
    /* FinalGroup */
    Buf out = {0};
    buf_put(&out, "[", 1);
    int emitted = 0;
    for (int g = 0; g < fin->ngroups; ++g) {
        if (emitted++) buf_put(&out, ",", 1);
        buf_put(&out, "{\"k\":[", 6);
        for (int k = 0; k < nkeys; ++k) { if (k) buf_put(&out, ",", 1); render(&out, fin->groups[g].keys[k]); }
        buf_put(&out, "],\"a\":[", 7);
        for (int a = 0; a < naggs; ++a) { if (a) buf_put(&out, ",", 1); render(&out, compute_final(&as[a], &fin->groups[g].st[a])); }
        buf_put(&out, "]}", 2);
    }
    if (nkeys == 0 && fin->ngroups == 0) {  /* group_final.go:108-117 */
        buf_put(&out, "{\"k\":[],\"a\":[", 13);
        for (int a = 0; a < naggs; ++a) { AggState s; agg_default(&as[a], &s); if (a) buf_put(&out, ",", 1); render(&out, compute_final(&as[a], &s)); }
        buf_put(&out, "]}", 2);
    }
    buf_put(&out, "]", 1);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (elapsed) *elapsed = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (rows_passed) *rows_passed = passed;
    free(key.p); free(tmp.p);
    for (int t = 0; t < threads; ++t) map_free(&st[t].map);   /* group payloads are left to process exit: test tool */
    free(st); free(th);
    return out.p;
}
void oracle_free(char* p) { free(p); }

/* ---- synthetic documents of the BASELINE.json configs (deterministic per (seed, row), thread-count independent) ----------- */
static unsigned long long splitmix(unsigned long long x) { x += 0x9e3779b97f4a7c15ULL; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL; x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL; return x ^ (x >> 31); }
typedef struct { int config; long long first, lo, hi; unsigned long long seed; char* buf; long long* offs; long long base; } GenJob;
#define ZIPF_VOCAB 100000
static double zipf_cdf[ZIPF_VOCAB];
static pthread_once_t zipf_once = PTHREAD_ONCE_INIT;
static void zipf_init(void) {
    double tot = 0, acc = 0;
    for (int i = 0; i < ZIPF_VOCAB; ++i) tot += pow((double)(i + 1), -1.1);
    for (int i = 0; i < ZIPF_VOCAB; ++i) { acc += pow((double)(i + 1), -1.1) / tot; zipf_cdf[i] = acc; }
    zipf_cdf[ZIPF_VOCAB - 1] = 1.0;
}
static int gen_doc(int config, unsigned long long seed, long long row, char* out) {
    unsigned long long r1 = splitmix(seed ^ (unsigned long long)row * 0x2545F4914F6CDD1DULL), r2 = splitmix(r1), r3 = splitmix(r2), r4 = splitmix(r3);
    if (config == 2)  /* {"id":i,"n":U[0,1e6),"f":U[0,1) 6 decimals,"type":"t0..15"} */
        return sprintf(out, "{\"id\":%lld,\"n\":%llu,\"f\":0.%06llu,\"type\":\"t%llu\"}", row, r1 % 1000000ULL, r2 % 1000000ULL, r3 % 16ULL);
    if (config == 3) {  /* TPC-H Q1 shaped lineitem */
        static const char* rf[] = {"A", "N", "R"}; static const char* ls[] = {"F", "O"};
        unsigned y = 1992 + (unsigned)(r1 % 7), m = 1 + (unsigned)(r2 % 12), d = 1 + (unsigned)(r3 % 28);
        return sprintf(out, "{\"l_orderkey\":%lld,\"l_quantity\":%llu,\"l_extendedprice\":%llu.%02llu,\"l_discount\":0.%02llu,\"l_tax\":0.%02llu,"
                            "\"l_returnflag\":\"%s\",\"l_linestatus\":\"%s\",\"l_shipdate\":\"%04u-%02u-%02u\"}",
                       row, 1 + r4 % 50ULL, 900 + (r1 >> 20) % 104000ULL, (r2 >> 20) % 100ULL, (r3 >> 20) % 11ULL, (r4 >> 20) % 9ULL,
                       rf[(r1 >> 40) % 3], ls[(r2 >> 40) % 2], y, m, d);
    }
    if (config == 4)  /* {"g":U[0,1e6),"x":U[0,1000),"y":float} */
        return sprintf(out, "{\"g\":%llu,\"x\":%llu,\"y\":%llu.%03llu}", r1 % 1000000ULL, r2 % 1000ULL, r3 % 1000ULL, r4 % 1000ULL);
    /* config 5: Zipf(s = 1.1) string key over a 100k vocabulary (inverse CDF, as tools/workloads.py generates the device
       columns), 10% MISSING + 10% null on k and v */
    {
        double u = (double)(r1 >> 11) / 9007199254740992.0;
        pthread_once(&zipf_once, zipf_init);
        int zlo = 0, zhi = ZIPF_VOCAB - 1;  /* first rank whose cumulative probability reaches u */
        while (zlo < zhi) { int mid = (zlo + zhi) >> 1; if (zipf_cdf[mid] >= u) zhi = mid; else zlo = mid + 1; }
        unsigned long long w = (unsigned long long)zlo;
        char k[64], v[64];
        unsigned long long mk = r2 % 10ULL, mv = r3 % 10ULL;
        if (mk == 0) k[0] = 0; else if (mk == 1) sprintf(k, "\"k\":null,"); else sprintf(k, "\"k\":\"w%06llu-%llx\",", w, splitmix(w) & 0xffffffULL);
        if (mv == 0) v[0] = 0; else if (mv == 1) sprintf(v, "\"v\":null,"); else sprintf(v, "\"v\":%lld,", (long long)(r4 % 1001000ULL) - 1000);
        return sprintf(out, "{%s%s\"id\":%lld}", k, v, row);
    }
}
/* Fills buf/offs with docs [first, first+n); returns bytes used (or -1 when cap is too small). Single-threaded per call. */
long long oracle_gen_docs(int config, unsigned long long seed, long long first, long long n, char* buf, long long cap, long long* offs) {
    long long at = 0;
    const int newline = config & 0x100;  /* NDJSON: a line end closes every document (trailing white space of the document) */
    config &= 0xff;
    for (long long i = 0; i < n; ++i) {
        if (at + 512 > cap) return -1;
        offs[i] = at;
        at += gen_doc(config, seed, first + i, buf + at);
        if (newline) buf[at++] = '\n';
    }
    offs[n] = at;
    return at;
}
