// json.hpp — a small validating JSON scanner (shared by the column shredder and the plan-JSON reader)
// and a DOM for plan JSON.  Numbers keep the int64 / float64 distinction the reference relies on
// (test/multistore/test_cases/integers/case_select.json: 9223372036854775807 survives exactly).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <map>

#include "common.hpp"

namespace n1 {
namespace json {

struct Scanner {
    const char* p;
    const char* end;
    bool ok = true;
    Scanner(const char* b, const char* e) : p(b), end(e) {}

    void ws() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    bool fail() { ok = false; return false; }

    // Scans a string starting at the opening quote.  raw_begin/raw_end delimit the bytes between the
    // quotes; has_escape tells whether unescape() is needed.
    bool string_raw(const char*& rb, const char*& re, bool& has_escape) {
        if (p >= end || *p != '"') return fail();
        ++p;
        rb = p;
        has_escape = false;
        while (p < end) {
            unsigned char c = (unsigned char)*p;
            if (c == '"') { re = p; ++p; return true; }
            if (c == '\\') {
                has_escape = true;
                ++p;
                if (p >= end) return fail();
                char e = *p;
                if (e == 'u') {
                    if (end - p < 5) return fail();
                    for (int i = 1; i <= 4; ++i) if (!isxdigit((unsigned char)p[i])) return fail();
                    p += 5;
                } else if (e == '"' || e == '\\' || e == '/' || e == 'b' || e == 'f' || e == 'n' || e == 'r' || e == 't') {
                    ++p;
                } else return fail();
                continue;
            }
            if (c < 0x20) return fail();
            ++p;
        }
        return fail();
    }

    static void put_utf8(std::string& out, u32 cp) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    static u32 hex4(const char* s) {
        u32 v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = s[i];
            v = v * 16 + (c <= '9' ? c - '0' : ((c | 0x20) - 'a' + 10));
        }
        return v;
    }
    // Go's encoding/json-style unescape: lone surrogates become U+FFFD.
    static void unescape(const char* rb, const char* re, std::string& out) {
        out.clear();
        for (const char* s = rb; s < re;) {
            if (*s != '\\') { out.push_back(*s++); continue; }
            char e = s[1];
            s += 2;
            switch (e) {
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    u32 cp = hex4(s);
                    s += 4;
                    if (cp >= 0xD800 && cp < 0xDC00) {
                        if (re - s >= 6 && s[0] == '\\' && s[1] == 'u') {
                            u32 lo = hex4(s + 2);
                            if (lo >= 0xDC00 && lo < 0xE000) { cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); s += 6; }
                            else cp = 0xFFFD;
                        } else cp = 0xFFFD;
                    } else if (cp >= 0xDC00 && cp < 0xE000) cp = 0xFFFD;
                    put_utf8(out, cp);
                    break;
                }
                default: out.push_back(e);
            }
        }
    }

    // Scans a number; is_int tells whether it is an integer literal that fits int64.
    bool number(bool& is_int, i64& iv, double& dv) {
        const char* s = p;
        if (p < end && *p == '-') ++p;
        if (p >= end) return fail();
        if (*p == '0') ++p;
        else if (*p >= '1' && *p <= '9') { while (p < end && *p >= '0' && *p <= '9') ++p; }
        else return fail();
        bool frac = false;
        if (p < end && *p == '.') {
            frac = true;
            ++p;
            if (p >= end || !(*p >= '0' && *p <= '9')) return fail();
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        if (p < end && (*p == 'e' || *p == 'E')) {
            frac = true;
            ++p;
            if (p < end && (*p == '+' || *p == '-')) ++p;
            if (p >= end || !(*p >= '0' && *p <= '9')) return fail();
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        if (!frac) {
            auto r = std::from_chars(s, p, iv);
            if (r.ec == std::errc() && r.ptr == p) { is_int = true; return true; }
        }
        is_int = false;
        auto r = std::from_chars(s, p, dv);
        if (r.ec == std::errc::result_out_of_range) {  // Go: ParseFloat range error -> document invalid
            return fail();
        }
        if (r.ec != std::errc()) return fail();
        return true;
    }

    // Grammar check of a number that nobody reads (json.Validate looks at syntax only: a skipped 1e999 does not make
    // the document invalid - the device shredder treats skipped numbers the same way - and no conversion is paid for)
    bool skip_number() {
        if (p < end && *p == '-') ++p;
        if (p >= end) return fail();
        if (*p == '0') ++p;
        else if (*p >= '1' && *p <= '9') { while (p < end && *p >= '0' && *p <= '9') ++p; }
        else return fail();
        if (p < end && *p == '.') {
            ++p;
            if (p >= end || !(*p >= '0' && *p <= '9')) return fail();
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        if (p < end && (*p == 'e' || *p == 'E')) {
            ++p;
            if (p < end && (*p == '+' || *p == '-')) ++p;
            if (p >= end || !(*p >= '0' && *p <= '9')) return fail();
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        return true;
    }

    bool literal(const char* w) {
        size_t n = strlen(w);
        if ((size_t)(end - p) < n || memcmp(p, w, n) != 0) return fail();
        p += n;
        return true;
    }

    // Validating skip of one value.
    bool skip() {
        ws();
        if (p >= end) return fail();
        char c = *p;
        if (c == '{') {
            ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                const char *rb, *re; bool esc;
                if (!string_raw(rb, re, esc)) return false;
                ws();
                if (p >= end || *p != ':') return fail();
                ++p;
                if (!skip()) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail();
            }
        }
        if (c == '[') {
            ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                if (!skip()) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail();
            }
        }
        if (c == '"') { const char *rb, *re; bool esc; return string_raw(rb, re, esc); }
        if (c == 't') return literal("true");
        if (c == 'f') return literal("false");
        if (c == 'n') return literal("null");
        return skip_number();
    }
};

// ---- DOM (plan JSON, small inputs) ---------------------------------------------------------------
struct Node {
    enum Kind { NUL, BOOL, INT, FLOAT, STR, ARR, OBJ } kind = NUL;
    bool b = false;
    i64 i = 0;
    double d = 0;
    std::string s;
    std::vector<Node> arr;
    std::vector<std::pair<std::string, Node>> obj;
    const Node* get(const char* name) const {
        for (auto& kv : obj) if (kv.first == name) return &kv.second;
        return nullptr;
    }
    std::string str_or(const char* name, const std::string& dflt) const {
        const Node* n = get(name);
        return n && n->kind == STR ? n->s : dflt;
    }
};

inline bool parse_node(Scanner& sc, Node& out) {
    sc.ws();
    if (sc.p >= sc.end) return sc.fail();
    char c = *sc.p;
    if (c == '{') {
        out.kind = Node::OBJ;
        ++sc.p; sc.ws();
        if (sc.p < sc.end && *sc.p == '}') { ++sc.p; return true; }
        for (;;) {
            sc.ws();
            const char *rb, *re; bool esc;
            if (!sc.string_raw(rb, re, esc)) return false;
            std::string key;
            if (esc) Scanner::unescape(rb, re, key); else key.assign(rb, re);
            sc.ws();
            if (sc.p >= sc.end || *sc.p != ':') return sc.fail();
            ++sc.p;
            out.obj.emplace_back(key, Node());
            if (!parse_node(sc, out.obj.back().second)) return false;
            sc.ws();
            if (sc.p < sc.end && *sc.p == ',') { ++sc.p; continue; }
            if (sc.p < sc.end && *sc.p == '}') { ++sc.p; return true; }
            return sc.fail();
        }
    }
    if (c == '[') {
        out.kind = Node::ARR;
        ++sc.p; sc.ws();
        if (sc.p < sc.end && *sc.p == ']') { ++sc.p; return true; }
        for (;;) {
            out.arr.emplace_back();
            if (!parse_node(sc, out.arr.back())) return false;
            sc.ws();
            if (sc.p < sc.end && *sc.p == ',') { ++sc.p; continue; }
            if (sc.p < sc.end && *sc.p == ']') { ++sc.p; return true; }
            return sc.fail();
        }
    }
    if (c == '"') {
        const char *rb, *re; bool esc;
        if (!sc.string_raw(rb, re, esc)) return false;
        out.kind = Node::STR;
        if (esc) Scanner::unescape(rb, re, out.s); else out.s.assign(rb, re);
        return true;
    }
    if (c == 't') { out.kind = Node::BOOL; out.b = true; return sc.literal("true"); }
    if (c == 'f') { out.kind = Node::BOOL; out.b = false; return sc.literal("false"); }
    if (c == 'n') { out.kind = Node::NUL; return sc.literal("null"); }
    bool ii;
    if (!sc.number(ii, out.i, out.d)) return false;
    out.kind = ii ? Node::INT : Node::FLOAT;
    return true;
}

inline bool parse(const std::string& text, Node& out) {
    Scanner sc(text.data(), text.data() + text.size());
    if (!parse_node(sc, out)) return false;
    sc.ws();
    return sc.p == sc.end;
}

inline void quote(const std::string& s, std::string& out) {  // JSON string, no HTML escaping (MarshalNoEscape)
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
                else out.push_back((char)c);
        }
    }
    out.push_back('"');
}

// strconv.FormatFloat(f,'f',-1,64) with -0 -> 0 (value/float.go:31-48): shortest round-trip digits, no exponent.
// Number text of a floatValue: Go's strconv.FormatFloat(f, 'f', -1, 64) - the shortest digits that read back as f, laid
// out in fixed notation (value/float.go:31-48; "-0" is written "0").  std::to_chars' own fixed format prints the exact
// binary value for large magnitudes (1e300 -> 301 exact digits), so the shortest digits are taken from its scientific
// form and laid out here.
inline std::string format_float(double f) {
    if (std::isnan(f)) return "\"NaN\"";
    if (std::isinf(f)) return f > 0 ? "\"+Infinity\"" : "\"-Infinity\"";
    if (f == 0) return "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, f, std::chars_format::scientific);
    std::string sci(buf, r.ptr), digits, out;
    size_t i = 0;
    if (sci[i] == '-') { out = "-"; ++i; }
    for (; i < sci.size() && sci[i] != 'e'; ++i) if (sci[i] != '.') digits.push_back(sci[i]);
    const int exp10 = std::atoi(sci.c_str() + i + 1);
    if (exp10 >= 0) {
        const size_t int_len = (size_t)exp10 + 1;
        if (digits.size() <= int_len) out += digits + std::string(int_len - digits.size(), '0');
        else out += digits.substr(0, int_len) + "." + digits.substr(int_len);
    } else {
        out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    }
    return out;
}

}  // namespace json
}  // namespace n1
