// table.cpp — the columnar shredder (replaces PrimaryScan + Fetch over datastore/file) and the HBM
// column store.
//
// Reference behaviour restated (file:line under /root/reference):
//   datastore/file/file.go:711-730  ScanEntries = sorted ReadDir, every non-directory entry is a document
//   datastore/file/file.go:732-749  fetch = ReadFile + value.NewValue(bytes); key = name minus extension
//   value/parsed.go:38-98           type sniff skips ' ', '\t', '\n'; invalid JSON -> BINARY (all fields MISSING)
//   value/parsed.go:159-207         Field(): first occurrence of the name; non-object -> MISSING
//   value/value.go:367-430          NewValue: integral float64 -> int64
#include "table.hpp"

#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <thread>

#include "json.hpp"

namespace n1 {

double now_sec() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

std::vector<std::string> split_path(const std::string& path) {
    std::vector<std::string> out;
    size_t s = 0;
    for (;;) {
        size_t e = path.find('\x1f', s);
        if (e == std::string::npos) { out.push_back(path.substr(s)); break; }
        out.push_back(path.substr(s, e - s));
        s = e + 1;
    }
    return out;
}
std::string join_path(const std::vector<std::string>& p, char sep) {
    std::string s;
    for (size_t i = 0; i < p.size(); ++i) { if (i) s.push_back(sep); s += p[i]; }
    return s;
}

int Table::add_column(const std::string& path) {
    if (appended || sealed) N1_THROW(N1GPU_E_INVALID, "columns must be declared before documents are appended");
    int f = find_column(path);
    if (f >= 0) return f;
    if (cols.size() >= 16) N1_THROW(N1GPU_E_INELIGIBLE, "more than 16 referenced columns");
    cols.emplace_back();
    cols.back().path = split_path(path);
    return (int)cols.size() - 1;
}
int Table::find_column(const std::string& path) const {
    auto p = split_path(path);
    for (size_t i = 0; i < cols.size(); ++i) if (cols[i].path == p) return (int)i;
    return -1;
}

// ---- path trie -----------------------------------------------------------------------------------
namespace {
struct Trie {
    struct Node { std::vector<std::pair<std::string, int>> kids; int col = -1; };
    std::vector<Node> nodes;
    Trie() { nodes.emplace_back(); }
    void add(const std::vector<std::string>& path, int col) {
        int n = 0;
        for (auto& name : path) {
            int next = -1;
            for (auto& k : nodes[n].kids) if (k.first == name) next = k.second;
            if (next < 0) { nodes.emplace_back(); next = (int)nodes.size() - 1; nodes[n].kids.emplace_back(name, next); }
            n = next;
        }
        nodes[n].col = col;
    }
    int child(int n, const char* b, const char* e) const {
        size_t len = (size_t)(e - b);
        for (auto& k : nodes[n].kids) if (k.first.size() == len && memcmp(k.first.data(), b, len) == 0) return k.second;
        return -1;
    }
};

struct ChunkCol {
    std::vector<i64> payload;
    std::vector<u8> tags;
    std::unordered_map<std::string, u32> index;
    std::vector<std::string> strings;
    u32 code(const std::string& s) {
        auto it = index.find(s);
        if (it != index.end()) return it->second;
        u32 c = (u32)strings.size();
        index.emplace(s, c);
        strings.push_back(s);
        return c;
    }
};

struct Shredder {
    const Trie& trie;
    std::vector<ChunkCol>& cols;
    i64 row = 0;
    std::string tmp;
    Shredder(const Trie& t, std::vector<ChunkCol>& c) : trie(t), cols(c) {}

    void put(int col, u8 tag, i64 payload) {
        if (cols[col].tags[row] != C_MISSING) return;  // first occurrence wins (FirstFind)
        cols[col].tags[row] = tag;
        cols[col].payload[row] = payload;
    }

    // value at trie node `n` (n has a column and/or children)
    bool value(json::Scanner& sc, int n) {
        sc.ws();
        if (sc.p >= sc.end) return sc.fail();
        const Trie::Node& node = trie.nodes[n];
        char c = *sc.p;
        if (c == '{') {
            if (node.col >= 0) put(node.col, C_OTHER, 0);
            if (node.kids.empty()) return sc.skip();
            return object(sc, n);
        }
        if (c == '[') {
            if (node.col >= 0) put(node.col, C_OTHER, 0);
            return sc.skip();
        }
        if (node.col < 0) return sc.skip();
        if (c == '"') {
            const char *rb, *re; bool esc;
            if (!sc.string_raw(rb, re, esc)) return false;
            if (cols[node.col].tags[row] == C_MISSING) {
                if (esc) json::Scanner::unescape(rb, re, tmp); else tmp.assign(rb, re);
                put(node.col, C_STRING, (i64)cols[node.col].code(tmp));
            }
            return true;
        }
        if (c == 't') { if (!sc.literal("true")) return false; put(node.col, C_TRUE, 0); return true; }
        if (c == 'f') { if (!sc.literal("false")) return false; put(node.col, C_FALSE, 0); return true; }
        if (c == 'n') { if (!sc.literal("null")) return false; put(node.col, C_NULL, 0); return true; }
        bool ii; i64 iv; double dv;
        if (!sc.number(ii, iv, dv)) return false;
        if (ii) put(node.col, C_INT, iv);
        else if (f_is_int(dv)) put(node.col, C_INT, go_i64(dv));  // NewValue canonicalisation
        else { i64 b; memcpy(&b, &dv, 8); put(node.col, C_FLOAT, b); }
        return true;
    }

    bool object(json::Scanner& sc, int n) {
        ++sc.p;  // '{'
        sc.ws();
        if (sc.p < sc.end && *sc.p == '}') { ++sc.p; return true; }
        for (;;) {
            sc.ws();
            const char *rb, *re; bool esc;
            if (!sc.string_raw(rb, re, esc)) return false;
            int kid;
            if (esc) { json::Scanner::unescape(rb, re, tmp); kid = trie.child(n, tmp.data(), tmp.data() + tmp.size()); }
            else kid = trie.child(n, rb, re);
            sc.ws();
            if (sc.p >= sc.end || *sc.p != ':') return sc.fail();
            ++sc.p;
            if (kid >= 0) { if (!value(sc, kid)) return false; }
            else if (!sc.skip()) return false;
            sc.ws();
            if (sc.p < sc.end && *sc.p == ',') { ++sc.p; continue; }
            if (sc.p < sc.end && *sc.p == '}') { ++sc.p; return true; }
            return sc.fail();
        }
    }

    void document(const char* b, const char* e) {
        const char* s = b;
        while (s < e && (*s == ' ' || *s == '\t' || *s == '\n')) ++s;  // identifyType, parsed.go:76-98
        bool valid = false;
        if (s < e && *s == '{') {
            json::Scanner sc(s, e);
            if (object(sc, 0)) { sc.ws(); valid = sc.p == sc.end; }
        }
        if (!valid)  // non-object or BINARY document: every field access is MISSING
            for (auto& c : cols) { c.tags[row] = C_MISSING; c.payload[row] = 0; }
        ++row;
    }
};
}  // namespace

void host_shred_docs(const std::vector<Column>& cols, const char* buf, const i64* offs, const i64* rows, i64 n,
                     std::vector<HostShredOut>& out) {
    Trie trie;
    for (size_t c = 0; c < cols.size(); ++c) trie.add(cols[c].path, (int)c);
    std::vector<ChunkCol> cc(cols.size());
    for (auto& c : cc) { c.payload.assign((size_t)n, 0); c.tags.assign((size_t)n, C_MISSING); }
    Shredder sh(trie, cc);
    for (i64 i = 0; i < n; ++i) sh.document(buf + offs[rows[i]], buf + offs[rows[i] + 1]);
    out.resize(cols.size());
    for (size_t c = 0; c < cols.size(); ++c) {
        out[c].tags.swap(cc[c].tags);
        out[c].payload.swap(cc[c].payload);
        out[c].strings.swap(cc[c].strings);
    }
}

void Table::append_json(const char* buf, const i64* offsets, i64 ndocs, int threads) {
    if (sealed) N1_THROW(N1GPU_E_INVALID, "table is sealed");
    if (threads == -1) { append_json_device(buf, offsets, ndocs); return; }
    if (device_shredded) N1_THROW(N1GPU_E_INVALID, "table already holds device-shredded rows");
    for (auto& col : cols)
        if (col.dict_global || col.codes_are_ranks) N1_THROW(N1GPU_E_INVALID, "append after dictionary import / pre-shredded columns is not supported");
    if (ndocs < 0) N1_THROW(N1GPU_E_INVALID, "negative document count");
    double t0 = now_sec();
    appended = true;
    Trie trie;
    for (size_t c = 0; c < cols.size(); ++c) trie.add(cols[c].path, (int)c);
    int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if ((i64)nthreads > (ndocs + 4095) / 4096) nthreads = (int)std::max<i64>(1, (ndocs + 4095) / 4096);
    std::vector<std::vector<ChunkCol>> chunks(nthreads);
    std::vector<std::thread> pool;
    auto work = [&](int t) {
        i64 lo = ndocs * t / nthreads, hi = ndocs * (t + 1) / nthreads;
        auto& cc = chunks[t];
        cc.resize(cols.size());
        for (auto& c : cc) { c.payload.assign((size_t)(hi - lo), 0); c.tags.assign((size_t)(hi - lo), C_MISSING); }
        Shredder sh(trie, cc);
        for (i64 d = lo; d < hi; ++d) sh.document(buf + offsets[d], buf + offsets[d + 1]);
    };
    if (nthreads == 1) work(0);
    else {
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    // merge chunk-local string codes into the column's staging dictionary (parallel over columns)
    std::vector<std::unordered_map<std::string, u32>> colindex(cols.size());
    auto merge = [&](size_t c) {
        Column& col = cols[c];
        auto& index = colindex[c];
        for (u32 i = 0; i < col.local_strings.size(); ++i) index.emplace(col.local_strings[i], i);
        size_t base = col.payload.size();
        col.payload.resize(base + (size_t)ndocs);
        col.tags.resize(base + (size_t)ndocs);
        size_t at = base;
        for (int t = 0; t < nthreads; ++t) {
            ChunkCol& cc = chunks[t][c];
            std::vector<u32> remap(cc.strings.size());
            for (size_t i = 0; i < cc.strings.size(); ++i) {
                auto it = index.find(cc.strings[i]);
                if (it == index.end()) {
                    u32 id = (u32)col.local_strings.size();
                    col.local_strings.push_back(cc.strings[i]);
                    index.emplace(cc.strings[i], id);
                    remap[i] = id;
                } else remap[i] = it->second;
            }
            size_t n = cc.tags.size();
            memcpy(col.tags.data() + at, cc.tags.data(), n);
            for (size_t r = 0; r < n; ++r)
                col.payload[at + r] = cc.tags[r] == C_STRING ? (i64)remap[(size_t)cc.payload[r]] : cc.payload[r];
            at += n;
        }
    };
    if (cols.size() > 1 && nthreads > 1) {
        std::vector<std::thread> mp;
        std::vector<std::string> errs(cols.size());
        for (size_t c = 0; c < cols.size(); ++c)
            mp.emplace_back([&, c] { try { merge(c); } catch (const std::exception& e) { errs[c] = e.what(); } });
        for (auto& th : mp) th.join();
        for (auto& e : errs) if (!e.empty()) N1_THROW(N1GPU_E_INVALID, "%s", e.c_str());
    } else {
        for (size_t c = 0; c < cols.size(); ++c) merge(c);
    }
    nrows += ndocs;
    json_bytes += offsets[ndocs] - offsets[0];
    shred_sec += now_sec() - t0;
}

void Table::load_dir(const std::string& dir, int threads) {
    DIR* d = opendir(dir.c_str());
    if (!d) N1_THROW(N1GPU_E_IO, "cannot open keyspace directory %s", dir.c_str());
    std::vector<std::string> names;
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        if (e->d_type == DT_DIR) continue;
        if (e->d_type != DT_REG) {  // links / file systems without d_type: ask
            struct stat st;
            if (stat((dir + "/" + n).c_str(), &st) != 0 || S_ISDIR(st.st_mode)) continue;
        }
        names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());  // ioutil.ReadDir order
    // The primary index scans IDS, not files: every non-directory entry yields the id "name minus its last extension"
    // (documentPathToId, file.go:757-761), and Fetch reads <id>.json (file.go:346) - a missing file is silently no document
    // (file.go:319-322), an unreadable one is an error the request reports while it carries on.  So foo.txt is a document only
    // if foo.json exists (and a.json + a.txt are TWO rows of a.json), .DS_Store and editor backups are none.
    for (auto& n : names) {
        const size_t dot = n.rfind('.');
        if (dot != std::string::npos) n.resize(dot);
        n += ".json";
    }
    // The reference opens and reads one file per document, serially (file.go:732-743).  Here the reads of contiguous
    // ranges of the sorted names run on all cores - they are system calls on (mostly cached) small files - and the
    // documents are then laid end to end in primary-key order for the shredder.
    const size_t nfiles = names.size();
    int nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > (nfiles + 255) / 256) nthreads = (int)std::max<size_t>(1, (nfiles + 255) / 256);
    struct Part { std::string bytes; std::vector<i64> sizes; std::string err; };
    std::vector<Part> parts((size_t)nthreads);
    auto read_range = [&](int t) {
        Part& part = parts[(size_t)t];
        const size_t lo = nfiles * (size_t)t / (size_t)nthreads, hi = nfiles * (size_t)(t + 1) / (size_t)nthreads;
        part.sizes.reserve(hi - lo);
        std::string path;
        for (size_t i = lo; i < hi; ++i) {
            path.assign(dir).append("/").append(names[i]);
            const int fd = open(path.c_str(), O_RDONLY | O_CLOEXEC);
            if (fd < 0) continue;  // ENOENT: the id denotes no document; anything else: the reference reports and continues
            const size_t before = part.bytes.size();
            size_t cap = 4096;
            struct stat st;
            if (fstat(fd, &st) == 0 && st.st_size > 0) cap = (size_t)st.st_size + 1;
            bool failed = false;
            for (;;) {
                part.bytes.resize(part.bytes.size() + cap);
                const ssize_t got = read(fd, &part.bytes[part.bytes.size() - cap], cap);
                if (got < 0) { part.bytes.resize(before); failed = true; break; }  // (a directory named x.json, an I/O error): no document
                part.bytes.resize(part.bytes.size() - cap + (size_t)got);
                if (got == 0) break;
            }
            close(fd);
            if (!failed) part.sizes.push_back((i64)(part.bytes.size() - before));
        }
    };
    if (nthreads == 1) read_range(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(read_range, t);
        for (auto& th : pool) th.join();
    }
    size_t total = 0;
    for (auto& part : parts) { if (!part.err.empty()) N1_THROW(N1GPU_E_IO, "%s", part.err.c_str()); total += part.bytes.size(); }
    // laid end to end - in pinned memory when the device shredder takes the text (its H2D copy then needs no staging)
    PinnedBuf pin;
    std::string heap;
    char* buf = nullptr;
    if (threads < 0 && have_device()) { pin.ensure(total + 64); buf = pin.as<char>(); }
    else { heap.resize(total + 64); buf = &heap[0]; }
    std::vector<i64> offsets;
    offsets.reserve(nfiles + 1);
    offsets.push_back(0);
    std::vector<size_t> part_at(parts.size() + 1, 0);
    for (size_t i = 0; i < parts.size(); ++i) {
        part_at[i + 1] = part_at[i] + parts[i].bytes.size();
        for (i64 sz : parts[i].sizes) offsets.push_back(offsets.back() + sz);
    }
    {
        std::vector<std::thread> pool;
        for (size_t i = 0; i < parts.size(); ++i)
            pool.emplace_back([&, i] { if (!parts[i].bytes.empty()) memcpy(buf + part_at[i], parts[i].bytes.data(), parts[i].bytes.size()); std::string().swap(parts[i].bytes); });
        for (auto& th : pool) th.join();
    }
    append_json(buf, offsets.data(), (i64)offsets.size() - 1, threads);  // ids without a document were skipped
}

void Table::set_column(int c, int width, const void* payload, const u8* tags, i64 n, const char* blob,
                       const i64* offs, i64 ndict) {
    if (sealed) N1_THROW(N1GPU_E_INVALID, "table is sealed");
    if (c < 0 || c >= (int)cols.size()) N1_THROW(N1GPU_E_INVALID, "no such column %d", c);
    if (width != 8 && width != 4) N1_THROW(N1GPU_E_INVALID, "payload width must be 8 or 4");
    if (n < 0) N1_THROW(N1GPU_E_INVALID, "negative row count");
    if (appended) N1_THROW(N1GPU_E_INVALID, "cannot mix appended documents and pre-shredded columns");
    for (auto& other : cols)
        if (other.device_set) N1_THROW(N1GPU_E_INVALID, "cannot mix host and device columns in one table");
    for (auto& other : cols)
        if (&other != &cols[c] && !other.tags.empty() && (i64)other.tags.size() != n)
            N1_THROW(N1GPU_E_INVALID, "column lengths differ (%lld vs %lld)", (long long)other.tags.size(), (long long)n);
    Column& col = cols[c];
    col.payload.resize((size_t)n);
    col.tags.resize((size_t)n);
    if (width == 8) memcpy(col.payload.data(), payload, (size_t)n * 8);
    else for (i64 i = 0; i < n; ++i) col.payload[(size_t)i] = ((const u32*)payload)[i];
    if (tags) memcpy(col.tags.data(), tags, (size_t)n);
    else memset(col.tags.data(), width == 8 ? C_INT : C_STRING, (size_t)n);
    col.dict.clear();
    if (blob && offs) {
        for (i64 i = 0; i < ndict; ++i) col.dict.emplace_back(blob + offs[i], blob + offs[i + 1]);
        for (size_t i = 1; i < col.dict.size(); ++i)
            if (!(col.dict[i - 1] < col.dict[i])) N1_THROW(N1GPU_E_INVALID, "dictionary must be sorted bytewise and unique");
    }
    col.codes_are_ranks = true;
    nrows = n;
}

void Table::adopt_dictionary(int ci, std::vector<std::string>& global) {
    Column& c = cols[(size_t)ci];
    std::vector<u32> remap(c.dict.size());
    size_t g = 0;  // both sorted: one walk
    for (size_t i = 0; i < c.dict.size(); ++i) {
        while (g < global.size() && global[g] < c.dict[i]) ++g;
        if (g == global.size() || global[g] != c.dict[i]) N1_THROW(N1GPU_E_INVALID, "global dictionary lacks a local string");
        remap[i] = (u32)g;
    }
    bool identity = remap.size() == global.size();
    if (!identity || !remap.empty()) for (size_t i = 0; i < remap.size() && identity; ++i) identity = remap[i] == (u32)i;
    if (!identity && !remap.empty()) {
        if (c.d_payload.p && (device_shredded || c.device_set)) remap_ranks_device(ci, remap);
        else for (size_t r = 0; r < c.tags.size(); ++r) if (c.tags[r] == C_STRING) c.payload[r] = remap[(size_t)c.payload[r]];
    }
    c.dict.swap(global);
    c.dict_global = true;
    if (device_shredded || c.device_set) { c.stats.ndict = (i64)c.dict.size(); c.stats.empty_rank = (!c.dict.empty() && c.dict[0].empty()) ? 0 : -1; }
}

void Table::build_dictionary(int c) {
    Column& col = cols[c];
    if (col.codes_are_ranks) return;
    std::vector<u32> order(col.local_strings.size());
    for (u32 i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](u32 a, u32 b) { return col.local_strings[a] < col.local_strings[b]; });
    std::vector<u32> rank(order.size());
    col.dict.resize(order.size());
    for (u32 r = 0; r < order.size(); ++r) { rank[order[r]] = r; col.dict[r] = col.local_strings[order[r]]; }
    for (size_t i = 0; i < col.tags.size(); ++i)
        if (col.tags[i] == C_STRING) col.payload[i] = rank[(size_t)col.payload[i]];
    col.local_strings.clear();
    col.local_strings.shrink_to_fit();
    col.codes_are_ranks = true;
}

int Table::scan_bytes(int c) const {
    const Column& col = cols[c];
    return col.width + (col.stats.uniform_tag() ? 0 : 1);
}

void Table::seal() {
    if (sealed) return;
    if (device_shredded) {  // columns, dictionaries and statistics already final in HBM
        bool any_set = false;
        for (auto& col : cols) any_set = any_set || col.device_set;
        if (any_set)
            for (auto& col : cols)
                if (!col.device_set) N1_THROW(N1GPU_E_INVALID, "column %s was never set", join_path(col.path, '.').c_str());
        sealed = true;
        return;
    }
    double t0 = now_sec();
    for (auto& col : cols)
        if ((i64)col.tags.size() != nrows) N1_THROW(N1GPU_E_INVALID, "column %s has %lld rows, table has %lld",
                                                    join_path(col.path, '.').c_str(), (long long)col.tags.size(), (long long)nrows);
    auto finish = [&](size_t c) {
        Column& col = cols[c];
        build_dictionary((int)c);
        ColumnStats st;
        for (size_t i = 0; i < col.tags.size(); ++i) {
            u8 t = col.tags[i];
            if (t == C_FLOAT) {  // pre-shredded input may hold integral floats: canonicalise (NewValue)
                double d; memcpy(&d, &col.payload[i], 8);
                if (f_is_int(d)) { col.tags[i] = t = C_INT; col.payload[i] = go_i64(d); }
                else {
                    if (!st.has_float) { st.flt_min = st.flt_max = d; st.has_float = true; }
                    else { st.flt_min = std::min(st.flt_min, d); st.flt_max = std::max(st.flt_max, d); }
                }
            }
            if (t == C_INT) {
                i64 v = col.payload[i];
                if (!st.has_int) { st.int_min = st.int_max = v; st.has_int = true; }
                else { st.int_min = std::min(st.int_min, v); st.int_max = std::max(st.int_max, v); }
            }
            if (t > C_OTHER) N1_THROW(N1GPU_E_INVALID, "bad class byte %d", (int)t);
            st.class_mask |= bit(t);
            st.absent_rows += t <= C_NULL;
        }
        st.ndict = (i64)col.dict.size();
        st.empty_rank = (!col.dict.empty() && col.dict[0].empty()) ? 0 : -1;
        if (!col.stats_forced) col.stats = st;
        u32 m = col.stats.class_mask;
        col.width = (m & M_NUM) ? 8 : ((m & bit(C_STRING)) ? 4 : 0);
    };
    {
        std::vector<std::thread> pool;
        std::vector<std::string> errs(cols.size());
        for (size_t c = 0; c < cols.size(); ++c)
            pool.emplace_back([&, c] { try { finish(c); } catch (const std::exception& e) { errs[c] = e.what(); } });
        for (auto& th : pool) th.join();
        for (auto& e : errs) if (!e.empty()) N1_THROW(N1GPU_E_INVALID, "%s", e.c_str());
    }
    shred_sec += now_sec() - t0;
    if (!segment_out.empty()) write_segment();
    t0 = now_sec();
    if (!have_device()) {
        // build / CPU-test container: dictionaries and statistics only; queries can still be compiled
        // (NVRTC -> sm_100a cubin) but never executed - there is no CPU fallback.
        for (auto& col : cols) { std::vector<i64>().swap(col.payload); std::vector<u8>().swap(col.tags); }
        sealed = true;
        return;
    }
    i64 pad = padded_rows();
    if (pad == 0) pad = ROW_PAD;
    for (auto& col : cols) {
        col.d_tags.alloc((size_t)pad);
        CK(cudaMemset(col.d_tags.p, C_MISSING, (size_t)pad));
        if (nrows) CK(cudaMemcpy(col.d_tags.p, col.tags.data(), (size_t)nrows, cudaMemcpyHostToDevice));
        if (col.width == 8) {
            col.d_payload.alloc((size_t)pad * 8);
            CK(cudaMemset(col.d_payload.p, 0, (size_t)pad * 8));
            if (nrows) CK(cudaMemcpy(col.d_payload.p, col.payload.data(), (size_t)nrows * 8, cudaMemcpyHostToDevice));
        } else if (col.width == 4) {
            std::vector<u32> narrow((size_t)nrows);
            for (i64 i = 0; i < nrows; ++i) narrow[(size_t)i] = (u32)col.payload[(size_t)i];
            col.d_payload.alloc((size_t)pad * 4);
            CK(cudaMemset(col.d_payload.p, 0, (size_t)pad * 4));
            if (nrows) CK(cudaMemcpy(col.d_payload.p, narrow.data(), (size_t)nrows * 4, cudaMemcpyHostToDevice));
        } else {
            col.d_payload.alloc(256);
        }
        std::vector<i64>().swap(col.payload);
        std::vector<u8>().swap(col.tags);
    }
    CK(cudaDeviceSynchronize());
    upload_sec += now_sec() - t0;
    sealed = true;
}

}  // namespace n1
