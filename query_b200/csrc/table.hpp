// table.hpp — a shredded keyspace: typed columns + per-column sorted dictionaries, resident in HBM.
#pragma once
#include <map>
#include <functional>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.hpp"

namespace n1 {

const i64 ROW_PAD = 4096;  // device arrays are padded to a multiple of this many rows (vector loads never fault)

// A column's dictionary: a vector of strings that results can keep alive after the table is gone (a result refers to its
// strings as (column, rank) and resolves them on demand - a million-group result copies no string at finalisation).
// The table is immutable once sealed, so sharing needs no copy-on-write.
class SharedDict {
    std::shared_ptr<std::vector<std::string>> p = std::make_shared<std::vector<std::string>>();
public:
    typedef std::vector<std::string>::const_iterator const_iterator;
    size_t size() const { return p->size(); }
    bool empty() const { return p->empty(); }
    const std::string& operator[](size_t i) const { return (*p)[i]; }
    std::string& operator[](size_t i) { return (*p)[i]; }
    const_iterator begin() const { return p->begin(); }
    const_iterator end() const { return p->end(); }
    void clear() { p = std::make_shared<std::vector<std::string>>(); }
    void reserve(size_t n) { p->reserve(n); }
    void resize(size_t n) { p->resize(n); }
    template <class... A> void emplace_back(A&&... a) { p->emplace_back(std::forward<A>(a)...); }
    void push_back(const std::string& s) { p->push_back(s); }
    void swap(std::vector<std::string>& o) { auto n = std::make_shared<std::vector<std::string>>(); n->swap(o); o.swap(*p); p = n; }
    SharedDict& operator=(std::vector<std::string>&& o) { p = std::make_shared<std::vector<std::string>>(std::move(o)); return *this; }
    bool operator!=(const SharedDict& o) const { return *p != *o.p; }
    bool operator==(const SharedDict& o) const { return *p == *o.p; }
    const std::vector<std::string>& vec() const { return *p; }
    std::shared_ptr<const std::vector<std::string>> share() const { return p; }
};

struct ColumnStats {
    u32 class_mask = 0;  // classes that occur (bit per C_*)
    bool has_int = false;
    i64 int_min = 0, int_max = 0;
    bool has_float = false;
    double flt_min = 0, flt_max = 0;
    i64 ndict = 0;
    i64 empty_rank = -1;  // rank of "" in the dictionary, -1 if absent
    i64 absent_rows = 0;  // rows whose class is MISSING or NULL (decides which way a COUNT(x) counter counts)
    bool uniform_tag() const { return class_mask != 0 && (class_mask & (class_mask - 1)) == 0; }
};

struct Column {
    std::vector<std::string> path;       // field names below the document root
    // host staging (dropped after seal unless keep_host)
    std::vector<i64> payload;            // 8 bytes per row while staging
    std::vector<u8> tags;
    SharedDict dict;                     // sorted, unique (after seal / import); results keep it alive (share())
    bool dict_global = false;            // dictionary imported (multi-GPU): do not rebuild at seal
    std::vector<std::string> local_strings;  // staging: distinct strings, codes index this until seal
    bool codes_are_ranks = false;        // set_column supplied ranks into `dict` already
    ColumnStats stats;
    bool stats_forced = false;
    int width = 8;                       // device payload width: 8, 4 or 0
    DevBuf d_payload, d_tags;
    bool device_set = false;             // filled by set_column_device
};

struct Table {
    std::vector<Column> cols;
    i64 nrows = 0;
    i64 global_rows = 0;  // rows of the whole keyspace over all partitions (0 = not declared)
    bool sealed = false;
    bool appended = false;
    double shred_sec = 0, upload_sec = 0;
    i64 json_bytes = 0;

    // persistent segment (segment.cpp): when set, seal() writes the shredded columns to this file under this source tag
    std::string segment_out, segment_source;
    void write_segment() const;
    bool load_segment(const std::string& file, const std::string& source);
    void load_ndjson(const std::string& file, int threads);  // one document per line; threads as for load_dir

    bool device_shredded = false;  // columns were produced in HBM by shred.cu (no host staging exists)

    // threads >= 0: host threads (0 = all cores); threads == -1: the device shredder (shred.cu)
    void append_json_device(const char* buf, const i64* offsets, i64 ndocs);
    // one document per line; offsets computed on the device.  `fill(lo, hi)` (optional) makes bytes [lo, hi) of text valid just
    // before they are copied: a file is read chunk by chunk while the chunks before it cross PCIe
    void append_ndjson_device(const char* text, i64 size, const std::function<void(i64, i64)>& fill = nullptr);
    void append_text_device(const char* buf, const i64* offsets, i64 ndocs, i64 text_size, const std::function<void(i64, i64)>& fill);
    int add_column(const std::string& path);
    int find_column(const std::string& path) const;
    void append_json(const char* buf, const i64* offsets, i64 ndocs, int threads);
    void load_dir(const std::string& dir, int threads);
    void set_column(int col, int width, const void* payload, const u8* tags, i64 nrows, const char* blob,
                    const i64* offs, i64 ndict);
    // The same for a column that already lives in device memory (caller-owned device pointers, copied device to
    // device): no host staging at all; statistics are computed by a kernel.  Every column of the table must then
    // be set this way.
    void set_column_device(int col, int width, const void* dev_payload, const u8* dev_tags, i64 nrows, const char* blob,
                           const i64* offs, i64 ndict);
    void build_dictionary(int col);  // local strings -> sorted dict, codes -> ranks
    // Replaces the column's (sorted) dictionary by a sorted superset and remaps the rows' ranks - staged on the host, or
    // already in HBM (device shredder, set_column_device).  Multi-GPU: the dictionary every partition agrees on.
    void adopt_dictionary(int col, std::vector<std::string>& global);
    void remap_ranks_device(int col, const std::vector<u32>& remap);  // rows of a column resident in HBM: rank -> remap[rank]
    void seal();
    int scan_bytes(int col) const;
    i64 padded_rows() const { return (nrows + ROW_PAD - 1) / ROW_PAD * ROW_PAD; }
};

// The host shredder applied to selected documents (the device shredder's fix-up rows).  Per column: tags[n],
// payload[n] (STRING: index into strings).
struct HostShredOut {
    std::vector<u8> tags;
    std::vector<i64> payload;
    std::vector<std::string> strings;
};
void host_shred_docs(const std::vector<Column>& cols, const char* buf, const i64* offs, const i64* rows, i64 n,
                     std::vector<HostShredOut>& out);

std::vector<std::string> split_path(const std::string& path);
std::string join_path(const std::vector<std::string>& p, char sep);

}  // namespace n1
