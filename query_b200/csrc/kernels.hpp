// kernels.hpp — host launchers of the static kernels in kernels.cu
#pragma once
#include "common.hpp"

namespace n1 {

struct OpsArr {
    int n;
    int op[64];
};

// one DISTINCT aggregate: where its result words live and how its value component is packed in a set entry
struct DistinctDesc {
    int sid;           // entry set this aggregate reads (aggregates over the same operand share one)
    int numbers_only;  // COUNTN / SUM / AVG DISTINCT: only numeric entries count (COUNT: every entry)
    int w_cnt, w_ilo, w_ihi, w_neg, w_fsum, w_nflt;  // word indices (-1 = not needed)
    int cbits, pbits, biased;
    int nfree;         // >= 0: offset packing (PackComp::nfree)
    i64 bias;
    int classes[8];
};
struct DistinctDescs {
    int n;
    DistinctDesc d[16];
};
void launch_distinct_finalize(const u64* set_keys, u64 set_cap, int set128, int abits, int key_bits, int kw, const u64* keys, u64 cap,
                              u64* acc, const DistinctDescs& D, cudaStream_t s);
void launch_merge_mailbox(const u64* mail, int nranks, u64 slot_base, u64 stride, u64 words, u64 seq, u64 cap, const OpsArr& ops,
                          u64* out_dev, u64* out_host, int* status, cudaStream_t s);
void launch_merge_words(const u64* all, int nranks, u64 cap, const OpsArr& ops, u64* out_dev, u64* out_host, cudaStream_t s);
void launch_init_words(u64* acc, u64 cap, const OpsArr& ops, cudaStream_t s);
void launch_fill_u64(u64* p, u64 n, u64 v, cudaStream_t s);
void launch_count_owners(int kw, const u64* keys, const u64* acc, u64 cap, int nranks, int gk_pos, int gk_bits,
                         unsigned long long* counts, cudaStream_t s);
void launch_export_records(int kw, const u64* keys, const u64* acc, u64 cap, int W, int nranks, int gk_pos, int gk_bits,
                           unsigned long long* cursor, u64* out, u64 out_cap, cudaStream_t s);
void launch_merge_records(int kw, u64* keys, u64* acc, u64 cap, const OpsArr& ops, const u64* recs, u64 n, int* status, cudaStream_t s);

}  // namespace n1
