// common.hpp — shared host-side definitions for libn1gpu.so
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/n1gpu.h"

namespace n1 {

typedef long long i64;           // same types as the device library (n1ql_device.cuh)
typedef unsigned long long u64;
typedef uint32_t u32;
typedef uint8_t u8;

enum : int { C_MISSING = 0, C_NULL = 1, C_FALSE = 2, C_TRUE = 3, C_INT = 4, C_FLOAT = 5, C_STRING = 6, C_OTHER = 7 };
enum : int { OP_ADD_U64 = 0, OP_ADD_F64 = 1, OP_MIN_I64 = 2, OP_MAX_I64 = 3, OP_MIN_U64 = 4, OP_MAX_U64 = 5, OP_OR_U64 = 6 };

inline u32 bit(int c) { return 1u << c; }
const u32 M_NUM = (1u << C_INT) | (1u << C_FLOAT);
const u32 M_BOOL = (1u << C_FALSE) | (1u << C_TRUE);
const u32 M_ALL_SCALAR = 0x7f;

// An error carrying one of the N1GPU_E_* codes; the ABI layer turns it into a status + message.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline std::string strf(const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return buf;
}

#define N1_THROW(code, ...) throw ::n1::Error((code), ::n1::strf(__VA_ARGS__))
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            N1_THROW(N1GPU_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

extern std::atomic<u64> g_launches;  // every kernel launch made by this library

// A host value (result side): class + payload; strings own their bytes.
struct HValue {
    u8 cls = C_MISSING;
    i64 bits = 0;
    std::string s;
    static HValue missing() { return HValue(); }
    static HValue null() { HValue v; v.cls = C_NULL; return v; }
    static HValue boolean(bool b) { HValue v; v.cls = b ? C_TRUE : C_FALSE; return v; }
    static HValue integer(i64 x) { HValue v; v.cls = C_INT; v.bits = x; return v; }
    static HValue flt(double d) { HValue v; v.cls = C_FLOAT; memcpy(&v.bits, &d, 8); return v; }
    static HValue str(const std::string& s) { HValue v; v.cls = C_STRING; v.s = s; return v; }
    double f() const { double d; memcpy(&d, &bits, 8); return d; }
    double num() const { return cls == C_INT ? (double)bits : f(); }
};

// Go's int64(float64) on amd64 and value.IsInt (value/integer.go:354-356)
inline i64 go_i64(double d) {
    if (!(d >= -9223372036854775808.0 && d < 9223372036854775808.0)) return INT64_MIN;
    return (i64)d;
}
inline bool f_is_int(double d) { return d == (double)go_i64(d); }
// value.NewValue(float64) (value/value.go:377-382)
inline HValue new_num(double d) { return f_is_int(d) ? HValue::integer(go_i64(d)) : HValue::flt(d); }

// caching allocators (mem.cpp)
void* dev_alloc(size_t n, size_t* actual);
void dev_free(void* p, size_t n);
void dev_pool_trim();
void* pin_alloc(size_t n, size_t* actual);
void pin_free(void* p, size_t n);

// Device buffer (pooled cudaMalloc) with RAII.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; return *this; }
    ~DevBuf() { release(); }
    void alloc(size_t n) {
        release();
        p = dev_alloc(n, &bytes);  // bytes = the (rounded-up) size actually held
    }
    void ensure(size_t n) { if (n > bytes) alloc(n); }
    void release() { if (p) dev_free(p, bytes); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    PinnedBuf() {}
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { if (p) pin_free(p, bytes); }
    void ensure(size_t n) {
        if (n <= bytes) return;
        if (p) pin_free(p, bytes);
        p = nullptr;
        p = pin_alloc(n, &bytes);
    }
    template <class T> T* as() const { return (T*)p; }
};

inline u64 mix64(u64 x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31; return x;
}
inline int bits_for(u64 n_values) {  // bits to represent values 0..n_values-1
    int b = 0;
    while (b < 64 && ((u64)1 << b) < n_values) ++b;
    return b;
}

double now_sec();
bool have_device();  // a CUDA device is visible to this process

}  // namespace n1
