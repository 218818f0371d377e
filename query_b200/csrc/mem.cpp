// mem.cpp — process-wide caching allocators for HBM and pinned host memory.  A keyspace reload or a per-request
// query state allocates and frees hundreds of MB; cudaMalloc/cudaFree/cudaMallocHost are synchronous and cost
// milliseconds, so freed blocks are kept in exact-size bins (sizes are rounded up so that repeated requests of
// the same shape hit) and reused.  180 GB of HBM makes a generous cache harmless; it is capped anyway.
#include <map>
#include <mutex>

#include "common.hpp"

namespace n1 {

namespace {
struct Pool {
    std::mutex mu;
    std::map<size_t, std::vector<void*>> bins;
    size_t cached = 0;
    size_t cap;
    explicit Pool(size_t c) : cap(c) {}
};
Pool g_dev(48ull << 30), g_pin(8ull << 30);

size_t round_size(size_t n) {
    if (n < 256) return 256;
    if (n < (1u << 20)) return (n + 4095) & ~(size_t)4095;
    return (n + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
}
}  // namespace

void* dev_alloc(size_t n, size_t* actual) {
    n = round_size(n);
    *actual = n;
    {
        std::lock_guard<std::mutex> lk(g_dev.mu);
        auto it = g_dev.bins.find(n);
        if (it != g_dev.bins.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            g_dev.cached -= n;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
        cudaGetLastError();
        dev_pool_trim();
        e = cudaMalloc(&p, n);
        if (e != cudaSuccess) N1_THROW(N1GPU_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
    }
    return p;
}

void dev_free(void* p, size_t n) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_dev.mu);
        if (g_dev.cached + n <= g_dev.cap) {
            g_dev.bins[n].push_back(p);
            g_dev.cached += n;
            return;
        }
    }
    cudaFree(p);
}

void dev_pool_trim() {
    std::lock_guard<std::mutex> lk(g_dev.mu);
    for (auto& b : g_dev.bins) for (void* p : b.second) cudaFree(p);
    g_dev.bins.clear();
    g_dev.cached = 0;
}

void* pin_alloc(size_t n, size_t* actual) {
    n = round_size(n);
    *actual = n;
    {
        std::lock_guard<std::mutex> lk(g_pin.mu);
        auto it = g_pin.bins.find(n);
        if (it != g_pin.bins.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            g_pin.cached -= n;
            return p;
        }
    }
    void* p = nullptr;
    CK(cudaMallocHost(&p, n));
    return p;
}

void pin_free(void* p, size_t n) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin.mu);
        if (g_pin.cached + n <= g_pin.cap) {
            g_pin.bins[n].push_back(p);
            g_pin.cached += n;
            return;
        }
    }
    cudaFreeHost(p);
}

}  // namespace n1
