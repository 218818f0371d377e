// mem.cpp — process-wide caching allocators for HBM (per device) and pinned host memory.  A keyspace reload or a per-request
// query state allocates and frees hundreds of MB; cudaMalloc/cudaFree/cudaMallocHost are synchronous and cost
// milliseconds, so freed blocks are kept in exact-size bins (sizes are rounded up so that repeated requests of
// the same shape hit) and reused.  180 GB of HBM makes a generous cache harmless; it is capped anyway.
#include <map>
#include <mutex>
#include <unordered_map>

#include "common.hpp"

namespace n1 {

namespace {
// bins are keyed by (device, size): a block of HBM belongs to the device that was current when it was allocated, and a
// process that drives several devices (n1gpu_init(other) / cudaSetDevice) must never be handed another GPU's pointer.
// Pinned host memory is device-agnostic: device -1.
struct Pool {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void*>> bins;
    std::unordered_map<void*, int> owner;  // live + cached HBM blocks -> device
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
    int dev = 0;
    CK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_dev.mu);
        auto it = g_dev.bins.find({dev, n});
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
    std::lock_guard<std::mutex> lk(g_dev.mu);
    g_dev.owner[p] = dev;
    return p;
}

// Callers free a block only after the work that uses it has completed (Query::~Query synchronises its stream first; a
// table is sealed with a device synchronisation and must outlive its queries, n1gpu.h), so a cached block can be handed
// out again at once.
void dev_free(void* p, size_t n) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_dev.mu);
        auto o = g_dev.owner.find(p);
        const int dev = o == g_dev.owner.end() ? 0 : o->second;
        if (g_dev.cached + n <= g_dev.cap) {
            g_dev.bins[{dev, n}].push_back(p);
            g_dev.cached += n;
            return;
        }
        if (o != g_dev.owner.end()) g_dev.owner.erase(o);
    }
    cudaFree(p);
}

void dev_pool_trim() {
    std::lock_guard<std::mutex> lk(g_dev.mu);
    for (auto& b : g_dev.bins) for (void* p : b.second) { cudaFree(p); g_dev.owner.erase(p); }
    g_dev.bins.clear();
    g_dev.cached = 0;
}

static void pin_pool_trim() {
    std::lock_guard<std::mutex> lk(g_pin.mu);
    for (auto& b : g_pin.bins) for (void* p : b.second) cudaFreeHost(p);
    g_pin.bins.clear();
    g_pin.cached = 0;
}

void* pin_alloc(size_t n, size_t* actual) {
    n = round_size(n);
    *actual = n;
    {
        std::lock_guard<std::mutex> lk(g_pin.mu);
        auto it = g_pin.bins.find({-1, n});
        if (it != g_pin.bins.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            g_pin.cached -= n;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, n);
    if (e != cudaSuccess) {  // page-locked memory is a scarce resource: drop the cached blocks and retry once
        cudaGetLastError();
        pin_pool_trim();
        e = cudaMallocHost(&p, n);
        if (e != cudaSuccess) N1_THROW(N1GPU_E_NOMEM, "cudaMallocHost(%zu bytes) failed: %s", n, cudaGetErrorString(e));
    }
    return p;
}

void pin_free(void* p, size_t n) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin.mu);
        if (g_pin.cached + n <= g_pin.cap) {
            g_pin.bins[{-1, n}].push_back(p);
            g_pin.cached += n;
            return;
        }
    }
    cudaFreeHost(p);
}

}  // namespace n1
