// atomic_probe.cu — throughput of the accumulator-update primitives a GROUP BY kernel can choose between, measured
// at full occupancy on the device it runs on (not part of the library; design input for codegen.cpp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/atomic_probe tools/atomic_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ u32 mix(u32 x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// index distribution: skew == 0 uniform over n; skew == 1 log-uniform (Zipf s=1) over n
__device__ __forceinline__ u32 draw(u32 h, u32 n, int skew) {
    if (!skew) return h % n;
    const float u = (h >> 8) * (1.0f / 16777216.0f);
    u32 r = (u32)(__expf(u * __logf((float)n))) - 1;
    return r < n ? r : n - 1;
}

enum { S_ADD32, S_ADD64, S_MIN32, S_MIN64, S_RMW64, S_RMW32, S_MATCH, S_CHECKMIN64, G_RED64, G_RED32, G_MIN64, G_ATOM64, G_CHECKMIN64, G_LD64, G_RED64_G4, G_RED64_G8, G_LD256, NKIND };
const char* kname[NKIND] = {"smem atomicAdd u32", "smem atomicAdd u64 (CAS loop)", "smem atomicMin s32", "smem atomicMin s64 (CAS loop)",
                            "smem plain RMW u64 (racy)", "smem plain RMW u32 (racy)", "match_any + plain RMW u64", "smem load+compare, atomicMin s64 if smaller",
                            "global red.add u64", "global red.add u32", "global red.min s64", "global atom.add u64 (returning)",
                            "global load+compare, red.min s64 if smaller", "global ld u64 (reference)", "global red.add u64, 4 lanes on 4 adjacent words", "global red.add u64, 8 lanes on 8 adjacent words",
                            "global ld 256-bit (one 32 B sector per lane)"};

template <int KIND>
__global__ void __launch_bounds__(256) probe(u64* g, u32 n, int skew, int iters, int smem_words, u64* sink) {
    extern __shared__ u64 s[];
    u32* s32 = (u32*)s;
    for (int i = threadIdx.x; i < smem_words; i += 256) s[i] = (KIND == S_MIN64 || KIND == S_CHECKMIN64) ? ~0ull >> 1 : 0;
    __syncthreads();
    u32 h = mix(blockIdx.x * 256 + threadIdx.x + 1);
    u64 acc = 0;
    const u32 sn = KIND == S_ADD32 || KIND == S_MIN32 || KIND == S_RMW32 ? (u32)smem_words * 2 : (u32)smem_words;
    for (int it = 0; it < iters; ++it) {
        h = mix(h + it);
        const u32 v = h >> 12;
        if (KIND == S_ADD32) atomicAdd(&s32[draw(h, sn < n ? sn : n, skew)], v);
        if (KIND == S_ADD64) atomicAdd(&s[draw(h, sn < n ? sn : n, skew)], (u64)v);
        if (KIND == S_MIN32) atomicMin((int*)&s32[draw(h, sn < n ? sn : n, skew)], (int)v);
        if (KIND == S_MIN64) atomicMin((long long*)&s[draw(h, sn < n ? sn : n, skew)], (long long)v);
        if (KIND == S_CHECKMIN64) { long long* p = (long long*)&s[draw(h, sn < n ? sn : n, skew)]; if ((long long)v < *(volatile long long*)p) atomicMin(p, (long long)v); }
        if (KIND == S_RMW64) { volatile u64* p = &s[draw(h, sn < n ? sn : n, skew)]; *p = *p + v; }
        if (KIND == S_RMW32) { volatile u32* p = &s32[draw(h, sn < n ? sn : n, skew)]; *p = *p + v; }
        if (KIND == S_MATCH) {
            const u32 idx = draw(h, sn < n ? sn : n, skew);
            const unsigned peers = __match_any_sync(0xffffffffu, idx);
            // leader adds the (here: popcount-scaled) value once: stands for an in-warp combine before the update
            if ((peers & ((1u << (threadIdx.x & 31)) - 1)) == 0) { volatile u64* p = &s[idx]; *p = *p + (u64)v * __popc(peers); }
            __syncwarp();
        }
        if (KIND == G_RED64) atomicAdd(&g[draw(h, n, skew)], (u64)v);
        if (KIND == G_RED32) atomicAdd(&((u32*)g)[draw(h, n, skew)], v);
        if (KIND == G_MIN64) atomicMin((long long*)&g[draw(h, n, skew)], (long long)v);
        if (KIND == G_ATOM64) acc += atomicAdd(&g[draw(h, n, skew)], (u64)v);
        if (KIND == G_CHECKMIN64) { long long* p = (long long*)&g[draw(h, n, skew)]; if ((long long)v < __ldcg(p)) atomicMin(p, (long long)v); }
        if (KIND == G_LD64) acc += __ldcg(&g[draw(h, n, skew)]);
        if (KIND == G_RED64_G4 || KIND == G_RED64_G8) {  // a group of lanes updates adjacent words of ONE random slot
            const int G = KIND == G_RED64_G4 ? 4 : 8;
            const u32 lane = threadIdx.x & 31;
            const u32 slot = __shfl_sync(0xffffffffu, draw(h, n / G, skew), lane & ~(G - 1));
            atomicAdd(&g[(size_t)slot * G + (lane & (G - 1))], (u64)v);
        }
        if (KIND == G_LD256) {
            const ulonglong4 x = *(const ulonglong4*)&g[(size_t)draw(h, n / 4, skew) * 4];
            acc += x.x + x.y + x.z + x.w;
        }
    }
    __syncthreads();
    if (KIND < G_RED64) for (int i = threadIdx.x; i < smem_words; i += 256) acc += s[i];
    if (acc == 0x1234567u) sink[0] = acc;
}

template <int KIND>
void run(u64* g, u32 n, int skew, int smem_words, u64* sink, int sms) {
    const int iters = 2048, grid = sms * 8;
    const size_t smem = (size_t)smem_words * 8;
    CK(cudaFuncSetAttribute(probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    if (KIND == G_MIN64 || KIND == G_CHECKMIN64) CK(cudaMemset(g, 0x7f, (size_t)n * 8));
    probe<KIND><<<grid, 256, smem>>>(g, n, skew, 64, smem_words, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    probe<KIND><<<grid, 256, smem>>>(g, n, skew, iters, smem_words, sink);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double ops = (double)grid * 256 * iters;
    printf("%-48s n=%-8u %-8s smem=%3zuKB  %8.1f Gops/s  (%.3f cyc/lane/SM @1.9GHz)\n", kname[KIND], n, skew ? "zipf" : "uniform", smem / 1024,
           ops / ms / 1e6, 1.9e9 * sms / (ops / ms * 1e3));
}

int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    u64 *g, *sink;
    const u32 NMAX = 8u << 20;
    CK(cudaMalloc(&g, (size_t)NMAX * 8));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(g, 0, (size_t)NMAX * 8));
    const int SW = 3072;  // 24 KB of shared memory per block: 8 blocks per SM
    for (int skew = 0; skew < 2; ++skew) {
        const u32 n = 100000;
        run<S_ADD32>(g, n, skew, SW, sink, sms);
        run<S_ADD64>(g, n, skew, SW, sink, sms);
        run<S_MIN32>(g, n, skew, SW, sink, sms);
        run<S_MIN64>(g, n, skew, SW, sink, sms);
        run<S_CHECKMIN64>(g, n, skew, SW, sink, sms);
        run<S_RMW64>(g, n, skew, SW, sink, sms);
        run<S_RMW32>(g, n, skew, SW, sink, sms);
        run<S_MATCH>(g, n, skew, SW, sink, sms);
        for (u32 gn : {100000u, 1000000u, 8u << 20}) {
            run<G_RED64>(g, gn, skew, 16, sink, sms);
            run<G_RED32>(g, gn, skew, 16, sink, sms);
            run<G_MIN64>(g, gn, skew, 16, sink, sms);
            run<G_CHECKMIN64>(g, gn, skew, 16, sink, sms);
            run<G_ATOM64>(g, gn, skew, 16, sink, sms);
            run<G_LD64>(g, gn, skew, 16, sink, sms);
            run<G_RED64_G4>(g, gn, skew, 16, sink, sms);
            run<G_RED64_G8>(g, gn, skew, 16, sink, sms);
            run<G_LD256>(g, gn, skew, 16, sink, sms);
        }
    }
    return 0;
}
