// jit.hpp — NVRTC compilation of generated scan kernels for sm_100a and launch through the CUDA driver
// API (entry points resolved with cudaGetDriverEntryPoint, so libn1gpu.so has no link-time dependency on
// libcuda and loads on a machine without a GPU).
#pragma once
#include <memory>
#include <string>

#include "common.hpp"

namespace n1 {

struct JitKernel {
    void* module = nullptr;    // CUmodule
    void* function = nullptr;  // CUfunction
    int regs = 0;
    int static_smem = 0;
    int dyn_smem = 0;
    int block = 256;            // threads per block the source was generated for (NQ_BLOCK)
    int max_blocks_per_sm = 0;  // occupancy at `block` threads
    std::string cubin;
    ~JitKernel();
};

// Compiles `source` (which #includes "n1ql_device.cuh") to an sm_100a cubin.  Works without a GPU.
std::string jit_compile_cubin(const std::string& source, std::string* log);
// Compiles (cached per process by source text) and loads the kernel `nq_scan` on the current device.
std::shared_ptr<JitKernel> jit_load(const std::string& source, int dyn_smem = 0, int block = 256);
// Launches nq_scan<<<grid, k.block, k.dyn_smem, stream>>>(params) where params is a by-value struct of `bytes` bytes.
void jit_launch(const JitKernel& k, int grid, cudaStream_t stream, void* params, size_t bytes, bool pdl = false);

// kernels compiled by NVRTC / requests served from the in-process caches (by source text) since load
void jit_stats(unsigned long long* compiled, unsigned long long* reused);

int device_sm_count();

}  // namespace n1
