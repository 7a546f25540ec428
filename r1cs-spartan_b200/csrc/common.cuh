// common.cuh -- error plumbing, device buffers and launch geometry shared by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include "field.cuh"
#include "ec.cuh"

// Status codes come from the C ABI header (SB_EINVAL mirrors the reference's Error::InvalidArgument,
// src/error.rs:5-14).
#include "../../include/spartan_b200.h"

struct SbError : std::runtime_error {
    int code;
    SbError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SB_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            throw SbError(_e == cudaErrorMemoryAllocation ? SB_ENOMEM : SB_ECUDA,                      \
                          std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                              std::to_string(__LINE__) + ")");                                         \
    } while (0)

#define SB_REQUIRE(cond, msg)                        \
    do {                                             \
        if (!(cond)) throw SbError(SB_EINVAL, msg);  \
    } while (0)

// B200: 148 SMs.  Streaming kernels run as persistent grids of SB_SMS * k CTAs with grid-stride loops.
constexpr int SB_SMS = 148;

// Counter of kernels this library launched (bench.py reports it as gpu_launches), host<->device byte
// counters, and the optional per-kernel CUDA-event profiler (sb_prof_enable / sb_prof_report).
#include <atomic>
extern std::atomic<unsigned long long> g_sb_launches;           // (atomic: a multi-GPU context drives one host thread per GPU)
extern std::atomic<unsigned long long> g_sb_h2d_bytes, g_sb_d2h_bytes;
extern bool g_sb_prof_on;
extern int g_sb_prof_tag;
void sb_prof_begin(const char* name, cudaStream_t stream);
void sb_prof_end(cudaStream_t stream);
#if defined(SB_EMUL)
// tests/emul (CUDA-on-CPU test shim, never part of the product build): the launch runs the kernel's threads on the host
#define SB_KERNEL_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })
#define SB_DYN_SMEM(name) unsigned char* name = emu::g_block->dyn_smem
#else
#define SB_KERNEL_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#define SB_LAUNCH_NAMED(name, kernel, grid, block, smem, stream, ...)                \
    do {                                                                             \
        if (g_sb_prof_on) sb_prof_begin(name, stream);                               \
        SB_KERNEL_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__);            \
        if (g_sb_prof_on) sb_prof_end(stream);                                       \
        g_sb_launches++;                                                             \
        SB_CUDA(cudaGetLastError());                                                 \
    } while (0)
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...) SB_LAUNCH_NAMED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)
// kernel names for the two instantiations of the group templates
#define SB_KNAME(F, base) (sizeof(F) == sizeof(Fq) ? base "<Fq>" : base "<Fq2>")

// RAII device allocation on a stream-ordered pool.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(size_t count, cudaStream_t stream) { alloc(count, stream); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(size_t count, cudaStream_t stream) {
        release();
        n = count; s = stream;
        if (count) SB_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), stream));
    }
    void release() {
        if (p) { cudaFreeAsync(p, s); p = nullptr; }
        n = 0;
    }
    ~DevBuf() { release(); }
    T* get() const { return p; }
    size_t bytes() const { return n * sizeof(T); }
};

// Pinned host staging buffer.
template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    PinnedBuf() {}
    explicit PinnedBuf(size_t count) { alloc(count); }
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    void alloc(size_t count) {
        release(); n = count;
        if (count) SB_CUDA(cudaHostAlloc((void**)&p, count * sizeof(T), cudaHostAllocDefault));
    }
    void release() { if (p) { cudaFreeHost(p); p = nullptr; } n = 0; }
    ~PinnedBuf() { release(); }
    T* get() const { return p; }
};

// Pinned host memory mapped into the device address space: kernels write small results (and a sequence flag) straight
// into it and the host polls, instead of a cudaMemcpyAsync + cudaStreamSynchronize pair per result.
template <class T>
struct MappedBuf {
    T* h = nullptr;    // host address
    T* d = nullptr;    // the same memory as the device sees it
    size_t n = 0;
    MappedBuf() {}
    MappedBuf(const MappedBuf&) = delete;
    MappedBuf& operator=(const MappedBuf&) = delete;
    void alloc(size_t count) {
        release(); n = count;
        if (!count) return;
        SB_CUDA(cudaHostAlloc((void**)&h, count * sizeof(T), cudaHostAllocMapped | cudaHostAllocPortable));
        SB_CUDA(cudaHostGetDevicePointer((void**)&d, (void*)h, 0));
        memset((void*)h, 0, count * sizeof(T));
    }
    void release() { if (h) { cudaFreeHost((void*)h); h = nullptr; d = nullptr; } n = 0; }
    ~MappedBuf() { release(); }
};

static inline int grid_for(size_t work_items, int block, int ctas_per_sm) {
    size_t need = (work_items + block - 1) / block;
    size_t cap = (size_t)SB_SMS * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- 128-bit vectorized global access for field elements (all tables are 16-byte aligned SoA of
// whole elements: one Fr = two LDG.128, one Fq = three)
template <class T>
SB_D T ldg_elem(const T* p) {
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(T) % 16 == 0, "element must be a multiple of 16 bytes");
    T out;
    const uint4* src = reinterpret_cast<const uint4*>(p);
    uint4* dst = reinterpret_cast<uint4*>(&out);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) dst[i] = __ldg(src + i);
    return out;
#else
    return *p;
#endif
}
template <class T>
SB_D void st_elem(T* p, const T& v) {
#if defined(__CUDA_ARCH__)
    uint4* dst = reinterpret_cast<uint4*>(p);
    const uint4* src = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) dst[i] = src[i];
#else
    *p = v;
#endif
}
