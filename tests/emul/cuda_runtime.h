// tests/emul/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: a minimal CUDA-on-CPU shim.
//
// The authoring container has no GPU, and GPU time is scarce; this header lets tests/emul/build.sh compile the
// product's own .cu sources (csrc/*.cu, unchanged) with g++ into tests/emul/libsb_emul_TESTONLY.so, in which every
// kernel launch runs its CUDA threads as host threads, block after block: __syncthreads() is a barrier over the
// block's live threads, __shared__ variables are statics (blocks run one at a time), warp shuffles exchange
// through a per-warp buffer, atomics are host atomics.  That exercises the INDEX LOGIC of the kernels (plans, sorts,
// chunking, reductions, sharding) and the whole host driver against the oracle on a CPU, in seconds, before a
// change is sent to a GPU box.  It does NOT exercise the PTX field arithmetic (the portable C++ path of field.cuh
// runs instead); the -m gpu tests cover that on the device.
//
// The product never builds, links or loads this: the library refuses to create a context unless SB_EMUL_TESTS=1 is set.
#pragma once
#define SB_EMUL 1
#include <atomic>
#include <barrier>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

namespace emu {
struct BlockState {
    std::barrier<>* block_bar;
    std::vector<std::unique_ptr<std::barrier<>>>* warp_bar;
    uint32_t (*xch)[32];
    unsigned char* dyn_smem;
    unsigned nthreads;
};
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern BlockState* g_block;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn);
}  // namespace emu
#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

static inline void __syncthreads() { emu::g_block->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (*emu::g_block->warp_bar)[emu::t_threadIdx.x >> 5]->arrive_and_wait(); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, int off) {
    const unsigned lane = emu::t_threadIdx.x & 31, warp = emu::t_threadIdx.x >> 5;
    const unsigned live = std::min(32u, emu::g_block->nthreads - 32 * warp);
    uint32_t* x = emu::g_block->xch[warp];
    x[lane] = v;
    (*emu::g_block->warp_bar)[warp]->arrive_and_wait();
    const uint32_t r = (lane + off < live) ? x[lane + off] : v;
    (*emu::g_block->warp_bar)[warp]->arrive_and_wait();
    return r;
}
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) {
    const unsigned lane = emu::t_threadIdx.x & 31, warp = emu::t_threadIdx.x >> 5;
    const unsigned live = std::min(32u, emu::g_block->nthreads - 32 * warp);
    uint32_t* x = emu::g_block->xch[warp];
    x[lane] = v;
    (*emu::g_block->warp_bar)[warp]->arrive_and_wait();
    const uint32_t r = ((unsigned)src < live) ? x[src] : v;
    (*emu::g_block->warp_bar)[warp]->arrive_and_wait();
    return r;
}
static inline uint32_t __shfl_xor_sync(unsigned m, uint32_t v, int mask) { return __shfl_sync(m, v, (int)((emu::t_threadIdx.x & 31) ^ (unsigned)mask)); }
static inline uint32_t __shfl_up_sync(unsigned m, uint32_t v, int off) {
    const int lane = (int)(emu::t_threadIdx.x & 31);
    const uint32_t r = __shfl_sync(m, v, lane >= off ? lane - off : lane);
    return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { return a < b ? a : b; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { return a > b ? a : b; }

// ------------------------------------------------------------------ runtime API (everything is synchronous)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
typedef void* cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2, cudaHostAllocPortable = 1 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = getenv("SB_EMUL_TESTS") ? (getenv("SB_EMUL_DEVICES") ? atoi(getenv("SB_EMUL_DEVICES")) : 1) : 0; return *n ? cudaSuccess : 100; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { return cudaMallocAsync(p, n, nullptr); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
static inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* ok, int, int) { *ok = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 4; return cudaSuccess; }
static inline cudaError_t cudaFuncSetAttribute(const void*, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
