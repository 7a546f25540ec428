// tests/emul/emul_runtime.cpp -- the block executor of the CUDA-on-CPU test shim (see cuda_runtime.h there).
#include <cuda_runtime.h>
#include <semaphore>

namespace emu {
thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
BlockState* g_block = nullptr;

namespace {
// persistent workers: worker i runs CUDA thread i of the current block; only the workers a block needs are woken
struct Worker {
    std::binary_semaphore go{0};
    std::thread th;
};
struct Pool {
    std::vector<std::unique_ptr<Worker>> w;
    std::atomic<unsigned> remaining{0};
    std::binary_semaphore done{0};
    const std::function<void()>* fn = nullptr;
    uint3 bidx{}; dim3 bdim, gdim;
    void worker(unsigned i) {
        for (;;) {
            w[i]->go.acquire();
            t_threadIdx = uint3{i, 0, 0}; t_blockIdx = bidx; t_blockDim = bdim; t_gridDim = gdim;
            (*fn)();
            g_block->block_bar->arrive_and_drop();                      // a finished thread no longer takes part in barriers
            (*g_block->warp_bar)[i >> 5]->arrive_and_drop();
            if (remaining.fetch_sub(1, std::memory_order_acq_rel) == 1) done.release();
        }
    }
    void ensure(unsigned n) {
        if (w.capacity() < 2048) w.reserve(2048);                       // workers index w[] concurrently: never reallocate
        while (w.size() < n) {
            unsigned i = (unsigned)w.size();
            w.emplace_back(new Worker);
            w[i]->th = std::thread([this, i] { worker(i); });
            w[i]->th.detach();
        }
    }
    void run_block(unsigned n, const std::function<void()>& f, uint3 b, dim3 bd, dim3 gd) {
        ensure(n);
        fn = &f; bidx = b; bdim = bd; gdim = gd;
        remaining.store(n, std::memory_order_release);
        for (unsigned i = 0; i < n; i++) w[i]->go.release();
        done.acquire();
    }
};
Pool& pool() { static Pool* p = new Pool; return *p; }     // leaked on purpose: no destructor races at exit
std::mutex g_launch_mutex;                                   // one kernel at a time, whichever host thread launches it
}  // namespace

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
    std::lock_guard<std::mutex> lg(g_launch_mutex);
    const unsigned n = block.x;
    if (n == 0 || grid.x == 0) return;
    std::vector<unsigned char> dyn(smem + 64);
    uint32_t (*xch)[32] = new uint32_t[(n + 31) / 32][32];
    for (unsigned b = 0; b < grid.x; b++) {
        std::barrier<> bar((std::ptrdiff_t)n);
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        for (unsigned w = 0; w < (n + 31) / 32; w++) wb.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, n - 32 * w)));
        BlockState st{&bar, &wb, xch, (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63), n};
        g_block = &st;
        if (n == 1) {      // single-thread blocks run inline
            t_threadIdx = uint3{0, 0, 0}; t_blockIdx = uint3{b, 0, 0}; t_blockDim = block; t_gridDim = grid;
            fn();
        } else {
            pool().run_block(n, fn, uint3{b, 0, 0}, block, grid);
        }
        g_block = nullptr;
    }
    delete[] xch;
}
}  // namespace emu
