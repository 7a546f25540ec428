// comm_shm.cu -- the sharded prover's allgather over POSIX shared memory, for ranks on ONE node.
//
// What the prover exchanges is tiny (96 bytes per sumcheck round, a few hundred per MSM) and is already on the
// host when it is exchanged -- the round result has to reach the host-side transcript anyway -- so the
// collective is latency-bound, not bandwidth-bound: a shared-memory mailbox with sequence counters costs about a
// microsecond, against tens of microseconds for a device-side NCCL collective plus its two host copies.
// (r1cs-spartan_b200/dist.py keeps a torch.distributed / NCCL backed hook as the alternative; SB_COMM=nccl.)
#include "../../include/spartan_b200.h"
#include <atomic>
#include <chrono>
#include <cstring>
#include <fcntl.h>
#include <new>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {
constexpr size_t SHM_SLOT = 16384;     // bytes per rank per buffer
constexpr int SHM_MAX_WORLD = 64;
struct ShmRegion {
    std::atomic<uint64_t> seq[SHM_MAX_WORLD];
    unsigned char pad[64];
    unsigned char slots[2][SHM_MAX_WORLD][SHM_SLOT];
};
struct ShmComm {
    ShmRegion* reg = nullptr;
    int rank = 0, world = 1;
    uint64_t calls = 0;
    std::string name;
    bool owner = false;
};
inline void cpu_relax() {
#if defined(__x86_64__)
    _mm_pause();
#endif
}
int shm_allgather(void* user, const void* send, void* recv, size_t bytes) {
    ShmComm* c = static_cast<ShmComm*>(user);
    if (bytes > SHM_SLOT) return 2;
    const uint64_t k = ++c->calls;
    const int buf = (int)(k & 1);
    memcpy(c->reg->slots[buf][c->rank], send, bytes);
    c->reg->seq[c->rank].store(k, std::memory_order_release);
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < c->world; r++) {
        unsigned spins = 0;
        while (c->reg->seq[r].load(std::memory_order_acquire) < k) {
            cpu_relax();
            if ((++spins & 0xfffff) == 0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0) return 3;   // a peer died
        }
        memcpy(static_cast<unsigned char*>(recv) + (size_t)r * bytes, c->reg->slots[buf][r], bytes);
    }
    // double buffering is enough: a rank can only reach call k+2 (same buffer) after every rank has published call
    // k+1, which each of them does after it has finished reading call k
    return 0;
}
}  // namespace

extern "C" {
// create != 0 on exactly one rank (which must run first, e.g. before a barrier); the others attach.
sb_status sb_comm_shm_open(const char* name, int rank, int world, int create, sb_comm* out) {
    if (!name || !out || world < 1 || world > SHM_MAX_WORLD || rank < 0 || rank >= world) return SB_EINVAL;
    int fd = shm_open(name, create ? (O_CREAT | O_RDWR | O_TRUNC) : O_RDWR, 0600);
    if (fd < 0) return SB_ECOMM;
    if (create && ftruncate(fd, sizeof(ShmRegion)) != 0) { close(fd); return SB_ECOMM; }
    void* p = mmap(nullptr, sizeof(ShmRegion), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return SB_ECOMM;
    ShmComm* c = new (std::nothrow) ShmComm;
    if (!c) { munmap(p, sizeof(ShmRegion)); return SB_ENOMEM; }
    c->reg = static_cast<ShmRegion*>(p); c->rank = rank; c->world = world; c->name = name; c->owner = create != 0;
    if (create) for (int r = 0; r < SHM_MAX_WORLD; r++) c->reg->seq[r].store(0, std::memory_order_relaxed);
    out->rank = rank; out->world = world; out->allgather = shm_allgather; out->user = c;
    return SB_OK;
}
void sb_comm_shm_close(sb_comm* comm) {
    if (!comm || !comm->user) return;
    ShmComm* c = static_cast<ShmComm*>(comm->user);
    munmap(c->reg, sizeof(ShmRegion));
    if (c->owner) shm_unlink(c->name.c_str());
    delete c;
    comm->user = nullptr;
}
}
