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
#include <sched.h>
#include <sys/stat.h>
#include <time.h>
#include <algorithm>
#include <unistd.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {
constexpr size_t SHM_SLOT = 16384;     // payload bytes per rank per buffer
constexpr int SHM_MAX_WORLD = 64;
// Sequence numbers are (epoch << 32) | call: sb_comm::barrier (called at the start of every proof / commitment / opening
// of a sharded context) opens a new epoch, so a rank that dropped out of an earlier operation (a rank-local error
// mid-proof) is re-aligned instead of silently pairing call k of one rank with call k' of another.  Every slot carries
// the sequence number and byte count it was written for and the reader requires both to match exactly: a peer that
// is ahead or behind, or disagrees on the payload size, gives SB_ECOMM -- never stale data.
struct ShmSlot {
    uint64_t seq;
    uint64_t bytes;
    unsigned char data[SHM_SLOT];
};
struct ShmRegion {
    std::atomic<uint64_t> seq[SHM_MAX_WORLD];
    unsigned char pad[64];
    ShmSlot slots[2][SHM_MAX_WORLD];
};
struct ShmComm {
    ShmRegion* reg = nullptr;
    int rank = 0, world = 1;
    uint64_t epoch = 0, calls = 0;
    std::string name;
    bool owner = false, in_process = false;
    double timeout_s = 120.0;
};
inline void cpu_relax() {
#if defined(__x86_64__)
    _mm_pause();
#endif
}
// wait until rank r has published at least `want`; false on timeout (a peer died)
bool wait_for(ShmComm* c, int r, uint64_t want) {
    const auto t0 = std::chrono::steady_clock::now();
    for (uint64_t spins = 0; c->reg->seq[r].load(std::memory_order_acquire) < want; spins++) {
        if (spins < 4096) cpu_relax();
        else if (spins < 65536) sched_yield();                           // back off: do not burn a core the peer may need
        else { struct timespec ts = {0, 50000}; nanosleep(&ts, nullptr); }
        if ((spins & 0x3fff) == 0x3fff && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > c->timeout_s) return false;
    }
    return true;
}
int shm_allgather(void* user, const void* send, void* recv, size_t bytes) {
    ShmComm* c = static_cast<ShmComm*>(user);
    if (bytes > SHM_SLOT) return 2;
    const uint64_t k = (c->epoch << 32) | ++c->calls;
    const int buf = (int)(c->calls & 1);
    ShmSlot& mine = c->reg->slots[buf][c->rank];
    mine.seq = k; mine.bytes = bytes;
    memcpy(mine.data, send, bytes);
    c->reg->seq[c->rank].store(k, std::memory_order_release);
    for (int r = 0; r < c->world; r++) {
        if (!wait_for(c, r, k)) return 3;                                 // a peer died
        const ShmSlot& s = c->reg->slots[buf][r];
        if (s.seq != k || s.bytes != bytes) return 4;                     // the peer is in another call / epoch, or disagrees on the size
        memcpy(static_cast<unsigned char*>(recv) + (size_t)r * bytes, s.data, bytes);
        if (c->reg->seq[r].load(std::memory_order_acquire) > k + 1) return 4;   // the slot may have been rewritten while it was read
    }
    // double buffering is enough: a rank can only reach call k+2 (same buffer) after every rank has published call
    // k+1, which each of them does after it has finished reading call k
    return 0;
}
// New epoch (sb_comm::barrier).  Every rank calls it once at the start of every library call on a sharded context, so
// the epoch numbers agree across ranks as long as the ranks make the same sequence of calls (SPMD), whether or not
// an earlier call failed on some of them.  A peer still waiting inside an allgather of the previous epoch is released
// by the jump of this rank's sequence number and fails that call with a mismatch instead of hanging.
int shm_barrier(void* user) {
    ShmComm* c = static_cast<ShmComm*>(user);
    c->epoch++; c->calls = 0;
    const uint64_t k = c->epoch << 32;
    c->reg->seq[c->rank].store(k, std::memory_order_release);
    for (int r = 0; r < c->world; r++)
        if (!wait_for(c, r, k)) return 3;
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
    if (const char* e = getenv("SB_COMM_TIMEOUT_S")) c->timeout_s = atof(e) > 0 ? atof(e) : c->timeout_s;
    if (create) for (int r = 0; r < SHM_MAX_WORLD; r++) c->reg->seq[r].store(0, std::memory_order_relaxed);
    out->rank = rank; out->world = world; out->allgather = shm_allgather; out->barrier = shm_barrier; out->user = c;
    return SB_OK;
}
// The same mailbox in plain process memory, for `world` contexts driven by threads of ONE process
// (sb_ctx_create_multi uses it); fills out[0 .. world).  Close every entry with sb_comm_shm_close.
sb_status sb_comm_local_open(int world, sb_comm* out) {
    if (!out || world < 1 || world > SHM_MAX_WORLD) return SB_EINVAL;
    ShmRegion* reg = new (std::nothrow) ShmRegion;
    if (!reg) return SB_ENOMEM;
    for (int r = 0; r < SHM_MAX_WORLD; r++) reg->seq[r].store(0, std::memory_order_relaxed);
    for (int r = 0; r < world; r++) {
        ShmComm* c = new (std::nothrow) ShmComm;
        if (!c) return SB_ENOMEM;
        c->reg = reg; c->rank = r; c->world = world; c->in_process = true; c->owner = r == 0;
        if (const char* e = getenv("SB_COMM_TIMEOUT_S")) c->timeout_s = atof(e) > 0 ? atof(e) : c->timeout_s;
        out[r].rank = r; out[r].world = world; out[r].allgather = shm_allgather; out[r].barrier = shm_barrier; out[r].user = c;
    }
    return SB_OK;
}
// A rank whose library call failed locally (bad argument on one rank only, out of memory, a CUDA error) tells its peers:
// its sequence number jumps to the next epoch, so a peer waiting for it inside an allgather stops waiting and fails that
// call with a mismatch at once instead of after the time-out.  The next library call re-aligns everybody (shm_barrier).
void sb_comm_shm_abort(sb_comm* comm) {
    if (!comm || !comm->user) return;
    ShmComm* c = static_cast<ShmComm*>(comm->user);
    c->reg->seq[c->rank].store((c->epoch + 1) << 32, std::memory_order_release);
}
void sb_comm_shm_close(sb_comm* comm) {
    if (!comm || !comm->user) return;
    ShmComm* c = static_cast<ShmComm*>(comm->user);
    if (c->in_process) { if (c->owner) delete c->reg; }          // close rank 0's entry last
    else {
        munmap(c->reg, sizeof(ShmRegion));
        if (c->owner) shm_unlink(c->name.c_str());
    }
    delete c;
    comm->user = nullptr;
}
}
