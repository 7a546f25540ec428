// kernels_fr.cu -- scalar-field kernels: eq pyramid, segmented sparse products, fused
// fold+evaluate sumcheck rounds, opening fold.  sm_100a; integer pipe + HBM bound, no tensor cores
// (no step of this path is a dense contraction).
#include "kernels_fr.cuh"

#include <map>
#include <string>
#include <vector>

std::atomic<unsigned long long> g_sb_launches{0};
std::atomic<unsigned long long> g_sb_h2d_bytes{0}, g_sb_d2h_bytes{0};
bool g_sb_prof_on = false;
int g_sb_prof_tag = -1;            // free-form tag attached to the records (the MSM code sets log2 of the job size)

// ------------------------------------------------------------------ per-kernel CUDA-event profiler
namespace {
struct ProfRec { const char* name; cudaEvent_t e0, e1; int tag; };
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
void sb_prof_begin(const char* name, cudaStream_t stream) {
    ProfRec r{name, prof_event(), prof_event(), g_sb_prof_tag};
    cudaEventRecord(r.e0, stream);
    g_prof_recs.push_back(r);
}
void sb_prof_end(cudaStream_t stream) { cudaEventRecord(g_prof_recs.back().e1, stream); }
// JSON array [[name, start_ms, end_ms, tag], ...] relative to the first recorded launch; clears the records
std::string sb_prof_timeline_collect() {
    std::string out = "[";
    if (!g_prof_recs.empty()) {
        for (auto& r : g_prof_recs) cudaEventSynchronize(r.e1);
        cudaEvent_t base = g_prof_recs[0].e0;
        bool first = true;
        for (auto& r : g_prof_recs) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, base, r.e0); cudaEventElapsedTime(&t1, base, r.e1);
            char buf[256];
            snprintf(buf, sizeof buf, "%s[\"%s\", %.4f, %.4f, %d]", first ? "" : ", ", r.name, t0, t1, r.tag);
            out += buf; first = false;
        }
        for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
        g_prof_recs.clear();
    }
    return out + "]";
}
// JSON object {"kernel": {"launches": n, "ms": total}, ...}; clears the records
std::string sb_prof_collect() {
    std::map<std::string, std::pair<int, double>> acc;
    for (auto& r : g_prof_recs) {
        cudaEventSynchronize(r.e1);
        float ms = 0; cudaEventElapsedTime(&ms, r.e0, r.e1);
        auto& a = acc[r.name]; a.first++; a.second += ms;
        g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : acc) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %d, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
        out += buf; first = false;
    }
    return out + "}";
}

// ------------------------------------------------------------------ helpers
template <class T>
SB_D T ldcg_elem(const T* p) {      // L2-coherent load (data produced by other CTAs of this launch)
    T out;
    const uint4* src = reinterpret_cast<const uint4*>(p);
    uint4* dst = reinterpret_cast<uint4*>(&out);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) dst[i] = __ldcg(src + i);
    return out;
}

SB_D Fr warp_reduce_fr(Fr v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < Fr::N; i++) o.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], off);
        v = Fr::add(v, o);
    }
    return v;
}

// Block tree reduction of K accumulators, then grid-level finish by the last CTA to arrive.
// `out` may be device memory or mapped pinned host memory; when `flag` is given (mapped pinned host memory) the last
// CTA publishes `seq` there after the results, system-wide fenced, and the host polls it instead of synchronising the
// stream (a sumcheck round is a few microseconds of kernel: the cudaMemcpyAsync + cudaStreamSynchronize pair that used
// to fetch its 96 bytes cost more than the round itself).
template <int K>
SB_D void reduce_finish(Fr (&acc)[K], Fr* out, Fr* block_partials, unsigned int* ticket, volatile uint32_t* flag, uint32_t seq) {
    __shared__ Fr sh[K][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        Fr v = warp_reduce_fr(acc[k]);
        if (lane == 0) sh[k][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            Fr v = lane < nwarps ? sh[k][lane] : Fr::zero();
            v = warp_reduce_fr(v);
            if (lane == 0) st_elem(&block_partials[(size_t)blockIdx.x * K + k], v);
        }
        if (lane == 0) {
            __threadfence();
            unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Fr s[K];
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = Fr::zero();
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; k++) s[k] = Fr::add(s[k], ldcg_elem(&block_partials[(size_t)i * K + k]));
    }
    __syncthreads();   // sh reuse
#pragma unroll
    for (int k = 0; k < K; k++) {
        Fr v = warp_reduce_fr(s[k]);
        if (lane == 0) sh[k][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            Fr v = lane < nwarps ? sh[k][lane] : Fr::zero();
            v = warp_reduce_fr(v);
            if (lane == 0) st_elem(&out[k], v);
        }
        if (lane == 0) {
            *ticket = 0;                // self-reset for the next launch on this stream
            if (flag) { __threadfence_system(); *flag = seq; }
        }
    }
}

// ------------------------------------------------------------------ eq pyramid
constexpr int EQ_SMALL_LOG = 10;

// levels of size 1 .. 2^kmax inside one CTA
__global__ void __launch_bounds__(1024) k_eq_small(Fr* pyr, const Fr* tau, uint32_t nv, uint32_t kmax) {
    if (threadIdx.x == 0) st_elem(&pyr[1], Fr::one());
    __syncthreads();
    for (uint32_t k = 0; k < kmax; k++) {
        size_t s = (size_t)1 << k;
        Fr t = ldg_elem(&tau[nv - 1 - k]);
        for (size_t b = threadIdx.x; b < s; b += blockDim.x) {
            Fr v = pyr[s + b];
            Fr hi = Fr::mul(v, t);
            st_elem(&pyr[2 * s + 2 * b], Fr::sub(v, hi));
            st_elem(&pyr[2 * s + 2 * b + 1], hi);
        }
        __syncthreads();
    }
}
// out[2b] = in[b] (1 - t), out[2b+1] = in[b] t     (one multiplication per input)
__global__ void __launch_bounds__(256) k_eq_double(const Fr* __restrict__ in, Fr* __restrict__ out, size_t s, const Fr* __restrict__ tau) {
    Fr t = ldg_elem(tau);
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < s; b += (size_t)gridDim.x * blockDim.x) {
        Fr v = ldg_elem(&in[b]);
        Fr hi = Fr::mul(v, t);
        st_elem(&out[2 * b], Fr::sub(v, hi));
        st_elem(&out[2 * b + 1], hi);
    }
}
__global__ void __launch_bounds__(256) k_eq_double_scaled3(const Fr* __restrict__ in, Fr* __restrict__ out3, size_t s,
                                                           const Fr* __restrict__ tau, const Fr* __restrict__ rabc) {
    Fr t = ldg_elem(tau);
    Fr omt = Fr::sub(Fr::one(), t);
    Fr c0[3], c1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { Fr r = ldg_elem(&rabc[k]); c0[k] = Fr::mul(r, omt); c1[k] = Fr::mul(r, t); }
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < s; b += (size_t)gridDim.x * blockDim.x) {
        Fr v = ldg_elem(&in[b]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            st_elem(&out3[(size_t)k * 2 * s + 2 * b], Fr::mul(v, c0[k]));
            st_elem(&out3[(size_t)k * 2 * s + 2 * b + 1], Fr::mul(v, c1[k]));
        }
    }
}

void launch_eq_pyramid(Fr* pyr, const Fr* tau_dev, uint32_t nv, cudaStream_t stream) {
    if (nv == 0) return;
    uint32_t top = nv - 1;                              // largest level: size 2^(nv-1)
    uint32_t kmax = top < (uint32_t)EQ_SMALL_LOG ? top : (uint32_t)EQ_SMALL_LOG;
    SB_LAUNCH(k_eq_small, 1, 1024, 0, stream, pyr, tau_dev, nv, kmax);
    for (uint32_t k = kmax; k < top; k++) {
        size_t s = (size_t)1 << k;
        SB_LAUNCH(k_eq_double, grid_for(s, 256, 8), 256, 0, stream, pyr + s, pyr + 2 * s, s, tau_dev + (nv - 1 - k));
    }
}
void launch_eq_full(Fr* out, const Fr* pyr, const Fr* tau_dev, uint32_t nv, cudaStream_t stream) {
    size_t s = (size_t)1 << (nv - 1);
    SB_LAUNCH(k_eq_double, grid_for(s, 256, 8), 256, 0, stream, pyr + s, out, s, tau_dev);
}
void launch_eq_full_scaled3(Fr* out3, const Fr* pyr, const Fr* tau_dev, const Fr* rabc_dev, uint32_t nv, cudaStream_t stream) {
    size_t s = (size_t)1 << (nv - 1);
    SB_LAUNCH(k_eq_double_scaled3, grid_for(s, 256, 8), 256, 0, stream, pyr + s, out3, s, tau_dev, rabc_dev);
}

// ------------------------------------------------------------------ segmented sparse product
__global__ void __launch_bounds__(128) k_segsum(Fr* __restrict__ out, Fr* __restrict__ partials, const SegItem* __restrict__ items,
                                                uint32_t n_items, const Fr* __restrict__ val, const uint32_t* __restrict__ idx,
                                                const Fr* __restrict__ x) {
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += gridDim.x * blockDim.x) {
        SegItem item = items[it];
        Fr acc = Fr::zero();
        for (uint32_t e = item.start; e < item.start + item.len; e++) {
            uint32_t g = __ldg(&idx[e]);
            Fr xv = ldg_elem(&x[g & ~SEG_UNIT_FLAG]);
            if (g & SEG_UNIT_FLAG) acc = Fr::add(acc, xv);
            else acc = Fr::add(acc, Fr::mul(ldg_elem(&val[e]), xv));
        }
        if (item.out & 0x80000000u) st_elem(&partials[item.out & 0x7fffffffu], acc);
        else st_elem(&out[item.out], acc);
    }
}
// one CTA per long segment: sum its partial slots
__global__ void __launch_bounds__(256) k_seg_fixup(Fr* __restrict__ out, const Fr* __restrict__ partials, const SegFixup* __restrict__ fix) {
    __shared__ Fr sh[32];
    SegFixup f = fix[blockIdx.x];
    Fr acc = Fr::zero();
    for (uint32_t i = threadIdx.x; i < f.pcount; i += blockDim.x) acc = Fr::add(acc, ldg_elem(&partials[f.pstart + i]));
    acc = warp_reduce_fr(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        Fr v = lane < nwarps ? sh[lane] : Fr::zero();
        v = warp_reduce_fr(v);
        if (lane == 0) st_elem(&out[f.seg], v);
    }
}
void launch_segsum(Fr* out, Fr* partials, const SegItem* items, uint32_t n_items, const SegFixup* fix, uint32_t n_fix,
                   const Fr* val, const uint32_t* idx, const Fr* x, cudaStream_t stream) {
    if (n_items) SB_LAUNCH(k_segsum, grid_for(n_items, 128, 16), 128, 0, stream, out, partials, items, n_items, val, idx, x);
    if (n_fix) SB_LAUNCH(k_seg_fixup, (int)n_fix, 256, 0, stream, out, partials, fix);
}

// ------------------------------------------------------------------ sumcheck rounds
// KIND 1: S(t) = sum_b E[b] (A(t,b) B(t,b) - C(t,b))   (3 tables + eq suffix weights)
// KIND 2: S(t) = sum_b M(t,b) Z(t,b)                   (2 tables; C, E unused)
// FOLD: tables are first folded with r (fix the lowest variable), the folded halves are written out
// for the next round, and the evaluation runs on the folded values still in registers.
template <int KIND, bool FOLD>
__global__ void __launch_bounds__(256, 2)
k_sc_round(const Fr* __restrict__ A, const Fr* __restrict__ B, const Fr* __restrict__ C, Fr* __restrict__ Ao, Fr* __restrict__ Bo,
           Fr* __restrict__ Co, const Fr* __restrict__ E, const Fr r /* the challenge, a kernel argument: no upload */, size_t h /* pairs after the fold */,
           Fr* out3, Fr* block_partials, unsigned int* ticket, volatile uint32_t* flag, uint32_t seq) {
    Fr acc[3];
    acc[0] = Fr::zero(); acc[1] = Fr::zero(); acc[2] = Fr::zero();
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < h; b += (size_t)gridDim.x * blockDim.x) {
        Fr a0, a1, b0, b1, c0, c1;
        if (FOLD) {
            Fr x0 = ldg_elem(&A[4 * b]), x1 = ldg_elem(&A[4 * b + 1]), x2 = ldg_elem(&A[4 * b + 2]), x3 = ldg_elem(&A[4 * b + 3]);
            a0 = Fr::add(x0, Fr::mul(r, Fr::sub(x1, x0)));
            a1 = Fr::add(x2, Fr::mul(r, Fr::sub(x3, x2)));
            st_elem(&Ao[2 * b], a0); st_elem(&Ao[2 * b + 1], a1);
            x0 = ldg_elem(&B[4 * b]); x1 = ldg_elem(&B[4 * b + 1]); x2 = ldg_elem(&B[4 * b + 2]); x3 = ldg_elem(&B[4 * b + 3]);
            b0 = Fr::add(x0, Fr::mul(r, Fr::sub(x1, x0)));
            b1 = Fr::add(x2, Fr::mul(r, Fr::sub(x3, x2)));
            st_elem(&Bo[2 * b], b0); st_elem(&Bo[2 * b + 1], b1);
            if (KIND == 1) {
                x0 = ldg_elem(&C[4 * b]); x1 = ldg_elem(&C[4 * b + 1]); x2 = ldg_elem(&C[4 * b + 2]); x3 = ldg_elem(&C[4 * b + 3]);
                c0 = Fr::add(x0, Fr::mul(r, Fr::sub(x1, x0)));
                c1 = Fr::add(x2, Fr::mul(r, Fr::sub(x3, x2)));
                st_elem(&Co[2 * b], c0); st_elem(&Co[2 * b + 1], c1);
            }
        } else {
            a0 = ldg_elem(&A[2 * b]); a1 = ldg_elem(&A[2 * b + 1]);
            b0 = ldg_elem(&B[2 * b]); b1 = ldg_elem(&B[2 * b + 1]);
            if (KIND == 1) { c0 = ldg_elem(&C[2 * b]); c1 = ldg_elem(&C[2 * b + 1]); }
        }
        // t = 2: T(2) = 2 T1 - T0
        Fr a2 = Fr::sub(Fr::dbl(a1), a0), b2 = Fr::sub(Fr::dbl(b1), b0);
        if (KIND == 1) {
            Fr c2 = Fr::sub(Fr::dbl(c1), c0);
            Fr e = ldg_elem(&E[b]);
            acc[0] = Fr::add(acc[0], Fr::mul(e, Fr::sub(Fr::mul(a0, b0), c0)));
            acc[1] = Fr::add(acc[1], Fr::mul(e, Fr::sub(Fr::mul(a1, b1), c1)));
            acc[2] = Fr::add(acc[2], Fr::mul(e, Fr::sub(Fr::mul(a2, b2), c2)));
        } else {
            acc[0] = Fr::add(acc[0], Fr::mul(a0, b0));
            acc[1] = Fr::add(acc[1], Fr::mul(a1, b1));
            acc[2] = Fr::add(acc[2], Fr::mul(a2, b2));
        }
    }
    reduce_finish<3>(acc, out3, block_partials, ticket, flag, seq);
}

// CTA size of a round over h work items.  The tables are powers of two and the machine is not: with 148 SMs x 2 CTAs x 256
// threads, 2^18 items are 3.46 per thread -- the last of four passes runs 46 % full and the kernel cannot exceed 86 % of
// its own inner-loop rate (round 1 measured 0.58 / 0.50 of the Fr-product ceiling at 2^20 against 0.73-0.77 at 2^22).
// 224-thread CTAs make 2^k items 0.988 x an integer number of passes for every k >= 16; the choice below takes the
// CTA size (a multiple of a warp) whose last pass is fullest.  SB_SC_BLOCK forces a size.  OFF by default: see below.
static int sc_block_for(size_t h, int cps) {
    static const int forced = getenv("SB_SC_BLOCK") ? atoi(getenv("SB_SC_BLOCK")) : 0;
    if (forced >= 32 && forced <= 256 && forced % 32 == 0) return forced;
    if (forced != 1) return 256;     // measured (round 2, sweep 7): the fuller last pass does not pay for 14 instead of 16 warps per SM (0.552 against 0.589 of the ceiling at 2^20); SB_SC_BLOCK=1 turns the choice below on
    int best = 256;
    double best_eff = 0;
    for (int blk = 256; blk >= 160; blk -= 32) {
        const double x = (double)h / ((double)SB_SMS * cps * blk);
        if (x <= 1.0) return best_eff > 0 ? best : 256;
        const double eff = x / (double)(size_t)(x + 0.999999);
        if (eff > best_eff + 0.02) { best_eff = eff; best = blk; }
    }
    return best;
}
template <int KIND>
static void launch_sc_round(const Fr* A, const Fr* B, const Fr* C, Fr* Ao, Fr* Bo, Fr* Co, const Fr* E, const Fr* r_host,
                            size_t m_in, const RoundOut& o, const RoundWs& ws, cudaStream_t stream) {
    const bool fold = r_host != nullptr;
    const Fr r = fold ? *r_host : Fr::zero();
    size_t h = fold ? m_in / 4 : m_in / 2;
    const int cps = ws.max_grid / SB_SMS > 1 ? ws.max_grid / SB_SMS : 1;                  // 2 resident CTAs per SM by default
    const int block = sc_block_for(h, cps);
    int grid = grid_for(h, block, cps);
    if (grid > ws.max_grid) grid = ws.max_grid;
    if (fold) SB_LAUNCH_NAMED(KIND == 1 ? "k_sc_round<sc1,fold>" : "k_sc_round<sc2,fold>", (k_sc_round<KIND, true>), grid, block, 0, stream, A, B, C, Ao, Bo, Co, E, r, h, o.out, ws.block_partials, ws.ticket, o.flag, o.seq);
    else SB_LAUNCH_NAMED(KIND == 1 ? "k_sc_round<sc1,first>" : "k_sc_round<sc2,first>", (k_sc_round<KIND, false>), grid, block, 0, stream, A, B, C, Ao, Bo, Co, E, r, h, o.out, ws.block_partials, ws.ticket, o.flag, o.seq);
}
void launch_sc1_round(const Fr* A, const Fr* B, const Fr* C, Fr* Ao, Fr* Bo, Fr* Co, const Fr* E, const Fr* r_host,
                      size_t m_in, const RoundOut& o, const RoundWs& ws, cudaStream_t stream) {
    launch_sc_round<1>(A, B, C, Ao, Bo, Co, E, r_host, m_in, o, ws, stream);
}
void launch_sc2_round(const Fr* M, const Fr* Z, Fr* Mo, Fr* Zo, const Fr* r_host, size_t m_in, const RoundOut& o,
                      const RoundWs& ws, cudaStream_t stream) {
    launch_sc_round<2>(M, Z, nullptr, Mo, Zo, nullptr, nullptr, r_host, m_in, o, ws, stream);
}

__global__ void k_final_fold3(const Fr* A, const Fr* B, const Fr* C, int ntab, const Fr r, Fr* out, volatile uint32_t* flag, uint32_t seq) {
    if (threadIdx.x != 0) return;           // three multiplications: one thread, so that the flag follows its own stores
    for (int k = 0; k < ntab; k++) {
        const Fr* T = k == 0 ? A : k == 1 ? B : C;
        Fr x0 = T[0], x1 = T[1];
        st_elem(&out[k], Fr::add(x0, Fr::mul(r, Fr::sub(x1, x0))));
    }
    if (flag) { __threadfence_system(); *flag = seq; }
}
void launch_final_fold3(const Fr* A, const Fr* B, const Fr* C, int ntab, const Fr* r_host, const RoundOut& o, cudaStream_t stream) {
    SB_LAUNCH(k_final_fold3, 1, 32, 0, stream, A, B, C, ntab, *r_host, o.out, o.flag, o.seq);
}

// ------------------------------------------------------------------ opening fold
__global__ void __launch_bounds__(256) k_open_fold(const Fr* __restrict__ in, Fr* __restrict__ r_out, Fr* __restrict__ q_out,
                                                   const Fr* __restrict__ p_ptr, size_t half) {
    Fr p = ldg_elem(p_ptr);
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < half; b += (size_t)gridDim.x * blockDim.x) {
        Fr x0 = ldg_elem(&in[2 * b]), x1 = ldg_elem(&in[2 * b + 1]);
        Fr q = Fr::sub(x1, x0);
        st_elem(&q_out[b], q);
        st_elem(&r_out[b], Fr::add(x0, Fr::mul(p, q)));
    }
}
// q[b] = in[2b+1] - in[2b]: the first quotient of an opening (open.rs:42 at i = 0), which does not depend on the point
__global__ void __launch_bounds__(256) k_pair_diff(const Fr* __restrict__ in, Fr* __restrict__ q_out, size_t half) {
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < half; b += (size_t)gridDim.x * blockDim.x)
        st_elem(&q_out[b], Fr::sub(ldg_elem(&in[2 * b + 1]), ldg_elem(&in[2 * b])));
}
void launch_pair_diff(const Fr* in, Fr* q_out, size_t half, cudaStream_t stream) {
    SB_LAUNCH(k_pair_diff, grid_for(half, 256, 8), 256, 0, stream, in, q_out, half);
}
void launch_open_fold(const Fr* in, Fr* r_out, Fr* q_out, const Fr* p_dev, size_t half, cudaStream_t stream) {
    SB_LAUNCH(k_open_fold, grid_for(half, 256, 8), 256, 0, stream, in, r_out, q_out, p_dev, half);
}

// ------------------------------------------------------------------ self-test / microbenchmarks
template <class F>
__global__ void k_binop(int op, const F* a, const F* b, F* out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        F x = a[i], y = b[i], o;
        if (op == 0) o = F::add(x, y);
        else if (op == 1) o = F::sub(x, y);
        else if (op == 2) o = F::mul(x, y);
        else o = F::mul_portable(x, y);
        out[i] = o;
    }
}
void launch_fr_binop(int op, const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t stream) {
    SB_LAUNCH(k_binop<Fr>, grid_for(n, 256, 4), 256, 0, stream, op, a, b, out, n);
}
void launch_fq_binop(int op, const Fq* a, const Fq* b, Fq* out, size_t n, cudaStream_t stream) {
    SB_LAUNCH(k_binop<Fq>, grid_for(n, 256, 4), 256, 0, stream, op, a, b, out, n);
}
template <class F>
__global__ void __launch_bounds__(256) k_mul_bench(F* inout, size_t n, int iters) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = inout[i], y = x, z = F::add(x, F::one());
    for (int it = 0; it < iters; it++) { x = F::mul(x, y); z = F::mul(z, y); }   // two independent chains
    inout[i] = F::add(x, z);
}
void launch_fr_mul_bench(Fr* inout, size_t n_threads, int iters, cudaStream_t stream) {
    SB_LAUNCH(k_mul_bench<Fr>, (int)((n_threads + 255) / 256), 256, 0, stream, inout, n_threads, iters);
}
void launch_fq_mul_bench(Fq* inout, size_t n_threads, int iters, cudaStream_t stream) {
    SB_LAUNCH(k_mul_bench<Fq>, (int)((n_threads + 255) / 256), 256, 0, stream, inout, n_threads, iters);
}
