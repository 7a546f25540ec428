// kernels_fr.cuh -- launchers for the scalar-field (Fr) kernels of the prover hot path.
// Every launcher is asynchronous on `stream`; pointers are device pointers unless noted.
#pragma once
#include "common.cuh"

// One work item of the segmented sparse product (see SegPlan in index.cuh).
struct SegItem {
    uint32_t start;   // first entry
    uint32_t len;     // number of entries (<= SEG_LMAX)
    uint32_t out;     // output slot; bit 31 set => slot in the partials array
};
struct SegFixup {
    uint32_t seg;     // output segment
    uint32_t pstart;  // first partial slot
    uint32_t pcount;  // number of partial slots
};
constexpr uint32_t SEG_LMAX = 32;
constexpr uint32_t SEG_UNIT_FLAG = 0x80000000u;   // entry coefficient is exactly 1: skip the multiply

// eq(tau, .) suffix pyramid (replaces src/data_structures/eq.rs:5-20):
// pyr[s + b] for s = 2^k, b < s holds prod_{i > j} eq_i(tau_i, b_{i-j-1}) with j = nv-1-k;
// i.e. the level of size 2^k covers variables nv-k .. nv-1.  pyr[1] = 1.  pyr has 2^nv entries.
void launch_eq_pyramid(Fr* pyr, const Fr* tau_dev, uint32_t nv, cudaStream_t stream);
// out[2b + b0] = eq_0(tau_0, b0) * pyr_level0[b]  (the full 2^nv table)
void launch_eq_full(Fr* out, const Fr* pyr, const Fr* tau_dev, uint32_t nv, cudaStream_t stream);
// out_k[2b + b0] = r_k * eq_0(tau_0, b0) * pyr_level0[b], k = 0..2, out = 3 tables of 2^nv back to back
void launch_eq_full_scaled3(Fr* out3, const Fr* pyr, const Fr* tau_dev, const Fr* rabc_dev, uint32_t nv, cudaStream_t stream);

// out[seg] = sum_e val[e] * x[idx[e]] over the plan (replaces sum_over_y r1cs_reader.rs:75-85 and,
// on the transposed plan, eval_on_x r1cs_reader.rs:91-117).  `out` must be zeroed by the caller.
void launch_segsum(Fr* out, Fr* partials, const SegItem* items, uint32_t n_items, const SegFixup* fix, uint32_t n_fix,
                   const Fr* val, const uint32_t* idx, const Fr* x, cudaStream_t stream);

// Workspace for the per-round reductions (one per prover).
struct RoundWs {
    Fr* block_partials;      // [max_grid][3]
    unsigned int* ticket;    // zero-initialised, self-resetting
    int max_grid;
};
// Where a round leaves its result: `out` (3 Fr; device memory or mapped pinned host memory) and, optionally, a flag
// word in mapped pinned host memory that receives `seq` once the result is visible to the host (see reduce_finish).
struct RoundOut {
    Fr* out;
    volatile uint32_t* flag;
    uint32_t seq;
};
// Sumcheck 1 round (replaces upstream prove_round as driven by ahp/prover.rs:199-207):
//   if r_host != nullptr: fold A,B,C (m_in entries each) with *r_host (HOST pointer: the challenge travels as a kernel
//   argument) into Ao,Bo,Co (m_in/2 entries), then
//   S(t) = sum_b E[b] (A(t,b) B(t,b) - C(t,b)), t = 0,1,2 over the (possibly folded) tables; o.out <- S.
void launch_sc1_round(const Fr* A, const Fr* B, const Fr* C, Fr* Ao, Fr* Bo, Fr* Co, const Fr* E, const Fr* r_host,
                      size_t m_in, const RoundOut& o, const RoundWs& ws, cudaStream_t stream);
// Sumcheck 2 round (ahp/prover.rs:258-266): S(t) = sum_b M(t,b) Z(t,b)
void launch_sc2_round(const Fr* M, const Fr* Z, Fr* Mo, Fr* Zo, const Fr* r_host, size_t m_in, const RoundOut& o,
                      const RoundWs& ws, cudaStream_t stream);
// out[k] = T_k[0] + r (T_k[1] - T_k[0]) for up to 3 two-entry tables (final fold; prover.rs:217-219)
void launch_final_fold3(const Fr* A, const Fr* B, const Fr* C, int ntab, const Fr* r_host, const RoundOut& o, cudaStream_t stream);
// open.rs:42-45: q[b] = in[2b+1] - in[2b]; r_out[b] = in[2b] + p (in[2b+1] - in[2b])
void launch_open_fold(const Fr* in, Fr* r_out, Fr* q_out, const Fr* p_dev, size_t half, cudaStream_t stream);
// q[b] = in[2b+1] - in[2b] alone (the quotient of the first opening round is independent of the point)
void launch_pair_diff(const Fr* in, Fr* q_out, size_t half, cudaStream_t stream);
// elementwise self-test helpers: out = a (op) b with the PTX path; op 0 add, 1 sub, 2 mul, 3 mul_portable
void launch_fr_binop(int op, const Fr* a, const Fr* b, Fr* out, size_t n, cudaStream_t stream);
void launch_fq_binop(int op, const Fq* a, const Fq* b, Fq* out, size_t n, cudaStream_t stream);
// integer-pipe microbenchmark: `iters` dependent Montgomery multiplications per thread
void launch_fr_mul_bench(Fr* inout, size_t n_threads, int iters, cudaStream_t stream);
void launch_fq_mul_bench(Fq* inout, size_t n_threads, int iters, cudaStream_t stream);
