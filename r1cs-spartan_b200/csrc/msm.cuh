// msm.cuh -- Pippenger bucket MSM over G1 / G2 with pre-shifted fixed bases, plus the fixed-base
// scalar multiplication used by keygen.
//
// Replaces upstream ark_ec::msm::VariableBaseMSM::multi_scalar_mul as called from
// src/commitment/commit.rs:25 (G1) and src/commitment/open.rs:49 (G2), and
// ark_ec::msm::FixedBaseMSM::multi_scalar_mul as called from src/commitment/setup.rs:61-70.
//
// Design (DESIGN.md "MSM"): the bases are the public parameters, fixed for the life of the handle,
// so at load time every base P_i is expanded into its window multiples 2^(c w) P_i (affine).  With
// those, every signed c-bit digit of every scalar lands in ONE shared set of 2^(c-1) buckets and
// the usual per-window Horner recombination (255 serial doublings) disappears: the MSM is
//     digits -> counting sort by (window, bucket) -> one thread per (window, bucket) accumulates
//     mixed additions -> windows merged per bucket -> sum_k k B_k by chunked running sums.
#pragma once
#include "common.cuh"

// Window layout of the 255-bit scalar: W windows, window w covers bits [shift[w], shift[w+1]).
// Widths are c or c-1 and as even as possible, and the TOP window is c-1 bits wide: every window is
// (nearly) fully populated, so signed digits spread evenly over the 2^(c-1) buckets, and the top
// digit (at most 0.906 * 2^(c-1) + carry, because r < 0.906 * 2^255) needs no carry-out window.
struct WinLayout {
    int W;
    int c;
    uint16_t shift[72];
};
WinLayout msm_layout(size_t m);

constexpr int MSM_MAX_LEVELS = 16;
constexpr int MSM_MAX_HALVINGS = 8;      // batched-affine pairwise rounds before the XYZZ accumulation
constexpr int MSM_INFO_WORDS = 32;       // entries, longest run, S0, levels, S1, R, halving totals[8] at 6, items[16] at 14, spare
constexpr int MSM_INFO_HTOT = 6, MSM_INFO_ITEMS = 14;

// Scratch of one MSM over one base table; kept with the table and reused by every proof (allocated on first
// use, grown on demand), so that the steady state performs no device allocation at all.
template <class F>
struct MsmScratch {
    DevBuf<uint32_t> codes, sorted, counts, offsets, cursors, info, perm, invperm;
    DevBuf<uint32_t> plan[MSM_MAX_LEVELS];
    DevBuf<uint32_t> hplan[MSM_MAX_HALVINGS];
    DevBuf<AffinePt<F>> affA, affB;        // outputs of the pairwise rounds (ping-pong)
    DevBuf<F> prefix;                      // per-thread prefix products of the simultaneous inversion
    DevBuf<XyzzPt<F>> ptsA, ptsB, block_out;
};

template <class F>
struct MsmBases {
    DevBuf<AffinePt<F>> tab;   // [W][m]: tab[w * m + i] = 2^(shift[w]) * P_i
    size_t m = 0;
    WinLayout lay{};
    mutable MsmScratch<F> scratch;
};

// One MSM in flight: phase A (digits, counting sort, run statistics) and phase B (chunked accumulation,
// bucket reduction) are separate so that a whole ladder of MSMs can be queued on several streams with a
// single host synchronisation in between.
template <class F>
struct MsmJob {
    const MsmBases<F>* bases = nullptr;
    const Fr* scalars = nullptr;
    size_t m = 0;
    XyzzPt<F>* out = nullptr;
    cudaStream_t stream = nullptr;
    uint32_t* info_host = nullptr;     // pinned, MSM_INFO_WORDS words (see k_scan_plan)
    bool top = false;                  // first (largest) level of an opening: its accumulation is profiled under its own name
    // Optional second (high-priority) stream + hand-over event: everything after the first accumulation level --
    // the latency-bound part of the job -- is queued there, so that it is not dispatched behind the
    // throughput-bound accumulation kernels of the other jobs; `stream` is joined to it again at the end.
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t tail_event = nullptr;
};
template <class F> void msm_begin(MsmJob<F>& job);
template <class F> void msm_finish(MsmJob<F>& job);   // job.stream must have been synchronised after msm_begin

// expand affine bases (device) into their window multiples
template <class F>
void msm_prepare(const AffinePt<F>* bases_dev, size_t m, MsmBases<F>& out, cudaStream_t stream);
// out_dev <- sum_i scalars[i] * P_i   (scalars: Montgomery Fr on the device; result XYZZ on the device)
template <class F>
void msm_run(const MsmBases<F>& bases, const Fr* scalars_dev, size_t m, XyzzPt<F>* out_dev, cudaStream_t stream);

// out[i] = scalars[i] * g for n scalars (device, Montgomery), affine results (device)
template <class F>
void fixed_base_mul(const AffinePt<F>& g_host, const Fr* scalars_dev, size_t n, AffinePt<F>* out_dev, cudaStream_t stream);
// XYZZ -> affine with simultaneous inversion
template <class F>
void batch_to_affine(const XyzzPt<F>* in_dev, AffinePt<F>* out_dev, size_t n, cudaStream_t stream);
