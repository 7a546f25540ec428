// msm.cuh -- batched Pippenger bucket MSM over G1 / G2 with pre-shifted fixed bases, plus the fixed-base
// scalar multiplication used by keygen.
//
// Replaces upstream ark_ec::msm::VariableBaseMSM::multi_scalar_mul as called from
// src/commitment/commit.rs:25 (G1) and src/commitment/open.rs:49 (G2, once per level of an opening), and
// ark_ec::msm::FixedBaseMSM::multi_scalar_mul as called from src/commitment/setup.rs:61-70.
//
// Design (DESIGN.md "MSM"):
//  * The bases are the public parameters, fixed for the life of the handle, so at load time every base P_i is
//    expanded into its window multiples 2^(shift_w) P_i (affine).  With those, every signed digit of every scalar
//    lands in ONE shared set of 2^(c-1) buckets and the per-window Horner recombination (255 serial doublings)
//    disappears.
//  * An opening needs one MSM per level of the quotient pyramid (open.rs:37-51): sizes 2^(nv-1), 2^(nv-2), ..., 1.
//    They are independent and individually latency-bound below ~2^14 points, so they are run as ONE pipeline: an
//    MsmGroup is a list of SLOTS (one MSM each, own bases / window layout / output point) whose buckets are laid
//    side by side in one index space; digits -> plan -> scatter -> chunked accumulation -> bucket reduction each
//    run ONCE over all slots.  A proof is three such pipelines (commit, two openings) instead of 41 MSMs.
//  * No host synchronisation inside a pipeline: everything the later kernels need (entry count, number of
//    accumulation levels, chunk plans) is computed on the device by k_scan_plan and read from device memory; the
//    host launches grids sized by upper bounds it can compute from the slot sizes alone, and a fixed number of
//    accumulation levels (unused levels exit at once).
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
WinLayout msm_layout(size_t m, size_t group_max_m);   // group_max_m: the largest slot of the group the MSM belongs to

constexpr int MSM_MAX_SLOTS = 32;
constexpr int MSM_MAX_LEVELS = 8;        // accumulation levels a pipeline launches (levels the plan does not use exit at once)
constexpr int MSM_MAX_HALVINGS = 8;      // batched-affine pairwise rounds before the XYZZ accumulation
constexpr int MSM_INFO_WORDS = 32;       // entries, longest run, S0, levels, S1; items[MSM_MAX_LEVELS] at 8; round totals[MSM_MAX_HALVINGS] at 16
constexpr int MSM_INFO_ITEMS = 8, MSM_INFO_HTOT = 16;

// One MSM of a group: m points, its own window layout, its own range of the shared table / bucket / reduction spaces.
struct MsmSlot {
    uint32_t m;          // points (= scalars)
    uint32_t mbase;      // first global point index (prefix sum of m over the slots)
    uint32_t ebase;      // table: tab[ebase + w * m + i] = 2^(shift[w]) * P_i; the same index space numbers the entries
    uint32_t bbase;      // first bucket; the slot owns nb = 2^(c-1) buckets
    uint32_t nb;
    uint32_t red_l;      // buckets per quad of lanes in the first reduction stage
    uint32_t rbase;      // first CTA of the first reduction stage; the slot owns rblocks CTAs
    uint32_t rblocks;
    WinLayout lay;
};

// Scratch of one group; kept with the tables and reused by every proof (allocated when the group is prepared),
// so that the steady state performs no device allocation at all.
template <class F>
struct MsmScratch {
    DevBuf<uint32_t> codes, sorted, counts, offsets, cursors, info, perm, invperm;
    DevBuf<uint32_t> plan[MSM_MAX_LEVELS];
    DevBuf<uint32_t> cta_sum, cta_max, cta_hist, cta_lsum, cta_hsum;     // per-CTA aggregates of the plan kernels
    DevBuf<uint32_t> hplan[MSM_MAX_HALVINGS];
    DevBuf<AffinePt<F>> affA, affB;        // outputs of the pairwise rounds (ping-pong)
    DevBuf<F> prefix;                      // per-thread prefix products of the simultaneous inversion
    DevBuf<XyzzPt<F>> ptsA, ptsB, block_out;
};

template <class F>
struct MsmGroup {
    DevBuf<AffinePt<F>> tab;         // every slot's pre-shifted table, back to back (MsmSlot::ebase)
    std::vector<MsmSlot> slots;
    DevBuf<MsmSlot> slots_dev;
    uint32_t mtot = 0, etot = 0, btot = 0, rtot = 0;
    uint32_t red_quads = 64;         // quads of lanes per CTA of the bucket reduction
    uint32_t s0 = 0;                 // chunk length of the first accumulation level (fixed when the group is prepared)
    uint32_t items_bound[MSM_MAX_LEVELS] = {};   // upper bounds on the chunk count of every level (grid sizes)
    uint32_t R = 0;                  // pairwise affine rounds in front of the XYZZ accumulation (fixed when the group is prepared)
    uint32_t round_bound[MSM_MAX_HALVINGS] = {}, round_k[MSM_MAX_HALVINGS] = {}, round_threads[MSM_MAX_HALVINGS] = {};
    mutable MsmScratch<F> scratch;
    size_t nslots() const { return slots.size(); }
};

// per-slot scalar arrays of one run (Montgomery Fr on the device); passed to the digit kernel by value
struct MsmScalarPtrs { const Fr* p[MSM_MAX_SLOTS]; };

// Expand the affine bases of every slot (device arrays, bases_dev[j] has m[j] points) into the group's tables and
// allocate its scratch.  Synchronous with respect to `stream` only.
template <class F>
void msm_group_prepare(const std::vector<const AffinePt<F>*>& bases_dev, const std::vector<size_t>& m, MsmGroup<F>& out, cudaStream_t stream);
// Queue the whole pipeline on `stream`: out_dev[j] <- sum_i scalars[j][i] * P_{j,i} (XYZZ) for every slot j.
// Nothing is synchronised; the caller reads out_dev in stream order.
template <class F>
void msm_group_run(const MsmGroup<F>& g, const MsmScalarPtrs& scalars, XyzzPt<F>* out_dev, cudaStream_t stream, const char* tag = nullptr);

// The three phases of msm_group_run, for callers that pipeline several groups: a group's `accum` must follow its `front`
// and its `tail` its `accum` (in stream order or through events); different groups are independent of one another.
template <class F> void msm_group_front(const MsmGroup<F>& g, const MsmScalarPtrs& scalars, cudaStream_t stream);
template <class F> void msm_group_accum(const MsmGroup<F>& g, cudaStream_t stream);
template <class F> void msm_group_tail(const MsmGroup<F>& g, XyzzPt<F>* out_dev, cudaStream_t stream);

// out[i] = scalars[i] * g for n scalars (device, Montgomery), affine results (device)
template <class F>
void fixed_base_mul(const AffinePt<F>& g_host, const Fr* scalars_dev, size_t n, AffinePt<F>* out_dev, cudaStream_t stream);
// XYZZ -> affine with simultaneous inversion
template <class F>
void batch_to_affine(const XyzzPt<F>* in_dev, AffinePt<F>* out_dev, size_t n, cudaStream_t stream);
