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

template <class F>
struct MsmBases {
    DevBuf<AffinePt<F>> tab;   // [W][m]: tab[w * m + i] = 2^(c w) * P_i
    size_t m = 0;
    int c = 0;                 // window bits
    int W = 0;                 // number of windows (c * W >= 256)
};

int msm_window_bits(size_t m);
inline int msm_num_windows(int c) { return 255 / c + 1; }

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
