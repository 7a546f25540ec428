// msm.cu -- bucket-method MSM over pre-shifted fixed bases (G1 over Fq, G2 over Fq2), fixed-base
// scalar multiplication and batched XYZZ -> affine conversion.  See msm.cuh for the design.
// Integer-pipe bound (Fq Montgomery products); bases are gathered with 128-bit loads.
#include "msm.cuh"

// ------------------------------------------------------------------ window geometry
WinLayout msm_layout(size_t m) {
    int lg = 0;
    while (((size_t)1 << lg) < m) lg++;
    // window bits: log2(m) - 3 balances the accumulation (m W additions) against the bucket reduction (~9 additions
    // per bucket today).  SB_MSM_C_OFFSET / SB_MSM_C_MAX shift the rule for experiments (a cheaper reduction would
    // favour one more bit); the layout is fixed when the bases are prepared, so the knobs only act at load time.
    static const int c_off = getenv("SB_MSM_C_OFFSET") ? atoi(getenv("SB_MSM_C_OFFSET")) : -3;
    static const int c_max = getenv("SB_MSM_C_MAX") ? atoi(getenv("SB_MSM_C_MAX")) : 16;
    int c = lg + c_off;
    if (c < 4) c = 4;
    if (c > c_max) c = c_max;
    if (c > 20) c = 20;
    WinLayout L{};
    L.c = c;
    const int top = c - 1, rest = 255 - top;
    const int wl = (rest + c - 1) / c;                 // lower windows
    const int base = rest / wl, extra = rest % wl;     // widths: `extra` windows of base+1, the others base (>= c-1)
    int s = 0;
    for (int w = 0; w < wl; w++) { L.shift[w] = (uint16_t)s; s += base + (w < extra ? 1 : 0); }
    L.shift[wl] = (uint16_t)s;                          // top window
    L.shift[wl + 1] = 255;
    L.W = wl + 1;
    return L;
}

// ------------------------------------------------------------------ digits + counting sort
constexpr uint32_t CODE_NONE = 0xffffffffu;

__global__ void __launch_bounds__(256) k_msm_digits(const Fr* __restrict__ scalars, size_t m, WinLayout lay,
                                                    uint32_t* __restrict__ codes, uint32_t* __restrict__ counts) {
    const int W = lay.W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) {
        Fr s = ldg_elem(&scalars[i]).to_canonical();          // "into_repr" (commit.rs:20-21, open.rs:46)
        uint32_t limb[9];
#pragma unroll
        for (int k = 0; k < 8; k++) limb[k] = s.l[k];
        limb[8] = 0;
        uint32_t carry = 0;
        for (int w = 0; w < W; w++) {
            const int bit = lay.shift[w], width = lay.shift[w + 1] - bit;
            const int lo = bit >> 5, sh = bit & 31;
            uint64_t v = (uint64_t)limb[lo] | ((uint64_t)limb[lo + 1] << 32);
            uint32_t d = ((uint32_t)(v >> sh) & ((1u << width) - 1)) + carry;
            uint32_t code;
            if (w + 1 < W && d > (1u << (width - 1))) {       // negative digit d - 2^width, borrow from the next window
                uint32_t nd = (1u << width) - d; carry = 1;
                code = nd ? ((nd - 1) | 0x80000000u) : CODE_NONE;
            } else {                                           // the top window is never recoded (see WinLayout)
                carry = 0; code = d ? (d - 1) : CODE_NONE;
            }
            codes[(size_t)w * m + i] = code;
            if (code != CODE_NONE) atomicAdd(&counts[code & 0x7fffffffu], 1u);
        }
    }
}

// Chunking of the bucket runs.  Level 0 (mixed additions of table entries, the throughput-bound part) works on
// chunks of S0 entries; every further level sums chunks of S1 partial sums of the level below, until each
// bucket is down to one point: levels = 1 + ceil(log_S1(ceil(maxrun / S0))).  S0 trades the number of partial
// sums (entries / S0 full additions, 1.4x the cost of a mixed one, and as many 384-byte points written) against
// the length of the dependent chain per thread; S1 is small because the later levels are pure latency: a run of
// r partial sums costs S1 * log_S1(r) dependent additions (r = 2048: 21 with S1 = 3, 90 with two levels of 45;
// measured at 2^17 constraints: S1 = 3: 15.7 ms, 4: 16.2, 8: 16.9, the old two/three equal levels: 18.0).
__host__ __device__ inline uint32_t msm_levels(uint32_t maxrun, uint32_t s0, uint32_t s1) {
    uint32_t levels = 1;
    uint64_t cover = s0;
    while (cover < maxrun && levels < (uint32_t)MSM_MAX_LEVELS) { cover *= s1; levels++; }
    return levels;
}

__device__ inline uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* sh, uint32_t& total) {
    const uint32_t tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t t = tid >= off ? sh[tid - off] : 0;
        __syncthreads();
        sh[tid] += t;
        __syncthreads();
    }
    uint32_t incl = sh[tid];
    total = sh[1023];
    __syncthreads();
    return incl - v;
}

struct PlanPtrs {
    uint32_t* offsets; uint32_t* cursors; uint32_t* info;
    uint32_t* perm; uint32_t* invperm;          // bucket order of the accumulation (nullptr: natural order)
    uint32_t* plan[MSM_MAX_LEVELS];
    uint32_t* hplan[MSM_MAX_HALVINGS];
};
constexpr uint32_t AFF_TARGET_RUN = 16;      // pairwise rounds run until the longest bucket run is at most this

// One CTA: exclusive scan of the B bucket counters (-> offsets, cursors), the run statistics, and every plan
// the accumulation needs, all chosen on the device:
//   R        pairwise (batched-affine) rounds: 0 when the MSM has fewer than aff_min entries, else the
//            smallest R with ceil(maxrun / 2^R) <= AFF_TARGET_RUN
//   hplan[r-1][b] = exclusive scan of ceil(count_b / 2^r), r = 1..R      (layout of the list after round r)
//   levels        chunking of the runs that remain after the R rounds (counts c'_b = ceil(count_b / 2^R))
//   plan[l][b]    = exclusive scan of ceil(c'_b / (S0 S1^l))              (chunk plan of accumulation level l)
// info: see MSM_INFO_WORDS.
__global__ void __launch_bounds__(1024) k_scan_plan(const uint32_t* __restrict__ counts, uint32_t B, PlanPtrs pp, uint32_t s0, uint32_t s1,
                                                    uint32_t aff_min) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t sh_max, sh_levels, sh_R;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) sh_max = 0;
    const uint32_t per = (B + 1023) / 1024;
    const uint32_t beg = min(tid * per, B), end = min(beg + per, B);
    uint32_t sum = 0, mx = 0;
    for (uint32_t i = beg; i < end; i++) { uint32_t cnt = counts[i]; sum += cnt; mx = max(mx, cnt); }
    __syncthreads();
    atomicMax(&sh_max, mx);
    uint32_t total;
    uint32_t run = block_exclusive_scan_1024(sum, sh, total);
    for (uint32_t i = beg; i < end; i++) { pp.offsets[i] = run; pp.cursors[i] = run; run += counts[i]; }
    if (tid == 0) {
        pp.offsets[B] = total;
        uint32_t R = 0;
        if (total >= aff_min) while (R < (uint32_t)MSM_MAX_HALVINGS && ((sh_max + (1u << R) - 1) >> R) > AFF_TARGET_RUN) R++;
        sh_R = R;
        pp.info[0] = total; pp.info[1] = sh_max; pp.info[2] = s0; pp.info[4] = s1; pp.info[5] = R;
    }
    __syncthreads();
    const uint32_t R = sh_R;
    for (uint32_t r = 1; r <= R; r++) {
        const uint32_t add = (1u << r) - 1;
        uint32_t s = 0;
        for (uint32_t i = beg; i < end; i++) s += (counts[i] + add) >> r;
        uint32_t tot;
        uint32_t x = block_exclusive_scan_1024(s, sh, tot);
        for (uint32_t i = beg; i < end; i++) { pp.hplan[r - 1][i] = x; x += (counts[i] + add) >> r; }
        if (tid == 0) { pp.hplan[r - 1][B] = tot; pp.info[MSM_INFO_HTOT + r - 1] = tot; }
    }
    if (tid == 0) {
        sh_levels = msm_levels((sh_max + (1u << R) - 1) >> R, s0, s1);
        pp.info[3] = sh_levels;
    }
    __syncthreads();
    const uint32_t levels = sh_levels, radd = (1u << R) - 1;
    // Accumulation order of the buckets: by decreasing chunk length of the first level (counting sort; key =
    // ceil(c'_b / ceil(c'_b / S0)) <= S0).  Thread p of the accumulation kernel takes chunk p in this order, so
    // the 32 lanes of a warp run loops of (almost) equal length instead of whatever neighbouring buckets hold,
    // and the shortest chunks -- not the longest -- are the ones left when the grid drains.
    __shared__ uint32_t sh_hist[1026];
    const bool sorted = pp.perm != nullptr && s0 <= 1024;
    if (sorted) {
        for (uint32_t k = tid; k <= s0; k += 1024) sh_hist[k] = 0;
        __syncthreads();
        for (uint32_t i = beg; i < end; i++) {
            const uint32_t c = (counts[i] + radd) >> R, nch = (c + s0 - 1) / s0;
            atomicAdd(&sh_hist[nch ? (c + nch - 1) / nch : 0], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t run2 = 0;
            for (int k = (int)s0; k >= 0; k--) { const uint32_t h = sh_hist[k]; sh_hist[k] = run2; run2 += h; }
        }
        __syncthreads();
        for (uint32_t i = beg; i < end; i++) {
            const uint32_t c = (counts[i] + radd) >> R, nch = (c + s0 - 1) / s0;
            const uint32_t pos = atomicAdd(&sh_hist[nch ? (c + nch - 1) / nch : 0], 1u);
            pp.perm[pos] = i; pp.invperm[i] = pos;
        }
        __syncthreads();
    }
    uint64_t div = s0;
    for (uint32_t l = 0; l < levels; l++, div *= s1) {
        uint32_t s = 0;
        for (uint32_t k = beg; k < end; k++) {
            const uint32_t i = sorted ? pp.perm[k] : k;
            s += (uint32_t)((((counts[i] + radd) >> R) + div - 1) / div);
        }
        uint32_t tot;
        uint32_t x = block_exclusive_scan_1024(s, sh, tot);
        for (uint32_t k = beg; k < end; k++) {
            const uint32_t i = sorted ? pp.perm[k] : k;
            pp.plan[l][k] = x; x += (uint32_t)((((counts[i] + radd) >> R) + div - 1) / div);
        }
        if (tid == 0) { pp.plan[l][B] = tot; pp.info[MSM_INFO_ITEMS + l] = tot; }
    }
}

__global__ void __launch_bounds__(256) k_msm_scatter(const uint32_t* __restrict__ codes, size_t total, uint32_t* __restrict__ cursors,
                                                     uint32_t* __restrict__ sorted) {
    for (size_t f = blockIdx.x * (size_t)blockDim.x + threadIdx.x; f < total; f += (size_t)gridDim.x * blockDim.x) {
        uint32_t code = codes[f];
        if (code == CODE_NONE) continue;
        uint32_t pos = atomicAdd(&cursors[code & 0x7fffffffu], 1u);
        sorted[pos] = (uint32_t)f | (code & 0x80000000u);      // f = w * m + i indexes the pre-shifted table
    }
}

// ------------------------------------------------------------------ bucket accumulation
// Bucket b's entries (all windows: the bases are pre-shifted) are one contiguous run of `sorted`.
// Runs can be extremely uneven -- the top window only sees the few high bits of a 255-bit scalar, so
// all of its digits fall into a handful of buckets -- so the work item is not a bucket but a CHUNK of at
// most S consecutive entries of one run.  Level 1 turns chunks of base indices into partial sums (mixed
// additions); every further level sums chunks of the previous level's partial sums, until each bucket
// is down to one point.  Work per thread is bounded by S at every level whatever the scalars are.

// one thread per chunk: out[p] = sum of the chunk's elements
constexpr int ACC_THREADS = 64;
// MODE 0: sum XYZZ partial sums (in_pts);  1: mixed additions of table entries (sorted -> tab, sign in bit 31);
// MODE 2: mixed additions of an affine list (aff_in), the output of the pairwise rounds
template <class F, int MODE>
__global__ void __launch_bounds__(ACC_THREADS) k_seg_accum(const AffinePt<F>* __restrict__ tab, const uint32_t* __restrict__ sorted,
                                                           const AffinePt<F>* __restrict__ aff_in, const XyzzPt<F>* __restrict__ in_pts,
                                                           const uint32_t* __restrict__ seg_off, const uint32_t* __restrict__ chunk_start,
                                                           uint32_t nseg, uint32_t S, const uint32_t* __restrict__ perm,
                                                           XyzzPt<F>* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= chunk_start[nseg]) return;
    uint32_t lo = 0, hi = nseg;                 // last b with chunk_start[b] <= p (skips empty runs)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&chunk_start[mid]) <= p) lo = mid; else hi = mid;
    }
    // balanced split of the run into its planned number of chunks (each at most S long): the threads of a
    // warp get chunks of nearly equal length instead of S, S, ..., remainder
    const uint32_t j = p - chunk_start[lo], nch = chunk_start[lo + 1] - chunk_start[lo];
    const uint32_t sb = perm ? __ldg(&perm[lo]) : lo;        // position in the accumulation order -> bucket (first level only)
    const uint32_t off = seg_off[sb], cnt = seg_off[sb + 1] - off;
    uint32_t beg, end;
    if (S & 0x80000000u) {                      // experiment knob SB_MSM_BALANCED=0: chunks of S, S, ..., remainder
        beg = off + j * (S & 0x7fffffffu); end = min(beg + (S & 0x7fffffffu), off + cnt);
    } else {
        beg = off + (uint32_t)(((uint64_t)j * cnt) / nch);
        end = off + (uint32_t)(((uint64_t)(j + 1) * cnt) / nch);
    }
    XyzzPt<F> acc = XyzzPt<F>::inf();
    for (uint32_t e = beg; e < end; e++) {
        if (MODE == 1) {
            uint32_t ent = __ldg(&sorted[e]);
            AffinePt<F> q = ldg_elem(&tab[ent & 0x7fffffffu]);
            if (ent & 0x80000000u) q.y = F::neg(q.y);
            acc = XyzzPt<F>::add_mixed(acc, q);
        } else if (MODE == 2) {
            acc = XyzzPt<F>::add_mixed(acc, ldg_elem(&aff_in[e]));
        } else {
            acc = XyzzPt<F>::add(acc, ldg_elem(&in_pts[e]));
        }
    }
    st_elem(&out[p], acc);
}

// ------------------------------------------------------------------ pairwise rounds in affine coordinates
// One round halves every bucket run: output element j of bucket b is the sum of input elements 2j and 2j+1 of
// that bucket (or a copy of element 2j when the run is odd).  An affine addition is one inversion plus 2M + 1S;
// each thread owns AFF_K consecutive output elements and shares ONE inversion among them (Montgomery's trick:
// +3M per element), so an addition costs about 5M + 1S + (one Fermat inversion) / AFF_K instead of the 8M + 2S of
// a mixed XYZZ addition.  The exceptional cases of the group law (an input at infinity, P = Q, P = -Q) are
// handled exactly; their denominator is replaced by 1 so that the shared inversion stays well defined.
constexpr int AFF_K = 128;
constexpr int AFF_THREADS = 64;
enum { PAIR_COPY_P = 0, PAIR_COPY_Q = 1, PAIR_ADD = 2, PAIR_DBL = 3, PAIR_INF = 4 };

template <class F>
__device__ __forceinline__ int pair_classify(const AffinePt<F>& P, const AffinePt<F>& Q, bool has2, F& d) {
    d = F::one();
    if (!has2 || Q.is_inf()) return PAIR_COPY_P;
    if (P.is_inf()) return PAIR_COPY_Q;
    if (P.x == Q.x) {
        if (P.y == Q.y && !P.y.is_zero()) { d = F::dbl(P.y); return PAIR_DBL; }
        return PAIR_INF;
    }
    d = F::sub(Q.x, P.x);
    return PAIR_ADD;
}

template <class F, bool FIRST>
__global__ void __launch_bounds__(AFF_THREADS) k_affine_round(const AffinePt<F>* __restrict__ tab, const uint32_t* __restrict__ sorted,
                                                              const AffinePt<F>* __restrict__ in_aff, const uint32_t* __restrict__ in_off,
                                                              const uint32_t* __restrict__ out_off, uint32_t B, uint32_t nthreads,
                                                              F* __restrict__ prefix, AffinePt<F>* __restrict__ out_aff) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = out_off[B];
    const uint32_t p0 = t * (uint32_t)AFF_K;
    if (t >= nthreads || p0 >= total) return;
    const uint32_t p1 = min(p0 + (uint32_t)AFF_K, total);
    auto load_in = [&](uint32_t idx) -> AffinePt<F> {
        if (FIRST) {
            uint32_t ent = __ldg(&sorted[idx]);
            AffinePt<F> q = ldg_elem(&tab[ent & 0x7fffffffu]);
            if (ent & 0x80000000u) q.y = F::neg(q.y);
            return q;
        }
        return ldg_elem(&in_aff[idx]);
    };
    uint32_t lo = 0, hi = B;                    // bucket of p0: last b with out_off[b] <= p0
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&out_off[mid]) <= p0) lo = mid; else hi = mid;
    }
    // pass 1: prefix products of the denominators
    uint32_t b = lo;
    F acc = F::one();
    for (uint32_t p = p0; p < p1; p++) {
        while (__ldg(&out_off[b + 1]) <= p) b++;
        const uint32_t j = p - __ldg(&out_off[b]);
        const uint32_t ib = __ldg(&in_off[b]), n_in = __ldg(&in_off[b + 1]) - ib;
        const bool has2 = 2 * j + 1 < n_in;
        AffinePt<F> P = load_in(ib + 2 * j), Q = has2 ? load_in(ib + 2 * j + 1) : P;
        F d;
        pair_classify(P, Q, has2, d);
        st_elem(&prefix[(size_t)(p - p0) * nthreads + t], acc);
        acc = F::mul(acc, d);
    }
    F inv = F::inv(acc);
    // pass 2, backwards: 1/d_k = inv * prefix_k, then inv *= d_k
    for (uint32_t p = p1; p-- > p0;) {
        while (__ldg(&out_off[b]) > p) b--;
        const uint32_t j = p - __ldg(&out_off[b]);
        const uint32_t ib = __ldg(&in_off[b]), n_in = __ldg(&in_off[b + 1]) - ib;
        const bool has2 = 2 * j + 1 < n_in;
        AffinePt<F> P = load_in(ib + 2 * j), Q = has2 ? load_in(ib + 2 * j + 1) : P;
        F d;
        const int kind = pair_classify(P, Q, has2, d);
        F dinv = F::mul(inv, ldg_elem(&prefix[(size_t)(p - p0) * nthreads + t]));
        inv = F::mul(inv, d);
        AffinePt<F> r;
        if (kind == PAIR_COPY_P) r = P;
        else if (kind == PAIR_COPY_Q) r = Q;
        else if (kind == PAIR_INF) r = AffinePt<F>::inf();
        else {
            F lam;
            if (kind == PAIR_ADD) lam = F::mul(F::sub(Q.y, P.y), dinv);
            else { F xx = F::sqr(P.x); lam = F::mul(F::add(F::dbl(xx), xx), dinv); }
            r.x = F::sub(F::sub(F::sqr(lam), P.x), Q.x);
            r.y = F::sub(F::mul(lam, F::sub(P.x, r.x)), P.y);
        }
        st_elem(&out_aff[p], r);
    }
}

// ---- cooperative variant (SB_MSM_AFFINE_COOP=1; prepared in round 1, NOT yet run on a GPU, hence off and not in the
// test matrix): the inversion is shared by the whole CTA instead of one per thread.  Thread t owns K consecutive
// output elements and reduces their denominators to one product a_t exactly as above; the CTA then computes the
// inclusive prefix products I_t = a_0..a_t and suffix products U_t = a_t..a_{T-1} (Hillis-Steele in shared memory,
// log2 T products per thread each), ONE thread inverts the total, and 1/a_t = total^-1 * I_{t-1} * U_{t+1}.  The
// inversion share per addition drops from (Fermat inversion) / K to (2 log2 T + 2 products) / K plus one inversion
// per T K additions, so K can be small (more threads, shorter chains): about 5M + 1S + 0.3M per addition against
// the 8M + 2S of the mixed XYZZ addition.
constexpr int AFFC_THREADS = 256;

template <class F>
__device__ F cta_batch_inverse(const F& a, F* sh) {
    const uint32_t t = threadIdx.x, T = blockDim.x;
    __shared__ F sh_total_inv;
    sh[t] = a;
    __syncthreads();
    for (uint32_t off = 1; off < T; off <<= 1) {               // inclusive prefix products
        F v = sh[t];
        if (t >= off) v = F::mul(sh[t - off], v);
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    const F before = t ? sh[t - 1] : F::one();
    if (t == 0) sh_total_inv = F::inv(sh[T - 1]);
    __syncthreads();
    sh[t] = a;
    __syncthreads();
    for (uint32_t off = 1; off < T; off <<= 1) {               // inclusive suffix products
        F v = sh[t];
        if (t + off < T) v = F::mul(v, sh[t + off]);
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    const F after = (t + 1 < T) ? sh[t + 1] : F::one();
    const F r = F::mul(F::mul(sh_total_inv, before), after);
    __syncthreads();                                            // sh is reused by the caller's next call
    return r;
}

template <class F, bool FIRST>
__global__ void __launch_bounds__(AFFC_THREADS) k_affine_round_coop(const AffinePt<F>* __restrict__ tab, const uint32_t* __restrict__ sorted,
                                                                    const AffinePt<F>* __restrict__ in_aff, const uint32_t* __restrict__ in_off,
                                                                    const uint32_t* __restrict__ out_off, uint32_t B, uint32_t nthreads, uint32_t K,
                                                                    F* __restrict__ prefix, AffinePt<F>* __restrict__ out_aff) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* sh = reinterpret_cast<F*>(smem_raw);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = out_off[B];
    const uint32_t p0 = t * K;
    const bool active = t < nthreads && p0 < total;             // inactive threads still take part in the barriers
    const uint32_t p1 = active ? min(p0 + K, total) : p0;
    auto load_in = [&](uint32_t idx) -> AffinePt<F> {
        if (FIRST) {
            uint32_t ent = __ldg(&sorted[idx]);
            AffinePt<F> q = ldg_elem(&tab[ent & 0x7fffffffu]);
            if (ent & 0x80000000u) q.y = F::neg(q.y);
            return q;
        }
        return ldg_elem(&in_aff[idx]);
    };
    uint32_t b = 0;
    if (active) {
        uint32_t lo = 0, hi = B;                                // bucket of p0: last b with out_off[b] <= p0
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&out_off[mid]) <= p0) lo = mid; else hi = mid;
        }
        b = lo;
    }
    F acc = F::one();
    for (uint32_t p = p0; p < p1; p++) {                        // pass 1: prefix products of this thread's denominators
        while (__ldg(&out_off[b + 1]) <= p) b++;
        const uint32_t j = p - __ldg(&out_off[b]);
        const uint32_t ib = __ldg(&in_off[b]), n_in = __ldg(&in_off[b + 1]) - ib;
        const bool has2 = 2 * j + 1 < n_in;
        AffinePt<F> P = load_in(ib + 2 * j), Q = has2 ? load_in(ib + 2 * j + 1) : P;
        F d;
        pair_classify(P, Q, has2, d);
        st_elem(&prefix[(size_t)(p - p0) * nthreads + t], acc);
        acc = F::mul(acc, d);
    }
    F inv = cta_batch_inverse<F>(acc, sh);                      // 1 / (product of this thread's denominators)
    for (uint32_t p = p1; p-- > p0;) {                          // pass 2, backwards
        while (__ldg(&out_off[b]) > p) b--;
        const uint32_t j = p - __ldg(&out_off[b]);
        const uint32_t ib = __ldg(&in_off[b]), n_in = __ldg(&in_off[b + 1]) - ib;
        const bool has2 = 2 * j + 1 < n_in;
        AffinePt<F> P = load_in(ib + 2 * j), Q = has2 ? load_in(ib + 2 * j + 1) : P;
        F d;
        const int kind = pair_classify(P, Q, has2, d);
        F dinv = F::mul(inv, ldg_elem(&prefix[(size_t)(p - p0) * nthreads + t]));
        inv = F::mul(inv, d);
        AffinePt<F> r;
        if (kind == PAIR_COPY_P) r = P;
        else if (kind == PAIR_COPY_Q) r = Q;
        else if (kind == PAIR_INF) r = AffinePt<F>::inf();
        else {
            F lam;
            if (kind == PAIR_ADD) lam = F::mul(F::sub(Q.y, P.y), dinv);
            else { F xx = F::sqr(P.x); lam = F::mul(F::add(F::dbl(xx), xx), dinv); }
            r.x = F::sub(F::sub(F::sqr(lam), P.x), Q.x);
            r.y = F::sub(F::mul(lam, F::sub(P.x, r.x)), P.y);
        }
        st_elem(&out_aff[p], r);
    }
}

template <class F>
__device__ XyzzPt<F> mul_small(const XyzzPt<F>& p, uint32_t k) {
    XyzzPt<F> acc = XyzzPt<F>::inf();
    for (int bit = 31 - __clz(k | 1); bit >= 0; bit--) {
        acc = XyzzPt<F>::dbl(acc);
        if ((k >> bit) & 1) acc = XyzzPt<F>::add(acc, p);
    }
    return k ? acc : XyzzPt<F>::inf();
}

constexpr int RED_THREADS = 64;
// tree-sum RED_THREADS points held one per thread; result valid in thread 0
template <class F>
__device__ XyzzPt<F> block_tree_sum(XyzzPt<F> v, XyzzPt<F>* sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int stride = RED_THREADS / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) sh[threadIdx.x] = XyzzPt<F>::add(sh[threadIdx.x], sh[threadIdx.x + stride]);
        __syncthreads();
    }
    return sh[0];
}
// stage 1: thread t owns merged buckets [t L, (t+1) L): sum_b (b+1) M_b = running sums + (t L) * (sum of M_b)
template <class F>
__global__ void __launch_bounds__(RED_THREADS) k_bucket_reduce1(const XyzzPt<F>* __restrict__ pts, const uint32_t* __restrict__ off,
                                                                const uint32_t* __restrict__ invperm, uint32_t B, uint32_t L,
                                                                XyzzPt<F>* __restrict__ block_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XyzzPt<F>* sh = reinterpret_cast<XyzzPt<F>*>(smem_raw);
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)t * L;
    XyzzPt<F> total = XyzzPt<F>::inf();
    if (lo < B) {
        uint32_t hi = (uint32_t)min((uint64_t)B, lo + L);
        XyzzPt<F> run = XyzzPt<F>::inf(), sum = XyzzPt<F>::inf();
        for (uint32_t b = hi; b-- > (uint32_t)lo;) {
            const uint32_t k = invperm ? __ldg(&invperm[b]) : b;      // bucket -> position in the accumulation order
            uint32_t o = __ldg(&off[k]);
            if (__ldg(&off[k + 1]) > o) run = XyzzPt<F>::add(run, ldg_elem(&pts[o]));
            sum = XyzzPt<F>::add(sum, run);
        }
        total = XyzzPt<F>::add(sum, mul_small(run, (uint32_t)lo));
    }
    XyzzPt<F> r = block_tree_sum(total, sh);
    if (threadIdx.x == 0) st_elem(&block_out[blockIdx.x], r);
}
template <class F>
__global__ void __launch_bounds__(RED_THREADS) k_bucket_reduce2(const XyzzPt<F>* __restrict__ block_out, uint32_t nblocks, XyzzPt<F>* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XyzzPt<F>* sh = reinterpret_cast<XyzzPt<F>*>(smem_raw);
    XyzzPt<F> acc = XyzzPt<F>::inf();
    for (uint32_t i = threadIdx.x; i < nblocks; i += blockDim.x) acc = XyzzPt<F>::add(acc, block_out[i]);
    XyzzPt<F> r = block_tree_sum(acc, sh);
    if (threadIdx.x == 0) st_elem(out, r);
}

// ------------------------------------------------------------------ base expansion / affine conversion
template <class F>
__global__ void __launch_bounds__(128) k_preshift(const AffinePt<F>* __restrict__ bases, size_t i0, size_t cnt, WinLayout lay,
                                                  XyzzPt<F>* __restrict__ tmp) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    XyzzPt<F> p = XyzzPt<F>::from_affine(ldg_elem(&bases[i0 + j]));
    st_elem(&tmp[j], p);
    for (int w = 1; w < lay.W; w++) {
        for (int k = lay.shift[w - 1]; k < lay.shift[w]; k++) p = XyzzPt<F>::dbl(p);
        st_elem(&tmp[(size_t)w * cnt + j], p);
    }
}

constexpr int BTA_K = 8;
// flat element f of `in` goes to out[(f / cnt) * m + i0 + (f % cnt)]
template <class F>
__global__ void __launch_bounds__(128) k_batch_to_affine(const XyzzPt<F>* __restrict__ in, AffinePt<F>* __restrict__ out, size_t n,
                                                         size_t cnt, size_t m, size_t i0) {
    size_t base = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * BTA_K;
    if (base >= n) return;
    F pre[BTA_K];
    F acc = F::one();
#pragma unroll 1
    for (int k = 0; k < BTA_K; k++) {
        pre[k] = acc;
        size_t f = base + k;
        if (f < n) {
            F zzz = ldg_elem(&in[f].ZZZ);
            if (!zzz.is_zero()) acc = F::mul(acc, zzz);
        }
    }
    F inv = F::inv(acc);
#pragma unroll 1
    for (int k = BTA_K - 1; k >= 0; k--) {
        size_t f = base + k;
        if (f >= n) continue;
        XyzzPt<F> p = ldg_elem(&in[f]);
        AffinePt<F> a;
        if (p.ZZZ.is_zero()) {
            a = AffinePt<F>::inf();
        } else {
            F zi3 = F::mul(inv, pre[k]);
            inv = F::mul(inv, p.ZZZ);
            F zi2 = F::mul(F::sqr(zi3), F::sqr(p.ZZ));
            a.x = F::mul(p.X, zi2);
            a.y = F::mul(p.Y, zi3);
        }
        st_elem(&out[(f / cnt) * m + i0 + (f % cnt)], a);
    }
}

template <class F>
void batch_to_affine(const XyzzPt<F>* in_dev, AffinePt<F>* out_dev, size_t n, cudaStream_t stream) {
    if (!n) return;
    size_t threads = (n + BTA_K - 1) / BTA_K;
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_batch_to_affine"), (k_batch_to_affine<F>), (int)((threads + 127) / 128), 128, 0, stream, in_dev, out_dev, n, n, n, (size_t)0);
}

template <class F>
void msm_prepare(const AffinePt<F>* bases_dev, size_t m, MsmBases<F>& out, cudaStream_t stream) {
    out.m = m; out.lay = msm_layout(m);
    const int W = out.lay.W;
    out.tab.alloc((size_t)W * m, stream);
    const size_t chunk = m < ((size_t)1 << 16) ? m : ((size_t)1 << 16);
    DevBuf<XyzzPt<F>> tmp((size_t)W * chunk, stream);
    for (size_t i0 = 0; i0 < m; i0 += chunk) {
        size_t cnt = m - i0 < chunk ? m - i0 : chunk;
        SB_LAUNCH_NAMED(SB_KNAME(F, "k_preshift"), (k_preshift<F>), (int)((cnt + 127) / 128), 128, 0, stream, bases_dev, i0, cnt, out.lay, tmp.get());
        size_t n = (size_t)W * cnt;
        size_t threads = (n + BTA_K - 1) / BTA_K;
        SB_LAUNCH_NAMED(SB_KNAME(F, "k_batch_to_affine"), (k_batch_to_affine<F>), (int)((threads + 127) / 128), 128, 0, stream, tmp.get(), out.tab.get(), n, cnt, m, i0);
    }
}

// Chunk sizes of the accumulation levels (see msm_levels); SB_MSM_S0 / SB_MSM_S1 override
static uint32_t msm_env_u32(const char* name, uint32_t dflt, uint32_t lo, uint32_t hi) {
    const char* e = getenv(name);
    long v = e ? atol(e) : (long)dflt;
    return (uint32_t)std::min<long>(std::max<long>(v, lo), hi);
}
// S0 by size: throughput-bound jobs take long chunks (fewer partial sums), latency-bound ones short chains
static uint32_t msm_s0(size_t entries) {
    static const uint32_t v = msm_env_u32("SB_MSM_S0", 0, 0, 1024);
    static const uint32_t big = msm_env_u32("SB_MSM_S0_BIG", 48, 2, 1024), small = msm_env_u32("SB_MSM_S0_SMALL", 24, 2, 1024);
    return v >= 2 ? v : (entries >= ((size_t)1 << 21) ? big : small);
}
// Accumulate the buckets in the order of decreasing first-level chunk length (see k_scan_plan): the lanes of a warp
// then run loops of equal length.  Measured: 2^20 constraints 64.3 -> 59.8 ms, 2^17: 15.8 -> 15.1 ms.  SB_MSM_SORTED=0
// restores the natural bucket order.
static bool msm_sorted() { static const bool v = msm_env_u32("SB_MSM_SORTED", 1, 0, 1) != 0; return v; }
static uint32_t msm_s1() { static const uint32_t v = msm_env_u32("SB_MSM_S1", 3, 2, 4096); return v; }

template <class T>
static inline void ensure(DevBuf<T>& b, size_t n, cudaStream_t s) { if (b.n < n) b.alloc(n, s); }

// MSMs with at least 2^SB_MSM_AFFINE_LOG2 entries run the pairwise affine rounds first (G2 only: over Fq the shared
// inversion costs more than it saves).  OFF by default: measured at 2^20 constraints on B200 the rounds LOSE
// (opening 34.3 ms against 26.2 ms) -- a thread's chain of AFF_K additions plus one Fermat inversion (~2e5
// instructions) is milliseconds long, and the later rounds have too few threads to fill 148 SMs; a smaller AFF_K
// makes the inversion share larger than the saving.  Kept (and parity-tested with the threshold forced down)
// for instances large enough to amortise it and as the base for a cheaper inversion (k_affine_round_coop,
// SB_MSM_AFFINE_COOP=1: one inversion per CTA; written at the end of round 1, not yet run on a GPU).
template <class F>
static uint32_t msm_affine_min() {
    static const uint32_t v = [] {
        const char* e = getenv("SB_MSM_AFFINE_LOG2");
        int lg = e ? atoi(e) : 0;
        return (lg <= 0 || lg >= 32) ? 0xffffffffu : (1u << lg);
    }();
    return sizeof(F) == sizeof(Fq2) ? v : 0xffffffffu;
}

template <class F>
void msm_begin(MsmJob<F>& job) {
    const MsmBases<F>& bases = *job.bases;
    MsmScratch<F>& sc = bases.scratch;
    cudaStream_t stream = job.stream;
    const size_t m = job.m;
    SB_REQUIRE(m == bases.m, "msm: scalar count does not match the prepared bases");
    if (g_sb_prof_on) { g_sb_prof_tag = 0; while (((size_t)2 << g_sb_prof_tag) <= m) g_sb_prof_tag++; }
    const uint32_t B = 1u << (bases.lay.c - 1);
    const size_t total = (size_t)bases.lay.W * m;             // upper bound on the number of entries
    SB_REQUIRE(total < ((size_t)1 << 31), "msm: too many (window, point) pairs for 31-bit table indices");
    ensure(sc.codes, total, stream); ensure(sc.sorted, total, stream);
    ensure(sc.counts, B, stream); ensure(sc.offsets, B + 1, stream); ensure(sc.cursors, B, stream); ensure(sc.info, MSM_INFO_WORDS, stream);
    PlanPtrs pp{};
    pp.offsets = sc.offsets.get(); pp.cursors = sc.cursors.get(); pp.info = sc.info.get();
    if (msm_sorted()) { ensure(sc.perm, B, stream); ensure(sc.invperm, B, stream); pp.perm = sc.perm.get(); pp.invperm = sc.invperm.get(); }
    for (int l = 0; l < MSM_MAX_LEVELS; l++) { ensure(sc.plan[l], B + 1, stream); pp.plan[l] = sc.plan[l].get(); }
    const uint32_t aff_min = msm_affine_min<F>();
    const bool may_halve = total >= aff_min;
    for (int r = 0; r < MSM_MAX_HALVINGS; r++) {
        if (may_halve) ensure(sc.hplan[r], B + 1, stream);
        pp.hplan[r] = sc.hplan[r].get();
    }
    SB_CUDA(cudaMemsetAsync(sc.counts.get(), 0, B * sizeof(uint32_t), stream));
    SB_LAUNCH(k_msm_digits, grid_for(m, 256, 8), 256, 0, stream, job.scalars, m, bases.lay, sc.codes.get(), sc.counts.get());
    SB_LAUNCH(k_scan_plan, 1, 1024, 0, stream, sc.counts.get(), B, pp, msm_s0(total), msm_s1(), aff_min);
    SB_CUDA(cudaMemcpyAsync(job.info_host, sc.info.get(), MSM_INFO_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SB_LAUNCH(k_msm_scatter, grid_for(total, 256, 8), 256, 0, stream, sc.codes.get(), total, sc.cursors.get(), sc.sorted.get());
}

// Pairwise affine rounds (large G2 MSMs), accumulation levels (chunks of S: the first level adds affine points
// into XYZZ partial sums, later levels add partial sums), then sum_k k B_k.  Everything was planned on the
// device in msm_begin; the host only needs the item counts to size the grids.
template <class F>
void msm_finish(MsmJob<F>& job) {
    const MsmBases<F>& bases = *job.bases;
    MsmScratch<F>& sc = bases.scratch;
    cudaStream_t stream = job.stream;
    const uint32_t B = 1u << (bases.lay.c - 1);
    static const uint32_t unbalanced = msm_env_u32("SB_MSM_BALANCED", 1, 0, 1) ? 0u : 0x80000000u;
    const uint32_t S0 = job.info_host[2] | unbalanced, levels = job.info_host[3], S1 = job.info_host[4] | unbalanced, R = job.info_host[5];
    const uint32_t* items = job.info_host + MSM_INFO_ITEMS;
    const uint32_t* htot = job.info_host + MSM_INFO_HTOT;
    if (g_sb_prof_on) { g_sb_prof_tag = 0; while (((size_t)2 << g_sb_prof_tag) <= job.m) g_sb_prof_tag++; }
    SB_REQUIRE(levels >= 1 && levels <= (uint32_t)MSM_MAX_LEVELS && (S0 & 0x7fffffffu) >= 2 && (S1 & 0x7fffffffu) >= 2 && R <= (uint32_t)MSM_MAX_HALVINGS, "msm: bad plan");
    // pairwise rounds
    const AffinePt<F>* aff = nullptr;
    const uint32_t* aff_off = nullptr;
    if (R) {
        ensure(sc.affA, std::max<uint32_t>(htot[0], 1), stream);
        if (R > 1) ensure(sc.affB, std::max<uint32_t>(htot[1], 1), stream);
        static const bool coop = msm_env_u32("SB_MSM_AFFINE_COOP", 0, 0, 1) != 0;
        static const uint32_t coop_k = msm_env_u32("SB_MSM_AFFINE_K", 64, 1, 4096);
        const uint32_t K = coop ? coop_k : (uint32_t)AFF_K;
        const uint32_t nth0 = (htot[0] + K - 1) / K;
        ensure(sc.prefix, (size_t)std::max<uint32_t>(nth0, 1) * K, stream);
        const uint32_t* in_off = sc.offsets.get();
        for (uint32_t r = 0; r < R; r++) {
            AffinePt<F>* outp = (r % 2 == 0) ? sc.affA.get() : sc.affB.get();
            const uint32_t nth = (htot[r] + K - 1) / K;
            if (coop) {
                const int cgrid = (int)((std::max<uint32_t>(nth, 1) + AFFC_THREADS - 1) / AFFC_THREADS);
                const size_t csmem = AFFC_THREADS * sizeof(F);
                if (r == 0)
                    SB_LAUNCH_NAMED(SB_KNAME(F, "k_affine_round_coop"), (k_affine_round_coop<F, true>), cgrid, AFFC_THREADS, csmem, stream,
                                    bases.tab.get(), sc.sorted.get(), aff, in_off, sc.hplan[r].get(), B, nth, K, sc.prefix.get(), outp);
                else
                    SB_LAUNCH_NAMED(SB_KNAME(F, "k_affine_round_coop"), (k_affine_round_coop<F, false>), cgrid, AFFC_THREADS, csmem, stream,
                                    bases.tab.get(), sc.sorted.get(), aff, in_off, sc.hplan[r].get(), B, nth, K, sc.prefix.get(), outp);
                aff = outp; in_off = sc.hplan[r].get();
                continue;
            }
            const int grid = (int)((std::max<uint32_t>(nth, 1) + AFF_THREADS - 1) / AFF_THREADS);
            if (r == 0)
                SB_LAUNCH_NAMED(job.top ? SB_KNAME(F, "k_affine_round:top") : SB_KNAME(F, "k_affine_round"), (k_affine_round<F, true>), grid, AFF_THREADS, 0, stream,
                                bases.tab.get(), sc.sorted.get(), aff, in_off, sc.hplan[r].get(), B, nth, sc.prefix.get(), outp);
            else
                SB_LAUNCH_NAMED(job.top ? SB_KNAME(F, "k_affine_round:top") : SB_KNAME(F, "k_affine_round"), (k_affine_round<F, false>), grid, AFF_THREADS, 0, stream,
                                bases.tab.get(), sc.sorted.get(), aff, in_off, sc.hplan[r].get(), B, nth, sc.prefix.get(), outp);
            aff = outp; in_off = sc.hplan[r].get();
        }
        aff_off = in_off;
    }
    ensure(sc.ptsA, std::max<size_t>(items[0], 1), stream);
    if (levels > 1) ensure(sc.ptsB, std::max<size_t>(items[1], 1), stream);
    const uint32_t* seg = R ? aff_off : sc.offsets.get();
    const XyzzPt<F>* in_pts = nullptr;
    const XyzzPt<F>* last_pts = nullptr;
    for (uint32_t l = 0; l < levels; l++) {
        XyzzPt<F>* outp = (l % 2 == 0) ? sc.ptsA.get() : sc.ptsB.get();
        const uint32_t* plan = sc.plan[l].get();
        const int grid = (int)((std::max<uint32_t>(items[l], 1) + ACC_THREADS - 1) / ACC_THREADS);
        if (l == 0 && !R)
            SB_LAUNCH_NAMED(job.top ? SB_KNAME(F, "k_seg_accum_mixed:top") : SB_KNAME(F, "k_seg_accum_mixed"), (k_seg_accum<F, 1>), grid, ACC_THREADS, 0, stream,
                            bases.tab.get(), sc.sorted.get(), aff, in_pts, seg, plan, B, S0, msm_sorted() ? sc.perm.get() : nullptr, outp);
        else if (l == 0)
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_seg_accum_affine"), (k_seg_accum<F, 2>), grid, ACC_THREADS, 0, stream,
                            bases.tab.get(), sc.sorted.get(), aff, in_pts, seg, plan, B, S0, msm_sorted() ? sc.perm.get() : nullptr, outp);
        else
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_seg_accum_full"), (k_seg_accum<F, 0>), grid, ACC_THREADS, 0, stream,
                            bases.tab.get(), sc.sorted.get(), aff, in_pts, seg, plan, B, S1, nullptr, outp);
        last_pts = outp; seg = plan; in_pts = outp;
        if (l == 0 && job.tail_stream && job.tail_event) {
            SB_CUDA(cudaEventRecord(job.tail_event, stream));
            SB_CUDA(cudaStreamWaitEvent(job.tail_stream, job.tail_event, 0));
            stream = job.tail_stream;
        }
    }
    const uint32_t* last_plan = sc.plan[levels - 1].get();
    // buckets per thread in the first reduction stage: each thread pays 2 L additions for its running sums and
    // ~1.5 log2(B) for the multiplication by its offset; small L = short chain, large L = less total work
    // (measured: L = 2 wins below 2^18 points, L = 4 above; the old L = 8 loses everywhere)
    static const uint32_t red_env = msm_env_u32("SB_MSM_RED_L", 0, 0, 64);
    const uint32_t red_l = red_env ? red_env : (job.m >= ((size_t)1 << 18) ? 4u : 2u);
    const uint32_t L = B >= red_l * RED_THREADS ? red_l : 1;
    const uint32_t nthreads = (B + L - 1) / L;
    const uint32_t nblocks = (nthreads + RED_THREADS - 1) / RED_THREADS;
    ensure(sc.block_out, nblocks, stream);
    const size_t smem = RED_THREADS * sizeof(XyzzPt<F>);
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_bucket_reduce1"), (k_bucket_reduce1<F>), (int)nblocks, RED_THREADS, smem, stream, last_pts, last_plan,
                    msm_sorted() ? sc.invperm.get() : nullptr, B, L, sc.block_out.get());
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_bucket_reduce2"), (k_bucket_reduce2<F>), 1, RED_THREADS, smem, stream, sc.block_out.get(), nblocks, job.out);
    if (stream != job.stream) {
        SB_CUDA(cudaEventRecord(job.tail_event, stream));
        SB_CUDA(cudaStreamWaitEvent(job.stream, job.tail_event, 0));
    }
    g_sb_prof_tag = -1;
}

template <class F>
void msm_run(const MsmBases<F>& bases, const Fr* scalars_dev, size_t m, XyzzPt<F>* out_dev, cudaStream_t stream) {
    static thread_local PinnedBuf<uint32_t> info(MSM_INFO_WORDS);
    MsmJob<F> job;
    job.bases = &bases; job.scalars = scalars_dev; job.m = m; job.out = out_dev; job.stream = stream; job.info_host = info.get();
    msm_begin(job);
    SB_CUDA(cudaStreamSynchronize(stream));
    msm_finish(job);
}

// ------------------------------------------------------------------ fixed-base multiplication (keygen)
constexpr int FB_W = 8;                       // window bits
constexpr int FB_NWIN = 32;                   // 32 * 8 = 256 >= 255

template <class F>
__global__ void k_fb_window_bases(AffinePt<F> g, XyzzPt<F>* __restrict__ wb) {
    int win = threadIdx.x;
    if (win >= FB_NWIN) return;
    XyzzPt<F> p = XyzzPt<F>::from_affine(g);
    for (int k = 0; k < win * FB_W; k++) p = XyzzPt<F>::dbl(p);
    wb[win] = p;
}
template <class F>
__global__ void __launch_bounds__(128) k_fb_table(const XyzzPt<F>* __restrict__ wb, XyzzPt<F>* __restrict__ tab) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= FB_NWIN * (1u << FB_W)) return;
    uint32_t win = t >> FB_W, k = t & ((1u << FB_W) - 1);
    tab[t] = mul_small(wb[win], k);
}
template <class F>
__global__ void __launch_bounds__(128) k_fixed_base(const AffinePt<F>* __restrict__ tab, const Fr* __restrict__ scalars, size_t n,
                                                    XyzzPt<F>* __restrict__ out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = ldg_elem(&scalars[i]).to_canonical();
    XyzzPt<F> acc = XyzzPt<F>::inf();
#pragma unroll
    for (int limb = 0; limb < 8; limb++) {
        uint32_t v = s.l[limb];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t d = (v >> (8 * k)) & 0xffu;
            if (d) acc = XyzzPt<F>::add_mixed(acc, ldg_elem(&tab[(uint32_t)(limb * 4 + k) * 256u + d]));
        }
    }
    st_elem(&out[i], acc);
}

template <class F>
void fixed_base_mul(const AffinePt<F>& g_host, const Fr* scalars_dev, size_t n, AffinePt<F>* out_dev, cudaStream_t stream) {
    const size_t tab_n = (size_t)FB_NWIN << FB_W;
    DevBuf<XyzzPt<F>> wb(FB_NWIN, stream), tabx(tab_n, stream);
    DevBuf<AffinePt<F>> tab(tab_n, stream);
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_fb_window_bases"), (k_fb_window_bases<F>), 1, 32, 0, stream, g_host, wb.get());
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_fb_table"), (k_fb_table<F>), (int)((tab_n + 127) / 128), 128, 0, stream, wb.get(), tabx.get());
    batch_to_affine<F>(tabx.get(), tab.get(), tab_n, stream);
    const size_t chunk = (size_t)1 << 22;
    DevBuf<XyzzPt<F>> tmp(n < chunk ? n : chunk, stream);
    for (size_t i0 = 0; i0 < n; i0 += chunk) {
        size_t cnt = n - i0 < chunk ? n - i0 : chunk;
        SB_LAUNCH_NAMED(SB_KNAME(F, "k_fixed_base"), (k_fixed_base<F>), (int)((cnt + 127) / 128), 128, 0, stream, tab.get(), scalars_dev + i0, cnt, tmp.get());
        batch_to_affine<F>(tmp.get(), out_dev + i0, cnt, stream);
    }
}

template void msm_prepare<Fq>(const AffinePt<Fq>*, size_t, MsmBases<Fq>&, cudaStream_t);
template void msm_prepare<Fq2>(const AffinePt<Fq2>*, size_t, MsmBases<Fq2>&, cudaStream_t);
template void msm_begin<Fq>(MsmJob<Fq>&);
template void msm_begin<Fq2>(MsmJob<Fq2>&);
template void msm_finish<Fq>(MsmJob<Fq>&);
template void msm_finish<Fq2>(MsmJob<Fq2>&);
template void msm_run<Fq>(const MsmBases<Fq>&, const Fr*, size_t, XyzzPt<Fq>*, cudaStream_t);
template void msm_run<Fq2>(const MsmBases<Fq2>&, const Fr*, size_t, XyzzPt<Fq2>*, cudaStream_t);
template void fixed_base_mul<Fq>(const AffinePt<Fq>&, const Fr*, size_t, AffinePt<Fq>*, cudaStream_t);
template void fixed_base_mul<Fq2>(const AffinePt<Fq2>&, const Fr*, size_t, AffinePt<Fq2>*, cudaStream_t);
template void batch_to_affine<Fq>(const XyzzPt<Fq>*, AffinePt<Fq>*, size_t, cudaStream_t);
template void batch_to_affine<Fq2>(const XyzzPt<Fq2>*, AffinePt<Fq2>*, size_t, cudaStream_t);
