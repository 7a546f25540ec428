// msm.cu -- bucket-method MSM over pre-shifted fixed bases (G1 over Fq, G2 over Fq2), fixed-base
// scalar multiplication and batched XYZZ -> affine conversion.  See msm.cuh for the design.
// Integer-pipe bound (Fq Montgomery products); bases are gathered with 128-bit loads.
#include "msm.cuh"

// ------------------------------------------------------------------ window geometry
WinLayout msm_layout(size_t m, size_t group_max_m) {
    int lg = 0;
    while (((size_t)1 << lg) < m) lg++;
    // window bits: log2(m) - 3 for the slots of a LARGE group (throughput-bound: m W additions against the per-bucket
    // work of the plan, the last accumulation level and the reduction), log2(m) for the slots of a group whose largest
    // slot has at most 2^17 points (latency-bound: fewer entries per bucket = fewer dependent accumulation levels).
    // Measured in round 2 (sweeps 7-9): at 2^20 constraints offsets -3 / -2 / -1 / 0 give 41.4 / 41.4 / 41.9 / 43.6 ms and a
    // cap of 17-19 bits changes nothing or loses; at 2^17 (the per-GPU share of an 8-GPU proof) -3 / -2 / -1 / 0 / +1 give
    // 14.1 / 13.4 / 12.9 / 12.7 / 13.9 ms.  SB_MSM_C_OFFSET / SB_MSM_C_MAX override the rule for experiments; the layout is
    // fixed when the bases are prepared, so the knobs only act at load time.
    static const int c_off_env = getenv("SB_MSM_C_OFFSET") ? atoi(getenv("SB_MSM_C_OFFSET")) : 99;
    static const int c_max = getenv("SB_MSM_C_MAX") ? atoi(getenv("SB_MSM_C_MAX")) : 16;
    const int c_off = c_off_env != 99 ? c_off_env : (std::max(m, group_max_m) <= ((size_t)1 << 17) ? 0 : -3);
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

// ------------------------------------------------------------------ digits + counting sort (all slots at once)
constexpr uint32_t CODE_NONE = 0xffffffffu;

// One thread per scalar of the group.  Signed-digit recoding of the canonical scalar ("into_repr", commit.rs:20-21,
// open.rs:46) in the window layout of the scalar's slot; code = global bucket index | sign << 31, stored at the entry's
// own table index f = ebase + w m + i (so that the scatter needs no slot lookup); bucket histogram by atomics.
__global__ void __launch_bounds__(256) k_msm_digits(MsmScalarPtrs sp, const MsmSlot* __restrict__ slots, uint32_t nslots, uint32_t mtot,
                                                    uint32_t* __restrict__ codes, uint32_t* __restrict__ counts) {
    __shared__ uint32_t sh_mbase[MSM_MAX_SLOTS + 1];
    __shared__ const Fr* sh_scal[MSM_MAX_SLOTS];
    if (threadIdx.x < nslots) { sh_mbase[threadIdx.x] = slots[threadIdx.x].mbase; sh_scal[threadIdx.x] = sp.p[threadIdx.x]; }
    if (threadIdx.x == nslots) sh_mbase[nslots] = mtot;
    __syncthreads();
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < mtot; g += (size_t)gridDim.x * blockDim.x) {
        uint32_t j = 0;
        while ((uint32_t)g >= sh_mbase[j + 1]) j++;
        const MsmSlot* sl = slots + j;
        const uint32_t i = (uint32_t)g - sh_mbase[j], m = __ldg(&sl->m), ebase = __ldg(&sl->ebase), bbase = __ldg(&sl->bbase);
        const int W = __ldg(&sl->lay.W);
        Fr s = ldg_elem(&sh_scal[j][i]).to_canonical();
        uint32_t limb[9];
#pragma unroll
        for (int k = 0; k < 8; k++) limb[k] = s.l[k];
        limb[8] = 0;
        uint32_t carry = 0;
        int bit = 0;
        for (int w = 0; w < W; w++) {
            const int next = __ldg(&sl->lay.shift[w + 1]), width = next - bit;
            const int lo = bit >> 5, sh = bit & 31;
            uint64_t v = (uint64_t)limb[lo] | ((uint64_t)limb[lo + 1] << 32);
            uint32_t d = ((uint32_t)(v >> sh) & ((1u << width) - 1)) + carry;
            uint32_t code;
            if (w + 1 < W && d > (1u << (width - 1))) {       // negative digit d - 2^width, borrow from the next window
                uint32_t nd = (1u << width) - d; carry = 1;
                code = nd ? ((bbase + nd - 1) | 0x80000000u) : CODE_NONE;
            } else {                                           // the top window is never recoded (see WinLayout)
                carry = 0; code = d ? (bbase + d - 1) : CODE_NONE;
            }
            codes[(size_t)ebase + (size_t)w * m + i] = code;
            if (code != CODE_NONE) atomicAdd(&counts[code & 0x7fffffffu], 1u);
            bit = next;
        }
    }
}

// Chunking of the bucket runs.  Level 0 (mixed additions of table entries, the throughput-bound part) works on
// chunks of at most S0 entries; every further level sums chunks of at most S1 partial sums of the level below, until
// each bucket is down to one point: levels = 1 + ceil(log_S1(ceil(maxrun / S0))).  S0 trades the number of partial
// sums (entries / S0 full additions, 1.4x the cost of a mixed one, and as many 384-byte points written) against
// the length of the dependent chain per thread; S1 is small because the later levels are pure latency: a run of
// r partial sums costs S1 * log_S1(r) dependent additions (measured at 2^17 constraints in round 1: S1 = 3 beats 4, 8
// and two/three equal levels).  A pipeline launches a fixed number `nlaunch` of levels without asking the device how
// many it needs, so the plan kernel RAISES S1 until nlaunch levels cover the longest run (adversarial scalars: all
// equal -> one run of m entries): the result is exact for any input, only the chain per thread grows.
__host__ __device__ inline uint32_t msm_levels(uint32_t maxrun, uint32_t s0, uint32_t s1) {
    uint32_t levels = 1;
    uint64_t cover = s0;
    while (cover < maxrun) { cover *= s1; levels++; }
    return levels;
}
__host__ __device__ inline uint32_t msm_pick_s1(uint32_t maxrun, uint32_t s0, uint32_t s1_min, uint32_t nlaunch) {
    uint32_t s1 = s1_min < 2 ? 2 : s1_min;
    if (nlaunch <= 1) return s1;                   // (a single level can only cover maxrun <= s0; the host never asks for that)
    while (msm_levels(maxrun, s0, s1) > nlaunch) s1++;
    return s1;
}

struct PlanPtrs {
    uint32_t* offsets; uint32_t* cursors; uint32_t* info;
    uint32_t* perm; uint32_t* invperm;          // bucket order of the accumulation (nullptr: natural order)
    uint32_t* plan[MSM_MAX_LEVELS];
    uint32_t* hplan[MSM_MAX_HALVINGS];          // hplan[r-1][b] = exclusive scan (bucket order) of ceil(count_b / 2^r): list layout after round r
    uint32_t R;                                 // pairwise affine rounds before the XYZZ accumulation (0: none)
};
// per-CTA aggregates the plan kernels hand to one another
struct PlanWs {
    uint32_t* cta_hsum;    // [R][nC]         outputs of pairwise round r of the CTA's buckets
    uint32_t* cta_sum;     // [nC]            entries of the CTA's buckets
    uint32_t* cta_max;     // [nC]            longest run among them
    uint32_t* cta_hist;    // [nC][s0 + 1]    buckets per first-level chunk length
    uint32_t* cta_lsum;    // [levels][nC]    chunks per level of the CTA's positions
};
constexpr int PLAN_T = 256, PLAN_V = 4, PLAN_TILE = PLAN_T * PLAN_V;   // a CTA plans 1024 consecutive buckets / positions

// first-level chunk length of a run of c entries: ceil(c / ceil(c / s0)) <= s0 (0 for an empty bucket)
__host__ __device__ inline uint32_t plan_key(uint32_t c, uint32_t s0) {
    const uint32_t nch = (c + s0 - 1) / s0;
    return nch ? (c + nch - 1) / nch : 0;
}
// sum (and max) of one value per thread over a CTA of PLAN_T threads; every thread gets the results
__device__ inline void plan_block_sum3(uint32_t& a, uint32_t& b, uint32_t& mx, uint32_t* sh /* 3 * PLAN_T */) {
    const uint32_t t = threadIdx.x;
    sh[t] = a; sh[PLAN_T + t] = b; sh[2 * PLAN_T + t] = mx;
    __syncthreads();
    for (uint32_t off = PLAN_T / 2; off > 0; off >>= 1) {
        if (t < off) {
            sh[t] += sh[t + off]; sh[PLAN_T + t] += sh[PLAN_T + t + off];
            sh[2 * PLAN_T + t] = max(sh[2 * PLAN_T + t], sh[2 * PLAN_T + t + off]);
        }
        __syncthreads();
    }
    a = sh[0]; b = sh[PLAN_T]; mx = sh[2 * PLAN_T];
    __syncthreads();
}
// exclusive scan of one value per thread over the CTA
__device__ inline uint32_t plan_block_exscan(uint32_t v, uint32_t* sh /* PLAN_T */) {
    const uint32_t t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (uint32_t off = 1; off < (uint32_t)PLAN_T; off <<= 1) {
        uint32_t x = t >= off ? sh[t - off] : 0;
        __syncthreads();
        sh[t] += x;
        __syncthreads();
    }
    const uint32_t r = sh[t] - v;
    __syncthreads();
    return r;
}

// The plan of a pipeline, computed on the device by four small multi-CTA kernels (the single-CTA version of round 1
// walked 2^16 counters a dozen times with dependent loads: 0.45 ms per MSM; these take a few microseconds each):
//   A  per CTA: entries, longest run, histogram of first-level chunk lengths
//   C  offsets / cursors (exclusive scan of the counters), the run statistics -> S1, levels (info), and the accumulation
//      order of the buckets: by decreasing chunk length of the first level (counting sort on the histograms) --
//      thread p of the accumulation kernel takes chunk p in this order, so the 32 lanes of a warp run loops of
//      (almost) equal length and the shortest chunks are the ones left when the grid drains
//   D  per CTA of positions in that order: chunks per level
//   F  plan[l][k] = exclusive scan over that order of ceil(count / (S0 S1^l))   (chunk plan of accumulation level l)
// info: [0] entries, [1] longest run, [2] S0, [3] levels, [4] S1, [MSM_INFO_ITEMS + l] chunks of level l.
// (with R pairwise affine rounds in front, the XYZZ accumulation sees runs of ceil(count / 2^R) points: chunk lengths,
// accumulation order and level plans are all derived from those)
__global__ void __launch_bounds__(PLAN_T) k_plan_a(const uint32_t* __restrict__ counts, uint32_t B, uint32_t s0, uint32_t R, PlanWs ws) {
    __shared__ uint32_t hist[1025];
    __shared__ uint32_t sh[3 * PLAN_T];
    const uint32_t t = threadIdx.x, nC = gridDim.x;
    for (uint32_t k = t; k <= s0; k += PLAN_T) hist[k] = 0;
    __syncthreads();
    uint32_t sum = 0, zero = 0, mx = 0;
    uint32_t hs[MSM_MAX_HALVINGS];
#pragma unroll
    for (int r = 0; r < MSM_MAX_HALVINGS; r++) hs[r] = 0;
    const uint32_t b0 = blockIdx.x * PLAN_TILE + t * PLAN_V;
    const uint32_t radd = (1u << R) - 1;
#pragma unroll
    for (int v = 0; v < PLAN_V; v++) {
        const uint32_t b = b0 + v;
        if (b < B) {
            const uint32_t c = __ldg(&counts[b]);
            sum += c; mx = max(mx, c);
            atomicAdd(&hist[plan_key((c + radd) >> R, s0)], 1u);
#pragma unroll
            for (int r = 0; r < MSM_MAX_HALVINGS; r++) if ((uint32_t)r < R) hs[r] += (c + (2u << r) - 1) >> (r + 1);
        }
    }
    plan_block_sum3(sum, zero, mx, sh);
    if (t == 0) { ws.cta_sum[blockIdx.x] = sum; ws.cta_max[blockIdx.x] = mx; }
    for (uint32_t r = 0; r < R; r++) {
        uint32_t h = hs[r], z0 = 0, z1 = 0;
        plan_block_sum3(h, z0, z1, sh);
        if (t == 0) ws.cta_hsum[(size_t)r * nC + blockIdx.x] = h;
    }
    for (uint32_t k = t; k <= s0; k += PLAN_T) ws.cta_hist[(size_t)blockIdx.x * (s0 + 1) + k] = hist[k];
}

__global__ void __launch_bounds__(PLAN_T) k_plan_c(const uint32_t* __restrict__ counts, uint32_t B, uint32_t s0, uint32_t s1_min, uint32_t nlaunch,
                                                   PlanWs ws, PlanPtrs pp) {
    __shared__ uint32_t pos[1025];       // per chunk length: next position in the accumulation order for this CTA's buckets
    __shared__ uint32_t tot[1025];
    __shared__ uint32_t sh[3 * PLAN_T];
    const uint32_t t = threadIdx.x, nC = gridDim.x, blk = blockIdx.x;
    // entries before this CTA, entries in total, longest run (every CTA derives them from the per-CTA aggregates)
    uint32_t before = 0, total = 0, mx = 0;
    for (uint32_t c = t; c < nC; c += PLAN_T) {
        const uint32_t s = ws.cta_sum[c];
        total += s; if (c < blk) before += s;
        mx = max(mx, ws.cta_max[c]);
    }
    plan_block_sum3(before, total, mx, sh);
    const uint32_t R = pp.R, radd = (1u << R) - 1;
    if (blk == 0 && t == 0) {
        const uint32_t mxr = (mx + radd) >> R;            // longest run the XYZZ accumulation sees
        const uint32_t s1 = msm_pick_s1(mxr, s0, s1_min, nlaunch);
        pp.offsets[B] = total;
        pp.info[0] = total; pp.info[1] = mx; pp.info[2] = s0; pp.info[3] = msm_levels(mxr, s0, s1); pp.info[4] = s1;
    }
    const bool sorted = pp.perm != nullptr;
    if (sorted) {
        // buckets of chunk length k: tot[k] in the whole group, pos[k] of them in the CTAs before this one
        for (uint32_t k = t; k <= s0; k += PLAN_T) {
            uint32_t all = 0, below = 0;
            for (uint32_t c = 0; c < nC; c++) {
                const uint32_t h = ws.cta_hist[(size_t)c * (s0 + 1) + k];
                all += h; if (c < blk) below += h;
            }
            tot[k] = all; pos[k] = below;
        }
        __syncthreads();
        if (t == 0) {        // longest chunks first
            uint32_t run = 0;
            for (int k = (int)s0; k >= 0; k--) { pos[k] += run; run += tot[k]; }
        }
        __syncthreads();
    }
    uint32_t c[PLAN_V], sum = 0;
    const uint32_t b0 = blk * PLAN_TILE + t * PLAN_V;
#pragma unroll
    for (int v = 0; v < PLAN_V; v++) { c[v] = (b0 + v < B) ? __ldg(&counts[b0 + v]) : 0; sum += c[v]; }
    uint32_t run = before + plan_block_exscan(sum, sh);
#pragma unroll
    for (int v = 0; v < PLAN_V; v++) {
        const uint32_t b = b0 + v;
        if (b < B) {
            pp.offsets[b] = run; pp.cursors[b] = run; run += c[v];
            if (sorted) {
                const uint32_t p = atomicAdd(&pos[plan_key((c[v] + radd) >> R, s0)], 1u);
                pp.perm[p] = b; pp.invperm[b] = p;
            }
        }
    }
    // layouts of the lists after every pairwise round
    for (uint32_t r = 0; r < R; r++) {
        uint32_t bef = 0, tot2 = 0, z = 0;
        for (uint32_t cc = t; cc < nC; cc += PLAN_T) {
            const uint32_t h = ws.cta_hsum[(size_t)r * nC + cc];
            tot2 += h; if (cc < blk) bef += h;
        }
        plan_block_sum3(bef, tot2, z, sh);
        uint32_t n[PLAN_V], hsum = 0;
#pragma unroll
        for (int v = 0; v < PLAN_V; v++) { n[v] = (c[v] + (2u << r) - 1) >> (r + 1); hsum += n[v]; }
        uint32_t x = bef + plan_block_exscan(hsum, sh);
#pragma unroll
        for (int v = 0; v < PLAN_V; v++) {
            if (b0 + v < B) { pp.hplan[r][b0 + v] = x; x += n[v]; }
        }
        if (blk == 0 && t == 0) { pp.hplan[r][B] = tot2; pp.info[MSM_INFO_HTOT + r] = tot2; }
    }
}

// D (WRITE = false): chunks per level of this CTA's positions -> ws.cta_lsum;  F (WRITE = true): the chunk plans
template <bool WRITE>
__global__ void __launch_bounds__(PLAN_T) k_plan_levels(const uint32_t* __restrict__ counts, uint32_t B, PlanWs ws, PlanPtrs pp) {
    __shared__ uint32_t sh[3 * PLAN_T];
    const uint32_t t = threadIdx.x, nC = gridDim.x, blk = blockIdx.x;
    const uint32_t s0 = pp.info[2], levels = pp.info[3], s1 = pp.info[4];
    uint32_t c[PLAN_V];
    const uint32_t k0 = blk * PLAN_TILE + t * PLAN_V;
#pragma unroll
    for (int v = 0; v < PLAN_V; v++) {
        const uint32_t k = k0 + v;
        c[v] = k < B ? ((__ldg(&counts[pp.perm ? pp.perm[k] : k]) + ((1u << pp.R) - 1)) >> pp.R) : 0;
    }
    uint64_t div = s0;
    for (uint32_t l = 0; l < levels; l++, div *= s1) {
        uint32_t n[PLAN_V], sum = 0;
#pragma unroll
        for (int v = 0; v < PLAN_V; v++) { n[v] = (uint32_t)((c[v] + div - 1) / div); sum += n[v]; }
        if (!WRITE) {
            uint32_t z0 = 0, z1 = 0;
            plan_block_sum3(sum, z0, z1, sh);
            if (t == 0) ws.cta_lsum[(size_t)l * nC + blk] = sum;
        } else {
            uint32_t before = 0, total = 0, z = 0;
            for (uint32_t cc = t; cc < nC; cc += PLAN_T) {
                const uint32_t s = ws.cta_lsum[(size_t)l * nC + cc];
                total += s; if (cc < blk) before += s;
            }
            plan_block_sum3(before, total, z, sh);
            uint32_t run = before + plan_block_exscan(sum, sh);
#pragma unroll
            for (int v = 0; v < PLAN_V; v++) {
                if (k0 + v < B) { pp.plan[l][k0 + v] = run; run += n[v]; }
            }
            if (blk == 0 && t == 0) { pp.plan[l][B] = total; pp.info[MSM_INFO_ITEMS + l] = total; }
        }
    }
}

__global__ void __launch_bounds__(256) k_msm_scatter(const uint32_t* __restrict__ codes, size_t total, uint32_t* __restrict__ cursors,
                                                     uint32_t* __restrict__ sorted) {
    for (size_t f = blockIdx.x * (size_t)blockDim.x + threadIdx.x; f < total; f += (size_t)gridDim.x * blockDim.x) {
        uint32_t code = codes[f];
        if (code == CODE_NONE) continue;
        uint32_t pos = atomicAdd(&cursors[code & 0x7fffffffu], 1u);
        sorted[pos] = (uint32_t)f | (code & 0x80000000u);      // f indexes the group's pre-shifted table
    }
}

// ------------------------------------------------------------------ bucket accumulation
// Bucket b's entries (all windows: the bases are pre-shifted) are one contiguous run of `sorted`.
// Runs can be extremely uneven (adversarial scalars), so the work item is not a bucket but a CHUNK of one run.
// Level 0 turns chunks of table entries into partial sums (mixed additions); every further level sums chunks of
// the previous level's partial sums, until each bucket is down to one point.  Work per thread is bounded at every
// level whatever the scalars are.  One thread per chunk: out[p] = sum of the chunk's elements.
constexpr int ACC_THREADS = 64;
// MIXED: level 0, mixed additions of table entries (sorted -> tab, sign in bit 31); otherwise level `level` >= 1, full
// additions of the partial sums of level - 1 (in_pts, laid out by the plan of level - 1 = seg_off).  Levels the plan
// does not use return at once (the host launches a fixed number of levels, see msm_group_run).
// (after pairwise affine rounds, level 0 reads the affine list they left -- aff_in, laid out by the last round's plan --
// instead of the table)
template <class F, bool MIXED>
__global__ void __launch_bounds__(ACC_THREADS) k_seg_accum(const AffinePt<F>* __restrict__ tab, const uint32_t* __restrict__ sorted,
                                                           const AffinePt<F>* __restrict__ aff_in,
                                                           const XyzzPt<F>* __restrict__ in_pts, const uint32_t* __restrict__ seg_off,
                                                           const uint32_t* __restrict__ chunk_start, uint32_t nseg, uint32_t level,
                                                           const uint32_t* __restrict__ info, const uint32_t* __restrict__ perm,
                                                           XyzzPt<F>* __restrict__ out) {
    if (!MIXED && level >= __ldg(&info[3])) return;
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= chunk_start[nseg]) return;
    uint32_t lo = 0, hi = nseg;                 // last b with chunk_start[b] <= p (skips empty runs)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&chunk_start[mid]) <= p) lo = mid; else hi = mid;
    }
    // balanced split of the run into its planned number of chunks: the threads of a warp get chunks of nearly
    // equal length instead of S, S, ..., remainder
    const uint32_t j = p - chunk_start[lo], nch = chunk_start[lo + 1] - chunk_start[lo];
    const uint32_t sb = (MIXED && perm) ? __ldg(&perm[lo]) : lo;        // position in the accumulation order -> bucket (level 0 only)
    const uint32_t off = seg_off[sb], cnt = seg_off[sb + 1] - off;
    const uint32_t beg = off + (uint32_t)(((uint64_t)j * cnt) / nch);
    const uint32_t end = off + (uint32_t)(((uint64_t)(j + 1) * cnt) / nch);
    XyzzPt<F> acc = XyzzPt<F>::inf();
    for (uint32_t e = beg; e < end; e++) {
        if (MIXED) {
            AffinePt<F> q;
            if (aff_in) q = ldg_elem(&aff_in[e]);
            else {
                uint32_t ent = __ldg(&sorted[e]);
                q = ldg_elem(&tab[ent & 0x7fffffffu]);
                if (ent & 0x80000000u) q.y = F::neg(q.y);
            }
            acc = XyzzPt<F>::add_mixed(acc, q);
        } else {
            acc = XyzzPt<F>::add(acc, ldg_elem(&in_pts[e]));
        }
    }
    st_elem(&out[p], acc);
}

// ------------------------------------------------------------------ pairwise rounds in affine coordinates
// One round halves every bucket run: output element j of bucket b is the sum of input elements 2j and 2j+1 of that
// bucket (or a copy of element 2j when the run is odd).  An affine addition is one inversion plus 2M + 1S; each thread
// owns K consecutive output elements and shares ONE inversion among them (Montgomery's trick: +3M per element), so an
// addition costs 5M + 1S + (one inversion) / K instead of the 8M + 2S of a mixed XYZZ addition -- over Fq2, 4 176
// against 6 912 integer products.  The inversion is the binary-GCD one (Fp::inv_fast, ~30 k instructions): with
// the Fermat ladder (~170 k) round 1 measured this LOSING at every K; the rounds only pay with the cheap inversion.
// The exceptional cases of the group law (an input at infinity, P = Q, P = -Q) are handled exactly; their
// denominator is replaced by 1 so that the shared inversion stays well defined.
constexpr int AFF_THREADS = 64;
enum { PAIR_COPY_P = 0, PAIR_COPY_Q = 1, PAIR_ADD = 2, PAIR_DBL = 3, PAIR_INF = 4 };

// denominator of the addition P + Q from the x coordinates alone; the y coordinates are only fetched in the rare cases
// that need them (an x coordinate that is zero: the point may be the encoding of infinity; equal x coordinates:
// doubling or cancellation).  load_y(k): y of input k (0 = P, 1 = Q).
template <class F, class LoadY>
__device__ __forceinline__ int pair_classify_x(const F& px, const F& qx, bool has2, LoadY load_y, F& d) {
    d = F::one();
    if (!has2) return PAIR_COPY_P;
    const bool pz = px.is_zero(), qz = qx.is_zero();
    if (qz && load_y(1).is_zero()) return PAIR_COPY_P;          // Q at infinity
    if (pz && load_y(0).is_zero()) return PAIR_COPY_Q;          // P at infinity
    if (px == qx) {
        const F py = load_y(0);
        if (py == load_y(1) && !py.is_zero()) { d = F::dbl(py); return PAIR_DBL; }
        return PAIR_INF;
    }
    d = F::sub(qx, px);
    return PAIR_ADD;
}

// The outputs of a round are dealt out warp by warp: a warp owns 32 * per consecutive outputs (per = ceil(total /
// nthreads) <= K) and lane l takes outputs base + l, base + l + 32, ...: in every iteration the 32 lanes work on 32
// CONSECUTIVE outputs -- the same bucket or two neighbouring ones, so the bucket tracking is (almost) uniform across the
// warp, the entry indices and the results are read and written as contiguous runs, and every lane runs the same number
// of iterations (a first version with per-thread contiguous shares and bucket-by-bucket loops had 17.9 of 32 lanes
// active: ncu, profiles/r02_ncu_affine_round_v1.txt).
template <class F, bool FIRST>
__global__ void __launch_bounds__(AFF_THREADS) k_affine_round(const AffinePt<F>* __restrict__ tab, const uint32_t* __restrict__ sorted,
                                                              const AffinePt<F>* __restrict__ in_aff, const uint32_t* __restrict__ in_off,
                                                              const uint32_t* __restrict__ out_off, uint32_t B, uint32_t nthreads,
                                                              F* __restrict__ prefix, AffinePt<F>* __restrict__ out_aff) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = out_off[B];
    const uint32_t per = (total + nthreads - 1) / nthreads;
    if (t >= nthreads || per == 0) return;
    const uint64_t first64 = (uint64_t)(t & ~31u) * per + (t & 31u);      // this lane's first output
    if (first64 >= total) return;
    const uint32_t first = (uint32_t)first64;
    // iterations of this lane: outputs first + 32 i < min(warp's end, total)
    const uint64_t wend = min((uint64_t)((t & ~31u) + 32) * per, (uint64_t)total);
    const uint32_t iters = (uint32_t)((wend - first + 31) / 32);
    // element `idx` of the input list: its address (first round: through the sorted entry, with the sign of the digit)
    auto in_ptr = [&](uint32_t idx, bool& negate) -> const AffinePt<F>* {
        if (FIRST) {
            const uint32_t ent = __ldg(&sorted[idx]);
            negate = (ent & 0x80000000u) != 0;
            return &tab[ent & 0x7fffffffu];
        }
        negate = false;
        return &in_aff[idx];
    };
    uint32_t lo = 0, hi = B;                    // bucket of the first output: last b with out_off[b] <= first
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&out_off[mid]) <= first) lo = mid; else hi = mid;
    }
    // the two inputs of output p (the bucket cursor b moves with p: forwards in the first pass, backwards in the second)
    struct Pair { const AffinePt<F>* p0; const AffinePt<F>* p1; bool has2, n0, n1; };
    uint32_t b = lo;
    auto locate = [&](uint32_t p) -> Pair {
        while (__ldg(&out_off[b + 1]) <= p) b++;
        while (__ldg(&out_off[b]) > p) b--;
        const uint32_t j = p - __ldg(&out_off[b]);
        const uint32_t ib = __ldg(&in_off[b]), n_in = __ldg(&in_off[b + 1]) - ib;
        Pair pr;
        pr.has2 = 2 * j + 1 < n_in; pr.n1 = false;
        pr.p0 = in_ptr(ib + 2 * j, pr.n0);
        pr.p1 = pr.has2 ? in_ptr(ib + 2 * j + 1, pr.n1) : pr.p0;
        return pr;
    };
    // (Four attempts on the stalls around these loads were measured in round 2 and removed -- profiles/r02_ncu_affine_round_v3..v6:
    //   prefetch instructions one iteration ahead (sweep 8): 41.6 against 41.0 ms per proof;
    //   cp.async staging of the next iteration's points and prefix product in shared memory (sweep 11): 39.4 against 39.2 ms, the
    //     kernel itself 4.47 against 3.80 ms -- the carve-out took L1 away from local memory (hit rate 86 % -> 72 %);
    //   a register-lean second pass (re-reading coordinates through laundered pointers: 148 registers, 12 warps per SM instead
    //     of 8; sweep 12): 39.1 against 39.4 ms, the kernel 3.91 ms -- more warps, same traffic, more instruction-cache misses;
    //   L1-bypassing loads and streaming stores for the once-read data (sweep 13): 40.4 against 39.4 ms, 4.53 ms, although the
    //     local loads then hit L1 90 % of the time.
    // What ncu counts as local traffic here is the OPERANDS of the outlined Fq2 products -- 14 STL.128 by the caller and 48 LD.E
    // by the callee per call, 35 M + 11 M warp-instructions per launch -- not spills of this loop.)
    // pass 1: prefix products of the denominators (x coordinates only)
    F acc = F::one();
    Pair cur = locate(first), nxt = cur;
    for (uint32_t i = 0; i < iters; i++) {
        if (i + 1 < iters) nxt = locate(first + 32 * (i + 1));
        const AffinePt<F>* pp0 = cur.p0; const AffinePt<F>* pp1 = cur.p1;
        const bool has2 = cur.has2, n0 = cur.n0, n1 = cur.n1;
        const F px = ldg_elem(&pp0->x), qx = has2 ? ldg_elem(&pp1->x) : px;
        F d;
        pair_classify_x(px, qx, has2, [&](int k) { F y = ldg_elem(k ? &pp1->y : &pp0->y); return (k ? n1 : n0) ? F::neg(y) : y; }, d);
        st_elem(&prefix[(size_t)i * nthreads + t], acc);
        acc = F::mul(acc, d);
        cur = nxt;
    }
    F inv = F::inv_fast(acc);
    // pass 2, backwards: 1/d_k = inv * prefix_k, then inv *= d_k   (cur is the last output's pair: the first pass left it there)
    for (uint32_t i = iters; i-- > 0;) {
        if (i > 0) nxt = locate(first + 32 * (i - 1));
        const uint32_t p = first + 32 * i;
        const bool has2 = cur.has2;
        AffinePt<F> P = ldg_elem(cur.p0), Q = has2 ? ldg_elem(cur.p1) : P;
        if (cur.n0) P.y = F::neg(P.y);
        if (has2 && cur.n1) Q.y = F::neg(Q.y);
        if (!has2) Q = P;
        F d;
        const int kind = pair_classify_x(P.x, Q.x, has2, [&](int k) { return k ? Q.y : P.y; }, d);
        const F dinv = F::mul(inv, ldg_elem(&prefix[(size_t)i * nthreads + t]));
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
        cur = nxt;
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

// ------------------------------------------------------------------ quad-cooperative point arithmetic (latency-bound stages)
// The bucket reduction is a chain of DEPENDENT point operations whatever the MSM size.  On one thread a G2 addition is 14
// Fq2 products one after another (~18 k instructions, 25-30 us on a lone warp), and round 1 measured the two reduction
// kernels at 1.2 + 0.65 ms per pipeline -- the latency floor of every commitment and opening, and what held the 8-GPU
// efficiency at 0.42.  Here FOUR adjacent lanes (a "quad") share every operation: each stage of the XYZZ formulas has up
// to four independent products, each lane takes one -- 4 product latencies per addition instead of 14, 3 per doubling
// instead of 9.
// The operands live in SHARED memory: every quad owns a few point slots and a scratch area of field elements; a lane
// picks POINTERS to the two operands of its product (the outlined field product reads its operands from memory
// anyway), stores the result in the quad's scratch, and the quad meets at a __syncwarp.  (The first version kept the
// points in registers and exchanged results by shuffles through address-taken temporaries: ncu,
// profiles/r02_ncu_bucket_reduce1_quad.txt, showed 4.4 KB of stack per thread, local loads hitting L1 only 56 % of the
// time, 715 MB of DRAM writes from a kernel whose results are a few kilobytes, ~10 cycles per instruction and warp.)
// Every lane of a warp must call these together (__syncwarp is warp-wide); the exceptional cases of the group law are
// resolved AFTER the common path, so the calls never diverge.
constexpr int QUAD_WS = 14;          // field elements of scratch per quad
constexpr int QUAD_SLOTS = 4;        // point slots per quad
template <class F>
SB_D const F* quad_ptr(int r, const F* a0, const F* a1, const F* a2, const F* a3) { return r == 0 ? a0 : r == 1 ? a1 : r == 2 ? a2 : a3; }
template <class F>
SB_D F* quad_comp(XyzzPt<F>* p, int r) { return r == 0 ? &p->X : r == 1 ? &p->Y : r == 2 ? &p->ZZ : &p->ZZZ; }
template <class F>
SB_D const F* quad_comp(const XyzzPt<F>* p, int r) { return r == 0 ? &p->X : r == 1 ? &p->Y : r == 2 ? &p->ZZ : &p->ZZZ; }
// dst <- src, one coordinate per lane (dst != src)
template <class F>
SB_D void quad_copy(XyzzPt<F>* dst, const XyzzPt<F>* src, int r) { *quad_comp(dst, r) = *quad_comp(src, r); }
template <class F>
SB_D void quad_set_inf(XyzzPt<F>* dst, int r) { *quad_comp(dst, r) = F::zero(); }

// dst <- p + q (add-2008-s); dst may be p or q.  s: the quad's scratch (QUAD_WS field elements)
template <class F>
__device__ __noinline__ void quad_add(XyzzPt<F>* dst, const XyzzPt<F>* p, const XyzzPt<F>* q, F* s) {
    const int r = threadIdx.x & 3;
    const bool p_inf = p->ZZ.is_zero(), q_inf = q->ZZ.is_zero();
    s[r] = F::mul(*quad_ptr(r, &p->X, &q->X, &p->Y, &q->Y), *quad_ptr(r, &q->ZZ, &p->ZZ, &q->ZZZ, &p->ZZZ));   // U1, U2, S1, S2
    __syncwarp();
    const F Pp = F::sub(s[1], s[0]), R = F::sub(s[3], s[2]);
    const bool p_zero = Pp.is_zero(), r_zero = R.is_zero();
    if (r == 0) s[4] = Pp;
    if (r == 2) s[5] = R;
    __syncwarp();
    s[6 + r] = F::mul(*quad_ptr(r, &s[4], &p->ZZ, &s[5], &p->ZZZ), *quad_ptr(r, &s[4], &q->ZZ, &s[5], &q->ZZZ));   // PP, ZZ1 ZZ2, R^2, ZZZ1 ZZZ2
    __syncwarp();
    const F m3 = F::mul(*quad_ptr(r, &s[4], &s[7], &s[0], &s[7]), s[6]);                                       // PPP, ZZ3, Q, (unused)
    if (r != 3) s[10 + r] = m3;
    __syncwarp();
    const F X3 = F::sub(F::sub(s[8], s[10]), F::dbl(s[12]));
    if (r == 2) s[13] = F::sub(s[12], X3);
    __syncwarp();
    const F m4 = F::mul(*quad_ptr(r, &s[2], &s[2], &s[5], &s[9]), *quad_ptr(r, &s[10], &s[10], &s[13], &s[10]));   // S1 PPP, (unused), R (Q - X3), ZZZ3
    if (r == 0) s[4] = m4;
    __syncwarp();
    if (p_inf | q_inf | p_zero) {                         // the same in all four lanes
        if (p_inf) { if (dst != q) quad_copy(dst, q, r); }
        else if (q_inf) { if (dst != p) quad_copy(dst, p, r); }
        else if (r_zero) { if (r == 0) { const XyzzPt<F> d = XyzzPt<F>::dbl(*p); *dst = d; } }
        else quad_set_inf(dst, r);
    } else {
        if (r == 0) dst->X = X3;
        else if (r == 1) dst->ZZ = m3;
        else if (r == 2) dst->Y = F::sub(m4, s[4]);
        else dst->ZZZ = m4;
    }
    __syncwarp();
}
// dst <- 2 p (dbl-2008-s-1); dst may be p
template <class F>
__device__ __noinline__ void quad_dbl(XyzzPt<F>* dst, const XyzzPt<F>* p, F* s) {
    const int r = threadIdx.x & 3;
    const bool p_inf = p->ZZ.is_zero();
    if (r == 0) s[0] = F::dbl(p->Y);                                                                           // U
    __syncwarp();
    const F m1 = F::mul(*quad_ptr(r, &s[0], &p->X, &s[0], &s[0]), *quad_ptr(r, &s[0], &p->X, &s[0], &s[0]));  // V = U^2, XX, (unused x 2)
    if (r == 0) s[1] = m1;                                                                                     // V
    if (r == 1) s[2] = F::add(F::dbl(m1), m1);                                                                 // M = 3 XX
    __syncwarp();
    s[3 + r] = F::mul(*quad_ptr(r, &s[0], &p->X, &s[2], &p->ZZ), *quad_ptr(r, &s[1], &s[1], &s[2], &s[1]));    // W, S, M^2, ZZ3
    __syncwarp();
    const F X3 = F::sub(s[5], F::dbl(s[4]));
    if (r == 0) s[7] = F::sub(s[4], X3);
    __syncwarp();
    const F m3 = F::mul(*quad_ptr(r, &s[2], &s[3], &s[3], &s[3]), *quad_ptr(r, &s[7], &p->Y, &p->ZZZ, &p->Y)); // M (S - X3), W Y, ZZZ3, (unused)
    if (r == 1) s[8] = m3;
    __syncwarp();
    if (!p_inf) {
        if (r == 0) { dst->Y = F::sub(m3, s[8]); dst->X = X3; }
        else if (r == 2) dst->ZZZ = m3;
        else if (r == 3) dst->ZZ = s[6];
    } else if (dst != p) quad_copy(dst, p, r);
    __syncwarp();
}
// acc <- k acc for a small k that is the same in every lane of the warp; base: another point slot of the quad (clobbered)
template <class F>
__device__ void quad_mul_small(XyzzPt<F>* acc, XyzzPt<F>* base, uint32_t k, F* s) {
    const int r = threadIdx.x & 3;
    if (k == 0) { quad_set_inf(acc, r); __syncwarp(); return; }
    quad_copy(base, acc, r);
    __syncwarp();
    for (int bit = 30 - __clz(k); bit >= 0; bit--) {
        quad_dbl(acc, acc, s);
        if ((k >> bit) & 1) quad_add(acc, acc, base, s);
    }
}

// ------------------------------------------------------------------ bucket reduction: sum_b (b + 1) B_b per slot
// Two stages of quads.  Stage 1: a CTA of nq quads owns G = nq * L consecutive buckets of one slot, quad t the buckets
// [t L, (t + 1) L) of them.  With i = t L + j the local index of a bucket,
//     sum_i (i + 1) B_i = sum_t [ sum_j (j + 1) B_{tL+j} ] + L * sum_t t * run_t,     run_t = sum_j B_{tL+j},
// the inner sums are running sums (2 L additions per quad), and sum_t t * run_t = sum_{t >= 1} (run_t + run_{t+1} + ...) is a
// suffix scan over the quads (log2 nq steps) followed by one tree.  The CTA writes A_c (the local sum above) and R_c = the
// sum of its buckets.  Stage 2 (one CTA per slot) does the same to the CTAs: total = sum_c A_c + G sum_c c R_c.
// Depth: 2 L + 2 log2(nq) + 3 quad operations in stage 1 and about 30 in stage 2, whatever the MSM size (round 1: 2 L + ~22
// doublings/additions of a scalar multiple + 6 tree levels on single threads).
constexpr int RED_QUADS = 64, RED_THREADS = 4 * RED_QUADS;      // largest CTA; groups of tiny slots launch fewer quads (MsmGroup::red_quads, a power of two >= 8)
// shared memory of a reduction CTA: QUAD_SLOTS point slots and QUAD_WS scratch elements per quad
template <class F>
static size_t red_smem_bytes(uint32_t nq) { return (size_t)nq * (QUAD_SLOTS * sizeof(XyzzPt<F>) + QUAD_WS * sizeof(F)); }
template <class F>
struct QuadMem {
    XyzzPt<F>* slot;      // this quad's QUAD_SLOTS points
    F* ws;                // this quad's scratch
    XyzzPt<F>* all;       // slot 0 of quad 0 (quad t's slots start at all + t * QUAD_SLOTS)
    __device__ QuadMem(unsigned char* raw) {
        const uint32_t nq = blockDim.x >> 2, qd = threadIdx.x >> 2;
        all = reinterpret_cast<XyzzPt<F>*>(raw);
        slot = all + (size_t)qd * QUAD_SLOTS;
        ws = reinterpret_cast<F*>(all + (size_t)nq * QUAD_SLOTS) + (size_t)qd * QUAD_WS;
    }
};
// suffix scan over the quads of a CTA of slot `k`: slot k of quad t <- sum_{t' >= t} (slot k of quad t'); slot `tmp` is clobbered
template <class F>
__device__ void quad_block_suffix_scan(QuadMem<F>& m, int k, int tmp) {
    const uint32_t qd = threadIdx.x >> 2, r = threadIdx.x & 3, nq = blockDim.x >> 2;
    __syncthreads();
    for (uint32_t off = 1; off < nq; off <<= 1) {
        F c = F::zero();
        if (qd + off < nq) c = *quad_comp(m.all + (size_t)(qd + off) * QUAD_SLOTS + k, r);
        __syncthreads();                                   // every quad has read its neighbour before anybody updates
        *quad_comp(m.slot + tmp, r) = c;
        __syncwarp();
        quad_add(m.slot + k, m.slot + k, m.slot + tmp, m.ws);
        __syncthreads();
    }
}
// tree sum over the quads of a CTA of slot `k`; the result is left in slot k of quad 0; slot `tmp` is clobbered
template <class F>
__device__ void quad_block_tree_sum(QuadMem<F>& m, int k, int tmp) {
    const uint32_t qd = threadIdx.x >> 2, r = threadIdx.x & 3;
    __syncthreads();
    for (uint32_t stride = blockDim.x >> 3; stride > 0; stride >>= 1) {
        const bool live = (qd & ~7u) < stride;             // whole warps (8 quads) drop out once they hold no active quad; idle quads of a live warp add infinity
        F c = F::zero();
        if (live && qd < stride) c = *quad_comp(m.all + (size_t)(qd + stride) * QUAD_SLOTS + k, r);
        __syncthreads();
        if (live) {
            *quad_comp(m.slot + tmp, r) = c;
            __syncwarp();
            quad_add(m.slot + k, m.slot + k, m.slot + tmp, m.ws);
        }
        __syncthreads();
    }
}
// CTA -> slot by the slots' rbase; the final points of the accumulation are found through the device-side plan
// (levels = info[3]: the last level's output buffer and chunk plan).  block_out[2 c] = A_c, block_out[2 c + 1] = R_c.
template <class F>
__global__ void __launch_bounds__(RED_THREADS) k_bucket_reduce1(const XyzzPt<F>* __restrict__ ptsA, const XyzzPt<F>* __restrict__ ptsB, PlanPtrs pp,
                                                                const MsmSlot* __restrict__ slots, uint32_t nslots, XyzzPt<F>* __restrict__ block_out) {
    SB_DYN_SMEM(smem_raw);
    QuadMem<F> m(smem_raw);
    enum { RUN = 0, SUM = 1, TMP = 2, AUX = 3 };
    uint32_t j = 0;
    while (j + 1 < nslots && blockIdx.x >= __ldg(&slots[j + 1].rbase)) j++;
    const MsmSlot* sl = slots + j;
    const uint32_t B = __ldg(&sl->nb), L = __ldg(&sl->red_l), bbase = __ldg(&sl->bbase);
    const uint32_t levels = __ldg(&pp.info[3]);
    const XyzzPt<F>* pts = ((levels - 1) & 1) ? ptsB : ptsA;
    const uint32_t* off = pp.plan[levels - 1];
    const uint32_t qd = threadIdx.x >> 2, r = threadIdx.x & 3;
    const uint64_t lo = ((uint64_t)(blockIdx.x - __ldg(&sl->rbase)) * (blockDim.x >> 2) + qd) * L;
    quad_set_inf(m.slot + RUN, r); quad_set_inf(m.slot + SUM, r);
    for (uint32_t jj = L; jj-- > 0;) {
        const uint64_t b = lo + jj;
        F c = F::zero();
        if (b < B) {
            const uint32_t k = pp.invperm ? __ldg(&pp.invperm[bbase + (uint32_t)b]) : bbase + (uint32_t)b;      // bucket -> position in the accumulation order
            const uint32_t o = __ldg(&off[k]);
            if (__ldg(&off[k + 1]) > o) c = ldg_elem(quad_comp(&pts[o], r));
        }
        *quad_comp(m.slot + TMP, r) = c;
        __syncwarp();
        quad_add(m.slot + RUN, m.slot + RUN, m.slot + TMP, m.ws);
        quad_add(m.slot + SUM, m.slot + SUM, m.slot + RUN, m.ws);
    }
    quad_block_suffix_scan(m, RUN, TMP);                               // RUN <- sum of the runs of the quads >= this one; quad 0: R_c
    if (qd == 0) st_elem(quad_comp(&block_out[2 * (size_t)blockIdx.x + 1], r), *quad_comp(m.slot + RUN, r));
    if (qd == 0) quad_set_inf(m.slot + RUN, r);
    __syncwarp();
    quad_mul_small(m.slot + RUN, m.slot + AUX, L, m.ws);
    quad_add(m.slot + SUM, m.slot + SUM, m.slot + RUN, m.ws);
    quad_block_tree_sum(m, SUM, TMP);
    if (qd == 0) st_elem(quad_comp(&block_out[2 * (size_t)blockIdx.x], r), *quad_comp(m.slot + SUM, r));
}
// stage 2: one CTA per slot; quad q takes the stage-1 CTAs [q cpt, (q + 1) cpt) of the slot
template <class F>
__global__ void __launch_bounds__(RED_THREADS) k_bucket_reduce2(const XyzzPt<F>* __restrict__ block_out, const MsmSlot* __restrict__ slots,
                                                                XyzzPt<F>* __restrict__ out) {
    SB_DYN_SMEM(smem_raw);
    QuadMem<F> m(smem_raw);
    enum { RUN = 0, SUM = 1, TMP = 2, ACC = 3 };       // sum_j R_{q cpt + j}, sum_j j R_{q cpt + j}, scratch, sum_c A_c
    const MsmSlot* sl = slots + blockIdx.x;
    const uint32_t r0 = __ldg(&sl->rbase), nblocks = __ldg(&sl->rblocks), L = __ldg(&sl->red_l);
    const uint32_t nq = blockDim.x >> 2, cpt = (nblocks + nq - 1) / nq, qd = threadIdx.x >> 2, r = threadIdx.x & 3;
    quad_set_inf(m.slot + RUN, r); quad_set_inf(m.slot + SUM, r); quad_set_inf(m.slot + ACC, r);
    for (uint32_t jj = cpt; jj-- > 0;) {
        const uint32_t c = qd * cpt + jj;
        F a = F::zero(), rr = F::zero();
        if (c < nblocks) { a = ldg_elem(quad_comp(&block_out[2 * (size_t)(r0 + c)], r)); rr = ldg_elem(quad_comp(&block_out[2 * (size_t)(r0 + c) + 1], r)); }
        *quad_comp(m.slot + TMP, r) = a;
        __syncwarp();
        quad_add(m.slot + ACC, m.slot + ACC, m.slot + TMP, m.ws);
        *quad_comp(m.slot + TMP, r) = rr;
        __syncwarp();
        quad_add(m.slot + RUN, m.slot + RUN, m.slot + TMP, m.ws);
        if (jj >= 1) quad_add(m.slot + SUM, m.slot + SUM, m.slot + RUN, m.ws);
    }
    quad_block_suffix_scan(m, RUN, TMP);
    // sum_c c R_c = sum_q [ sum_j j R + cpt * q * run_q ]; the total weighs it by G = quads * L buckets per stage-1 CTA
    if (qd == 0) quad_set_inf(m.slot + RUN, r);
    __syncwarp();
    quad_mul_small(m.slot + RUN, m.slot + TMP, cpt, m.ws);
    quad_add(m.slot + SUM, m.slot + SUM, m.slot + RUN, m.ws);
    quad_mul_small(m.slot + SUM, m.slot + TMP, nq * L, m.ws);
    quad_add(m.slot + SUM, m.slot + SUM, m.slot + ACC, m.ws);
    quad_block_tree_sum(m, SUM, TMP);
    if (qd == 0) st_elem(quad_comp(&out[blockIdx.x], r), *quad_comp(m.slot + SUM, r));
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

// ------------------------------------------------------------------ group preparation and the pipeline
static uint32_t msm_env_u32(const char* name, uint32_t dflt, uint32_t lo, uint32_t hi) {
    const char* e = getenv(name);
    long v = e ? atol(e) : (long)dflt;
    return (uint32_t)std::min<long>(std::max<long>(v, lo), hi);
}
// S0 by size: throughput-bound groups take long chunks (fewer partial sums), latency-bound ones short chains
static uint32_t msm_s0(size_t entries) {      // (at most 1024: the plan kernels keep one histogram bin per chunk length in shared memory)
    static const uint32_t v = msm_env_u32("SB_MSM_S0", 0, 0, 1024);
    static const uint32_t big = msm_env_u32("SB_MSM_S0_BIG", 48, 2, 1024), small = msm_env_u32("SB_MSM_S0_SMALL", 24, 2, 1024);
    return v >= 2 ? v : (entries >= ((size_t)1 << 22) ? big : small);
}
// Accumulate the buckets in the order of decreasing first-level chunk length (see k_scan_plan): the lanes of a warp
// then run loops of equal length.  Measured in round 1: 2^20 constraints 64.3 -> 59.8 ms.  SB_MSM_SORTED=0 restores
// the natural bucket order.
static bool msm_sorted() { static const bool v = msm_env_u32("SB_MSM_SORTED", 1, 0, 1) != 0; return v; }
static uint32_t msm_s1_min() { static const uint32_t v = msm_env_u32("SB_MSM_S1", 3, 2, 4096); return v; }
// accumulation levels a pipeline launches; the device raises S1 when that many levels of the minimal S1 would not
// cover the longest run (S0 * 3^4 = 3888 entries per bucket at the defaults, 10x the runs of uniform scalars)
// Pairwise affine rounds (see k_affine_round): R rounds for groups of at least 2^SB_MSM_AFFINE_LOG2 entries, each thread
// sharing one inversion among at most SB_MSM_AFFINE_K additions.  Measured on B200 at 2^20 constraints (round 2, gpurun sweep 5):
// R = 0: 49.3 ms per proof; R = 4, K = 128: 47.0; R = 4, K = 256: 41.8; R = 5: 42.2; R = 4, K = 256 and the G1 commitment
// too: 41.4 -- the defaults.  Below 2^21 entries per group (2^17 constraints) the rounds lose a few percent (too few threads
// left in the later rounds) and stay off.  SB_MSM_AFFINE_ROUNDS=0 disables them; SB_MSM_AFFINE_G1=0 keeps G1 on XYZZ.
template <class F>
static uint32_t msm_affine_rounds(uint64_t etot) {
    static const uint32_t lg = msm_env_u32("SB_MSM_AFFINE_LOG2", 21, 0, 40), rounds = msm_env_u32("SB_MSM_AFFINE_ROUNDS", 4, 0, MSM_MAX_HALVINGS);
    static const bool g1 = msm_env_u32("SB_MSM_AFFINE_G1", 1, 0, 1) != 0;
    if (sizeof(F) != sizeof(Fq2) && !g1) return 0;
    return etot >= ((uint64_t)1 << lg) ? rounds : 0;
}
static uint32_t msm_affine_kmax() { static const uint32_t v = msm_env_u32("SB_MSM_AFFINE_K", 256, 1, 1024); return v; }
// CTAs of the round kernel an SM holds (registers decide: 4 over Fq2, 6-8 over Fq); the rounds are sized to exactly one wave
template <class F>
static uint32_t msm_affine_ctas_per_sm() {
    static const uint32_t forced = msm_env_u32("SB_MSM_AFFINE_CTAS", 0, 0, 16);
    if (forced) return forced;
    static uint32_t cached = 0;
    if (!cached) {
        int a = 0, b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_affine_round<F, true>, AFF_THREADS, 0) != cudaSuccess) a = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_affine_round<F, false>, AFF_THREADS, 0) != cudaSuccess) b = 4;
        cached = (uint32_t)std::max(1, std::min(a, b));
    }
    return cached;
}
static uint32_t msm_nlaunch() { static const uint32_t v = msm_env_u32("SB_MSM_LEVELS", 5, 2, MSM_MAX_LEVELS); return v; }

template <class F>
void msm_group_prepare(const std::vector<const AffinePt<F>*>& bases_dev, const std::vector<size_t>& ms, MsmGroup<F>& out, cudaStream_t stream) {
    const size_t J = ms.size();
    SB_REQUIRE(J >= 1 && J <= (size_t)MSM_MAX_SLOTS && bases_dev.size() == J, "msm: bad slot count");
    out.slots.assign(J, MsmSlot{});
    uint64_t mtot = 0, etot = 0, btot = 0, rtot = 0;
    static const uint32_t red_env = msm_env_u32("SB_MSM_RED_L", 0, 0, 64);
    {   // quads per reduction CTA: 64, fewer when every slot of the group is tiny
        size_t mmax = 0;
        for (size_t j = 0; j < J; j++) mmax = std::max(mmax, ms[j]);
        const uint32_t nbmax = 1u << (msm_layout(mmax, mmax).c - 1);
        static const uint32_t quads_cap = msm_env_u32("SB_MSM_RED_QUADS", 32, 8, RED_QUADS);
        out.red_quads = 8;
        while (out.red_quads * 2 <= quads_cap && out.red_quads * 4 < nbmax) out.red_quads *= 2;
    }
    size_t group_max = 0;
    for (size_t j = 0; j < J; j++) group_max = std::max(group_max, ms[j]);
    for (size_t j = 0; j < J; j++) {
        MsmSlot& s = out.slots[j];
        SB_REQUIRE(ms[j] >= 1 && ms[j] < ((size_t)1 << 31), "msm: slot size out of range");
        s.m = (uint32_t)ms[j]; s.lay = msm_layout(ms[j], group_max);
        s.nb = 1u << (s.lay.c - 1);
        // buckets per quad in the first reduction stage (see k_bucket_reduce1): 2 L chained additions per quad.  Measured in
        // round 2 (sweep 11), quads per CTA x L at 2^20 / 2^17 constraints: 64 x 4: 39.4 / 11.1 ms; 32 x 8: 39.4 / 10.6;
        // 32 x 4: 39.5 / 11.2; 16 x 8: 39.8 / 11.2; 32 x 16: 40.0 / 11.0 -- 32 quads (one warp per scheduler: the quads'
        // products do not queue behind one another on the integer pipe) with 8 buckets each
        const uint32_t red_l = red_env ? red_env : 8u;
        s.red_l = s.nb >= red_l * out.red_quads ? red_l : 1;
        const uint32_t nquads = (s.nb + s.red_l - 1) / s.red_l;
        s.rblocks = (nquads + out.red_quads - 1) / out.red_quads;
        s.mbase = (uint32_t)mtot; s.ebase = (uint32_t)etot; s.bbase = (uint32_t)btot; s.rbase = (uint32_t)rtot;
        mtot += s.m; etot += (uint64_t)s.lay.W * s.m; btot += s.nb; rtot += s.rblocks;
    }
    SB_REQUIRE(etot < ((uint64_t)1 << 31) && mtot < ((uint64_t)1 << 31), "msm: too many (window, point) pairs for 31-bit table indices");
    out.mtot = (uint32_t)mtot; out.etot = (uint32_t)etot; out.btot = (uint32_t)btot; out.rtot = (uint32_t)rtot;
    out.R = msm_affine_rounds<F>(etot);
    out.s0 = msm_s0(etot >> out.R);
    {
        // the XYZZ accumulation sees at most etot / 2^R + btot points
        const uint64_t e0 = (etot >> out.R) + (out.R ? btot : 0);
        uint64_t div = out.s0;
        for (int l = 0; l < MSM_MAX_LEVELS; l++, div *= msm_s1_min()) out.items_bound[l] = (uint32_t)(e0 / div + btot);
        for (uint32_t r = 0; r < out.R; r++) {
            const uint64_t bound = (etot >> (r + 1)) + btot;            // outputs of round r
            const uint64_t resident = (uint64_t)SB_SMS * msm_affine_ctas_per_sm<F>() * AFF_THREADS;  // threads one wave of the round kernel holds
            uint64_t k = (bound + resident - 1) / resident;
            k = std::min<uint64_t>(std::max<uint64_t>(k, std::min<uint32_t>(8, msm_affine_kmax())), msm_affine_kmax());
            out.round_bound[r] = (uint32_t)bound; out.round_k[r] = (uint32_t)k; out.round_threads[r] = (uint32_t)(((bound + k - 1) / k + 31) / 32 * 32);   // whole warps (see k_affine_round)
        }
    }
    out.slots_dev.alloc(J, stream);
    SB_CUDA(cudaMemcpyAsync(out.slots_dev.get(), out.slots.data(), J * sizeof(MsmSlot), cudaMemcpyHostToDevice, stream));
    // tables: window multiples of every base, converted to affine in batches
    out.tab.alloc(etot, stream);
    for (size_t j = 0; j < J; j++) {
        const MsmSlot& s = out.slots[j];
        const size_t m = s.m;
        const int W = s.lay.W;
        const size_t chunk = m < ((size_t)1 << 16) ? m : ((size_t)1 << 16);
        DevBuf<XyzzPt<F>> tmp((size_t)W * chunk, stream);
        for (size_t i0 = 0; i0 < m; i0 += chunk) {
            size_t cnt = m - i0 < chunk ? m - i0 : chunk;
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_preshift"), (k_preshift<F>), (int)((cnt + 127) / 128), 128, 0, stream, bases_dev[j], i0, cnt, s.lay, tmp.get());
            size_t n = (size_t)W * cnt;
            size_t threads = (n + BTA_K - 1) / BTA_K;
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_batch_to_affine"), (k_batch_to_affine<F>), (int)((threads + 127) / 128), 128, 0, stream, tmp.get(),
                            out.tab.get() + s.ebase, n, cnt, m, i0);
        }
    }
    // scratch of the pipeline, sized by the upper bounds (no allocation, no host round trip while proving)
    MsmScratch<F>& sc = out.scratch;
    sc.codes.alloc(etot, stream); sc.sorted.alloc(etot, stream);
    sc.counts.alloc(btot, stream); sc.offsets.alloc(btot + 1, stream); sc.cursors.alloc(btot, stream); sc.info.alloc(MSM_INFO_WORDS, stream);
    sc.perm.alloc(btot, stream); sc.invperm.alloc(btot, stream);
    {
        const size_t nC = (btot + PLAN_TILE - 1) / PLAN_TILE;
        sc.cta_sum.alloc(nC, stream); sc.cta_max.alloc(nC, stream); sc.cta_hist.alloc(nC * (out.s0 + 1), stream);
        sc.cta_lsum.alloc(nC * MSM_MAX_LEVELS, stream);
    }
    for (uint32_t l = 0; l < msm_nlaunch(); l++) sc.plan[l].alloc(btot + 1, stream);
    sc.ptsA.alloc(std::max<uint32_t>(out.items_bound[0], 1), stream);
    sc.ptsB.alloc(std::max<uint32_t>(out.items_bound[1], 1), stream);
    sc.block_out.alloc(2 * (size_t)rtot, stream);
    if (out.R) {
        const size_t nC = (btot + PLAN_TILE - 1) / PLAN_TILE;
        sc.cta_hsum.alloc(nC * out.R, stream);
        size_t pre = 0;
        for (uint32_t r = 0; r < out.R; r++) {
            sc.hplan[r].alloc(btot + 1, stream);
            pre = std::max<size_t>(pre, (size_t)out.round_threads[r] * out.round_k[r]);
        }
        sc.affA.alloc(out.round_bound[0], stream);
        if (out.R > 1) sc.affB.alloc(out.round_bound[1], stream);
        sc.prefix.alloc(pre, stream);
    }
    SB_CUDA(cudaStreamSynchronize(stream));           // `slots` (host) was the source of an asynchronous copy
}

// digits + histogram -> plan -> scatter -> accumulation levels -> bucket reduction, all slots at once, no host round trip.
// The pipeline is queued in three phases so that a caller can pipeline several groups (see msm_groups_run):
//   front  digits, plan, scatter                                    (short, launch-bound)
//   accum  pairwise affine rounds, first accumulation level         (throughput-bound: this is where the time goes)
//   tail   later accumulation levels, bucket reduction              (latency-bound: ~2 ms of dependent point additions)
template <class F>
static PlanPtrs msm_plan_ptrs(const MsmGroup<F>& g) {
    MsmScratch<F>& sc = g.scratch;
    PlanPtrs pp{};
    pp.offsets = sc.offsets.get(); pp.cursors = sc.cursors.get(); pp.info = sc.info.get();
    if (msm_sorted()) { pp.perm = sc.perm.get(); pp.invperm = sc.invperm.get(); }
    for (uint32_t l = 0; l < msm_nlaunch(); l++) pp.plan[l] = sc.plan[l].get();
    pp.R = g.R;
    for (uint32_t r = 0; r < g.R; r++) pp.hplan[r] = sc.hplan[r].get();
    return pp;
}
template <class F>
static void msm_prof_tag(const MsmGroup<F>& g) {
    if (g_sb_prof_on) { g_sb_prof_tag = 0; while (((size_t)2 << g_sb_prof_tag) <= g.mtot) g_sb_prof_tag++; }
}
template <class F>
void msm_group_front(const MsmGroup<F>& g, const MsmScalarPtrs& scalars, cudaStream_t stream) {
    MsmScratch<F>& sc = g.scratch;
    const uint32_t J = (uint32_t)g.nslots(), B = g.btot, nlaunch = msm_nlaunch();
    PlanPtrs pp = msm_plan_ptrs(g);
    msm_prof_tag(g);
    SB_CUDA(cudaMemsetAsync(sc.counts.get(), 0, (size_t)B * sizeof(uint32_t), stream));
    SB_LAUNCH(k_msm_digits, grid_for(g.mtot, 256, 8), 256, 0, stream, scalars, g.slots_dev.get(), J, g.mtot, sc.codes.get(), sc.counts.get());
    PlanWs ws{sc.cta_hsum.get(), sc.cta_sum.get(), sc.cta_max.get(), sc.cta_hist.get(), sc.cta_lsum.get()};
    const int nC = (int)((B + PLAN_TILE - 1) / PLAN_TILE);
    SB_LAUNCH(k_plan_a, nC, PLAN_T, 0, stream, sc.counts.get(), B, g.s0, g.R, ws);
    SB_LAUNCH(k_plan_c, nC, PLAN_T, 0, stream, sc.counts.get(), B, g.s0, msm_s1_min(), nlaunch, ws, pp);
    SB_LAUNCH_NAMED("k_plan_levels<sum>", (k_plan_levels<false>), nC, PLAN_T, 0, stream, sc.counts.get(), B, ws, pp);
    SB_LAUNCH_NAMED("k_plan_levels<write>", (k_plan_levels<true>), nC, PLAN_T, 0, stream, sc.counts.get(), B, ws, pp);
    SB_LAUNCH(k_msm_scatter, grid_for(g.etot, 256, 8), 256, 0, stream, sc.codes.get(), (size_t)g.etot, sc.cursors.get(), sc.sorted.get());
    g_sb_prof_tag = -1;
}
template <class F>
void msm_group_accum(const MsmGroup<F>& g, cudaStream_t stream) {
    MsmScratch<F>& sc = g.scratch;
    const uint32_t B = g.btot;
    PlanPtrs pp = msm_plan_ptrs(g);
    msm_prof_tag(g);
    // pairwise affine rounds: every bucket run is halved R times
    const AffinePt<F>* aff = nullptr;
    const uint32_t* seg0 = sc.offsets.get();
    for (uint32_t r = 0; r < g.R; r++) {
        AffinePt<F>* outp = (r % 2 == 0) ? sc.affA.get() : sc.affB.get();
        const int grid = (int)((g.round_threads[r] + AFF_THREADS - 1) / AFF_THREADS);
        if (r == 0)
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_affine_round"), (k_affine_round<F, true>), grid, AFF_THREADS, 0, stream, g.tab.get(), sc.sorted.get(), aff, seg0,
                            sc.hplan[r].get(), B, g.round_threads[r], sc.prefix.get(), outp);
        else
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_affine_round"), (k_affine_round<F, false>), grid, AFF_THREADS, 0, stream, g.tab.get(), sc.sorted.get(), aff, seg0,
                            sc.hplan[r].get(), B, g.round_threads[r], sc.prefix.get(), outp);
        aff = outp; seg0 = sc.hplan[r].get();
    }
    const int grid = (int)((std::max<uint32_t>(g.items_bound[0], 1) + ACC_THREADS - 1) / ACC_THREADS);
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_seg_accum_mixed"), (k_seg_accum<F, true>), grid, ACC_THREADS, 0, stream, g.tab.get(), sc.sorted.get(), aff,
                    (const XyzzPt<F>*)nullptr, seg0, sc.plan[0].get(), B, 0u, sc.info.get(), pp.perm, sc.ptsA.get());
    g_sb_prof_tag = -1;
}
template <class F>
void msm_group_tail(const MsmGroup<F>& g, XyzzPt<F>* out_dev, cudaStream_t stream) {
    MsmScratch<F>& sc = g.scratch;
    const uint32_t J = (uint32_t)g.nslots(), B = g.btot, nlaunch = msm_nlaunch();
    PlanPtrs pp = msm_plan_ptrs(g);
    msm_prof_tag(g);
    for (uint32_t l = 1; l < nlaunch; l++) {
        XyzzPt<F>* outp = (l % 2 == 0) ? sc.ptsA.get() : sc.ptsB.get();
        const XyzzPt<F>* inp = (l % 2 == 0) ? sc.ptsB.get() : sc.ptsA.get();
        const int grid = (int)((std::max<uint32_t>(g.items_bound[l], 1) + ACC_THREADS - 1) / ACC_THREADS);
        SB_LAUNCH_NAMED(SB_KNAME(F, "k_seg_accum_full"), (k_seg_accum<F, false>), grid, ACC_THREADS, 0, stream, g.tab.get(), sc.sorted.get(),
                        (const AffinePt<F>*)nullptr, inp, sc.plan[l - 1].get(), sc.plan[l].get(), B, l, sc.info.get(), (const uint32_t*)nullptr, outp);
    }
    const size_t smem = red_smem_bytes<F>(g.red_quads);
    const int red_threads = 4 * (int)g.red_quads;
    // more than the default 48 KB of dynamic shared memory: a function attribute, set once per DEVICE (a multi-GPU context
    // drives several devices from one process) and per instantiation
    static std::atomic<uint64_t> attr_set{0};
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (!((attr_set.load(std::memory_order_acquire) >> (dev & 63)) & 1)) {
        const int cap = (int)red_smem_bytes<F>(RED_QUADS);
        SB_CUDA(cudaFuncSetAttribute((const void*)k_bucket_reduce1<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        SB_CUDA(cudaFuncSetAttribute((const void*)k_bucket_reduce2<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        attr_set.fetch_or((uint64_t)1 << (dev & 63), std::memory_order_release);
    }
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_bucket_reduce1"), (k_bucket_reduce1<F>), (int)g.rtot, red_threads, smem, stream, sc.ptsA.get(), sc.ptsB.get(), pp,
                    g.slots_dev.get(), J, sc.block_out.get());
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_bucket_reduce2"), (k_bucket_reduce2<F>), (int)J, red_threads, smem, stream, sc.block_out.get(), g.slots_dev.get(), out_dev);
    g_sb_prof_tag = -1;
}
template <class F>
void msm_group_run(const MsmGroup<F>& g, const MsmScalarPtrs& scalars, XyzzPt<F>* out_dev, cudaStream_t stream, const char* tag) {
    (void)tag;
    msm_group_front(g, scalars, stream);
    msm_group_accum(g, stream);
    msm_group_tail(g, out_dev, stream);
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

template void msm_group_prepare<Fq>(const std::vector<const AffinePt<Fq>*>&, const std::vector<size_t>&, MsmGroup<Fq>&, cudaStream_t);
template void msm_group_prepare<Fq2>(const std::vector<const AffinePt<Fq2>*>&, const std::vector<size_t>&, MsmGroup<Fq2>&, cudaStream_t);
template void msm_group_front<Fq>(const MsmGroup<Fq>&, const MsmScalarPtrs&, cudaStream_t);
template void msm_group_front<Fq2>(const MsmGroup<Fq2>&, const MsmScalarPtrs&, cudaStream_t);
template void msm_group_accum<Fq>(const MsmGroup<Fq>&, cudaStream_t);
template void msm_group_accum<Fq2>(const MsmGroup<Fq2>&, cudaStream_t);
template void msm_group_tail<Fq>(const MsmGroup<Fq>&, XyzzPt<Fq>*, cudaStream_t);
template void msm_group_tail<Fq2>(const MsmGroup<Fq2>&, XyzzPt<Fq2>*, cudaStream_t);
template void msm_group_run<Fq>(const MsmGroup<Fq>&, const MsmScalarPtrs&, XyzzPt<Fq>*, cudaStream_t, const char*);
template void msm_group_run<Fq2>(const MsmGroup<Fq2>&, const MsmScalarPtrs&, XyzzPt<Fq2>*, cudaStream_t, const char*);
template void fixed_base_mul<Fq>(const AffinePt<Fq>&, const Fr*, size_t, AffinePt<Fq>*, cudaStream_t);
template void fixed_base_mul<Fq2>(const AffinePt<Fq2>&, const Fr*, size_t, AffinePt<Fq2>*, cudaStream_t);
template void batch_to_affine<Fq>(const XyzzPt<Fq>*, AffinePt<Fq>*, size_t, cudaStream_t);
template void batch_to_affine<Fq2>(const XyzzPt<Fq2>*, AffinePt<Fq2>*, size_t, cudaStream_t);
