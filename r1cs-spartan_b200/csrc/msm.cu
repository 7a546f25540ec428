// msm.cu -- bucket-method MSM over pre-shifted fixed bases (G1 over Fq, G2 over Fq2), fixed-base
// scalar multiplication and batched XYZZ -> affine conversion.  See msm.cuh for the design.
// Integer-pipe bound (Fq Montgomery products); bases are gathered with 128-bit loads.
#include "msm.cuh"

// ------------------------------------------------------------------ window geometry
WinLayout msm_layout(size_t m) {
    int lg = 0;
    while (((size_t)1 << lg) < m) lg++;
    int c = lg - 3;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
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

// exclusive scan of `total` counters by one CTA; offsets[total] = grand total; cursors = copy of offsets
__global__ void __launch_bounds__(1024) k_scan_exclusive(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                                         uint32_t* __restrict__ cursors, uint32_t total, uint32_t* __restrict__ info) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t sh_max;
    if (threadIdx.x == 0) sh_max = 0;
    const uint32_t tid = threadIdx.x;
    const uint32_t chunk = (total + 1023) / 1024;
    const uint32_t beg = tid * chunk, end = min(beg + chunk, total);
    uint32_t sum = 0, mx = 0;
    for (uint32_t i = beg; i < end; i++) { uint32_t cnt = counts[i]; sum += cnt; mx = max(mx, cnt); }
    sh[tid] = sum;
    __syncthreads();
    atomicMax(&sh_max, mx);
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t v = tid >= off ? sh[tid - off] : 0;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    uint32_t run = sh[tid] - sum;
    for (uint32_t i = beg; i < end; i++) { offsets[i] = run; cursors[i] = run; run += counts[i]; }
    if (tid == 1023) { offsets[total] = sh[1023]; info[0] = sh[1023]; info[1] = sh_max; }
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

// chunk_start = exclusive scan of ceil(len_b / S) over the runs seg_off[b] .. seg_off[b+1]; one CTA
__global__ void __launch_bounds__(1024) k_chunk_plan(const uint32_t* __restrict__ seg_off, uint32_t nseg, uint32_t S,
                                                     uint32_t* __restrict__ chunk_start) {
    __shared__ uint32_t sh[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (nseg + 1023) / 1024;
    const uint32_t beg = tid * per, end = min(beg + per, nseg);
    uint32_t sum = 0;
    for (uint32_t i = beg; i < end; i++) sum += (seg_off[i + 1] - seg_off[i] + S - 1) / S;
    sh[tid] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t v = tid >= off ? sh[tid - off] : 0;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    uint32_t run = sh[tid] - sum;
    for (uint32_t i = beg; i < end; i++) { chunk_start[i] = run; run += (seg_off[i + 1] - seg_off[i] + S - 1) / S; }
    if (tid == 1023) chunk_start[nseg] = sh[1023];
}

// one thread per chunk: out[p] = sum of the chunk's elements
template <class F, bool MIXED>
__global__ void __launch_bounds__(128) k_seg_accum(const AffinePt<F>* __restrict__ tab, const uint32_t* __restrict__ sorted,
                                                   const XyzzPt<F>* __restrict__ in_pts, const uint32_t* __restrict__ seg_off,
                                                   const uint32_t* __restrict__ chunk_start, uint32_t nseg, uint32_t S,
                                                   XyzzPt<F>* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= chunk_start[nseg]) return;
    uint32_t lo = 0, hi = nseg;                 // last b with chunk_start[b] <= p (skips empty runs)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&chunk_start[mid]) <= p) lo = mid; else hi = mid;
    }
    uint32_t beg = seg_off[lo] + (p - chunk_start[lo]) * S;
    uint32_t end = min(beg + S, seg_off[lo + 1]);
    XyzzPt<F> acc = XyzzPt<F>::inf();
    for (uint32_t e = beg; e < end; e++) {
        if (MIXED) {
            uint32_t ent = __ldg(&sorted[e]);
            AffinePt<F> q = ldg_elem(&tab[ent & 0x7fffffffu]);
            if (ent & 0x80000000u) q.y = F::neg(q.y);
            acc = XyzzPt<F>::add_mixed(acc, q);
        } else {
            acc = XyzzPt<F>::add(acc, ldg_elem(&in_pts[e]));
        }
    }
    st_elem(&out[p], acc);
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
                                                                uint32_t B, uint32_t L, XyzzPt<F>* __restrict__ block_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XyzzPt<F>* sh = reinterpret_cast<XyzzPt<F>*>(smem_raw);
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)t * L;
    XyzzPt<F> total = XyzzPt<F>::inf();
    if (lo < B) {
        uint32_t hi = (uint32_t)min((uint64_t)B, lo + L);
        XyzzPt<F> run = XyzzPt<F>::inf(), sum = XyzzPt<F>::inf();
        for (uint32_t b = hi; b-- > (uint32_t)lo;) {
            uint32_t o = __ldg(&off[b]);
            if (__ldg(&off[b + 1]) > o) run = XyzzPt<F>::add(run, ldg_elem(&pts[o]));
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

template <class F>
void msm_begin(MsmJob<F>& job) {
    const MsmBases<F>& bases = *job.bases;
    cudaStream_t stream = job.stream;
    const size_t m = job.m;
    SB_REQUIRE(m == bases.m, "msm: scalar count does not match the prepared bases");
    const uint32_t B = 1u << (bases.lay.c - 1);
    const size_t total = (size_t)bases.lay.W * m;             // upper bound on the number of entries
    SB_REQUIRE(total < ((size_t)1 << 31), "msm: too many (window, point) pairs for 31-bit table indices");
    job.codes.alloc(total, stream); job.sorted.alloc(total, stream);
    job.counts.alloc(B, stream); job.offsets.alloc(B + 1, stream); job.cursors.alloc(B, stream); job.info.alloc(2, stream);
    SB_CUDA(cudaMemsetAsync(job.counts.get(), 0, job.counts.bytes(), stream));
    SB_LAUNCH(k_msm_digits, grid_for(m, 256, 8), 256, 0, stream, job.scalars, m, bases.lay, job.codes.get(), job.counts.get());
    SB_LAUNCH(k_scan_exclusive, 1, 1024, 0, stream, job.counts.get(), job.offsets.get(), job.cursors.get(), B, job.info.get());
    SB_CUDA(cudaMemcpyAsync(job.info_host, job.info.get(), 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SB_LAUNCH(k_msm_scatter, grid_for(total, 256, 8), 256, 0, stream, job.codes.get(), total, job.cursors.get(), job.sorted.get());
}

// Chunk size S and level count from the measured longest run: two levels (chunks of S, then at most S
// partial sums per bucket) while the longest run is <= 2^12, more levels beyond (adversarial scalars).
template <class F>
void msm_finish(MsmJob<F>& job) {
    const MsmBases<F>& bases = *job.bases;
    cudaStream_t stream = job.stream;
    const uint32_t B = 1u << (bases.lay.c - 1);
    const size_t entries = job.info_host[0];
    const size_t maxrun = job.info_host[1] ? job.info_host[1] : 1;
    int levels = maxrun <= 1 ? 1 : maxrun <= (1u << 12) ? 2 : maxrun <= (1u << 18) ? 3 : 4;
    uint32_t S = 1;
    while (true) {                                             // smallest S with S^levels >= maxrun
        size_t pw = 1;
        for (int i = 0; i < levels; i++) pw *= S;
        if (pw >= maxrun) break;
        S++;
    }
    if (S < 2) S = 2;
    const size_t bound1 = entries / S + B + 1;
    const size_t bound2 = bound1 / S + B + 1;
    job.ptsA.alloc(bound1, stream); job.ptsB.alloc(bound2, stream);
    job.planA.alloc(B + 1, stream); job.planB.alloc(B + 1, stream);
    const uint32_t* seg = job.offsets.get();
    const XyzzPt<F>* in_pts = nullptr;
    size_t elems = entries, run = maxrun;
    int level = 0;
    const XyzzPt<F>* last_pts = nullptr; const uint32_t* last_plan = nullptr;
    while (true) {
        uint32_t* plan = (level % 2 == 0) ? job.planA.get() : job.planB.get();
        XyzzPt<F>* outp = (level % 2 == 0) ? job.ptsA.get() : job.ptsB.get();
        size_t items = elems / S + B + 1;
        SB_LAUNCH(k_chunk_plan, 1, 1024, 0, stream, seg, B, S, plan);
        if (level == 0)
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_seg_accum_mixed"), (k_seg_accum<F, true>), (int)((items + 127) / 128), 128, 0, stream,
                            bases.tab.get(), job.sorted.get(), in_pts, seg, plan, B, S, outp);
        else
            SB_LAUNCH_NAMED(SB_KNAME(F, "k_seg_accum_full"), (k_seg_accum<F, false>), (int)((items + 127) / 128), 128, 0, stream,
                            bases.tab.get(), job.sorted.get(), in_pts, seg, plan, B, S, outp);
        last_pts = outp; last_plan = plan;
        run = (run + S - 1) / S;
        if (run <= 1) break;
        seg = plan; in_pts = outp; elems = items; level++;
    }
    const uint32_t L = B >= 8 * RED_THREADS ? 8 : 1;
    const uint32_t nthreads = (B + L - 1) / L;
    const uint32_t nblocks = (nthreads + RED_THREADS - 1) / RED_THREADS;
    job.block_out.alloc(nblocks, stream);
    const size_t smem = RED_THREADS * sizeof(XyzzPt<F>);
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_bucket_reduce1"), (k_bucket_reduce1<F>), (int)nblocks, RED_THREADS, smem, stream, last_pts, last_plan, B, L, job.block_out.get());
    SB_LAUNCH_NAMED(SB_KNAME(F, "k_bucket_reduce2"), (k_bucket_reduce2<F>), 1, RED_THREADS, smem, stream, job.block_out.get(), nblocks, job.out);
}

template <class F>
void msm_run(const MsmBases<F>& bases, const Fr* scalars_dev, size_t m, XyzzPt<F>* out_dev, cudaStream_t stream) {
    static thread_local PinnedBuf<uint32_t> info(2);
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
