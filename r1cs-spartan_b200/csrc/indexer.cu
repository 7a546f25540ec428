// indexer.cu -- device-side construction of the sparse-product plans (see indexer.cuh).
//
// Everything here is integer / index work over the caller's CSR arrays: HBM- and atomics-bound, a few passes over
// nnz entries (2^20 constraints: 6.8 M entries, 2^24: 109 M).  Field elements are only moved and compared with 1.
// The ORDER of the entries inside a column segment and of the items inside a length class comes from atomic cursors
// and may differ from run to run; the plans' results do not (field addition is exact, so a segment's sum does not
// depend on the order of its terms) -- tests compare Az/Bz/Cz and M(r_x, .) bit for bit against the oracle.
#include "indexer.cuh"

namespace {
constexpr int IDX_T = 256;

__global__ void __launch_bounds__(IDX_T) k_idx_validate_rows(const uint64_t* __restrict__ rp, size_t n, uint64_t cap, uint32_t* err) {
    bool bad = false;
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r <= n; r += (size_t)gridDim.x * blockDim.x) {
        const uint64_t a = rp[r];
        if (r == 0 ? a != 0 : false) bad = true;
        if (a > cap) bad = true;
        if (r < n && a > rp[r + 1]) bad = true;
    }
    if (bad) atomicOr(err, (uint32_t)IDX_ERR_ROWPTR);
}
__global__ void __launch_bounds__(IDX_T) k_idx_validate_cols(const uint32_t* __restrict__ col, size_t nnz, uint32_t ncols, uint32_t* err) {
    bool bad = false;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nnz; e += (size_t)gridDim.x * blockDim.x)
        if (col[e] >= ncols) bad = true;
    if (bad) atomicOr(err, (uint32_t)IDX_ERR_COL);
}
__global__ void __launch_bounds__(IDX_T) k_idx_row_segments(const uint64_t* __restrict__ rp, size_t lo, size_t nl, uint32_t base, bool close,
                                                            uint32_t* __restrict__ seg_ptr) {
    const uint64_t r0 = rp[lo];
    const size_t cnt = nl + (close ? 1 : 0);
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < cnt; r += (size_t)gridDim.x * blockDim.x)
        seg_ptr[r] = base + (uint32_t)(rp[lo + r] - r0);
}
__global__ void __launch_bounds__(IDX_T) k_idx_flags(const uint32_t* __restrict__ col, const Fr* __restrict__ val, uint32_t* __restrict__ idx, size_t cnt) {
    const Fr one = Fr::one();
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < cnt; e += (size_t)gridDim.x * blockDim.x)
        idx[e] = col[e] | (ldg_elem(&val[e]) == one ? SEG_UNIT_FLAG : 0u);
}
__global__ void __launch_bounds__(IDX_T) k_idx_col_count(const uint32_t* __restrict__ col, size_t nnz, uint32_t lo, uint32_t hi, uint32_t* __restrict__ cnt) {
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nnz; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = col[e];
        if (y >= lo && y < hi) atomicAdd(&cnt[y - lo], 1u);
    }
}
__global__ void __launch_bounds__(IDX_T) k_idx_col_scatter(const uint64_t* __restrict__ rp, size_t n_rows, const uint32_t* __restrict__ col,
                                                           const Fr* __restrict__ val, size_t nnz, uint32_t lo, uint32_t hi, uint32_t tag_base,
                                                           uint32_t* __restrict__ cursor, uint32_t* __restrict__ idx_out, Fr* __restrict__ val_out) {
    const Fr one = Fr::one();
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nnz; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = col[e];
        if (y < lo || y >= hi) continue;
        // row of entry e: the last r with row_ptr[r] <= e (rows may be empty, and one row may hold most of the matrix:
        // a thread per ENTRY with a search, not a thread per row)
        size_t a = 0, b = n_rows;
        while (b - a > 1) {
            const size_t mid = (a + b) >> 1;
            if (rp[mid] <= e) a = mid; else b = mid;
        }
        const Fr v = ldg_elem(&val[e]);
        const uint32_t pos = atomicAdd(&cursor[y - lo], 1u);
        idx_out[pos] = (tag_base + (uint32_t)a) | (v == one ? SEG_UNIT_FLAG : 0u);
        st_elem(&val_out[pos], v);
    }
}

// ---- exclusive scan of u32 (three passes: per-tile sums, scan of the tile sums by one CTA, tile-local scans)
constexpr int SCAN_V = 4, SCAN_TILE = IDX_T * SCAN_V;
__device__ inline uint32_t idx_block_exscan(uint32_t v, uint32_t* sh, uint32_t& total) {
    const uint32_t t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (uint32_t off = 1; off < (uint32_t)IDX_T; off <<= 1) {
        const uint32_t x = t >= off ? sh[t - off] : 0;
        __syncthreads();
        sh[t] += x;
        __syncthreads();
    }
    const uint32_t r = sh[t] - v;
    total = sh[IDX_T - 1];
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(IDX_T) k_scan_tile_sums(const uint32_t* __restrict__ data, size_t n, uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t sh[IDX_T];
    const size_t b0 = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_V;
    uint32_t s = 0;
#pragma unroll
    for (int v = 0; v < SCAN_V; v++) if (b0 + v < n) s += data[b0 + v];
    uint32_t total;
    idx_block_exscan(s, sh, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(IDX_T) k_scan_tile_offsets(uint32_t* __restrict__ tile_sum, size_t ntiles) {      // one CTA; tile_sum[ntiles] <- grand total
    __shared__ uint32_t sh[IDX_T];
    uint32_t carry = 0;
    for (size_t base = 0; base < ntiles; base += IDX_T) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < ntiles ? tile_sum[i] : 0;
        uint32_t total;
        const uint32_t ex = idx_block_exscan(v, sh, total);
        if (i < ntiles) tile_sum[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sum[ntiles] = carry;
}
__global__ void __launch_bounds__(IDX_T) k_scan_apply(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ tile_sum, size_t ntiles) {
    __shared__ uint32_t sh[IDX_T];
    const size_t b0 = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_V;
    uint32_t c[SCAN_V], s = 0;
#pragma unroll
    for (int v = 0; v < SCAN_V; v++) { c[v] = b0 + v < n ? data[b0 + v] : 0; s += c[v]; }
    uint32_t total;
    uint32_t run = tile_sum[blockIdx.x] + idx_block_exscan(s, sh, total);
#pragma unroll
    for (int v = 0; v < SCAN_V; v++) if (b0 + v < n) { data[b0 + v] = run; run += c[v]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = tile_sum[ntiles];
}

// ---- items of a segmentation: a segment of len <= SEG_LMAX is one item; a longer one is split into chunks of SEG_LMAX (the
// last one shorter) that write partial sums, added up by k_seg_fixup.  Items are emitted by decreasing length so that the
// threads of a warp of k_segsum carry similar trip counts.
__global__ void __launch_bounds__(IDX_T) k_seg_hist(const uint32_t* __restrict__ seg_ptr, size_t nseg, PlanCounts* counts) {
    __shared__ uint32_t hist[SEG_LMAX + 1];
    __shared__ uint32_t nfix, npart;
    if (threadIdx.x <= SEG_LMAX) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { nfix = 0; npart = 0; }
    __syncthreads();
    for (size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x; s < nseg; s += (size_t)gridDim.x * blockDim.x) {
        const uint32_t len = seg_ptr[s + 1] - seg_ptr[s];
        if (len == 0) continue;
        if (len <= SEG_LMAX) { atomicAdd(&hist[len], 1u); continue; }
        const uint32_t full = len / SEG_LMAX, rem = len % SEG_LMAX;
        atomicAdd(&hist[SEG_LMAX], full);
        if (rem) atomicAdd(&hist[rem], 1u);
        atomicAdd(&nfix, 1u);
        atomicAdd(&npart, full + (rem ? 1u : 0u));
    }
    __syncthreads();
    if (threadIdx.x <= SEG_LMAX && hist[threadIdx.x]) atomicAdd(&counts->hist[threadIdx.x], hist[threadIdx.x]);
    if (threadIdx.x == 0) { if (nfix) atomicAdd(&counts->n_fix, nfix); if (npart) atomicAdd(&counts->n_partials, npart); }
}
// short segments: one item each; split segments: reserve their fixup and partial slots, list them for k_seg_emit_long
__global__ void __launch_bounds__(IDX_T) k_seg_emit(const uint32_t* __restrict__ seg_ptr, size_t nseg, PlanCursors* cur, SegItem* __restrict__ items,
                                                    SegFixup* __restrict__ fix) {
    for (size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x; s < nseg; s += (size_t)gridDim.x * blockDim.x) {
        const uint32_t beg = seg_ptr[s], len = seg_ptr[s + 1] - beg;
        if (len == 0) continue;
        if (len <= SEG_LMAX) {
            const uint32_t pos = atomicAdd(&cur->item[len], 1u);
            items[pos] = SegItem{beg, len, (uint32_t)s};
            continue;
        }
        const uint32_t cnt = (len + SEG_LMAX - 1) / SEG_LMAX;
        const uint32_t f = atomicAdd(&cur->fix, 1u), p0 = atomicAdd(&cur->partial, cnt);
        fix[f] = SegFixup{(uint32_t)s, p0, cnt};
    }
}
// one CTA per split segment: its chunks
__global__ void __launch_bounds__(IDX_T) k_seg_emit_long(const uint32_t* __restrict__ seg_ptr, const SegFixup* __restrict__ fix, PlanCursors* cur,
                                                         SegItem* __restrict__ items) {
    const SegFixup f = fix[blockIdx.x];
    const uint32_t beg = seg_ptr[f.seg], len = seg_ptr[f.seg + 1] - beg;
    for (uint32_t i = threadIdx.x; i < f.pcount; i += blockDim.x) {
        const uint32_t o = i * SEG_LMAX, l = len - o < SEG_LMAX ? len - o : SEG_LMAX;
        const uint32_t pos = atomicAdd(&cur->item[l], 1u);
        items[pos] = SegItem{beg + o, l, (f.pstart + i) | 0x80000000u};
    }
}
}  // namespace

void launch_idx_validate_rows(const uint64_t* row_ptr, size_t n, uint64_t nnz_cap, uint32_t* err, cudaStream_t stream) {
    SB_LAUNCH(k_idx_validate_rows, grid_for(n + 1, IDX_T, 8), IDX_T, 0, stream, row_ptr, n, nnz_cap, err);
}
void launch_idx_validate_cols(const uint32_t* col, size_t nnz, uint32_t ncols, uint32_t* err, cudaStream_t stream) {
    if (nnz) SB_LAUNCH(k_idx_validate_cols, grid_for(nnz, IDX_T, 8), IDX_T, 0, stream, col, nnz, ncols, err);
}
void launch_idx_row_segments(const uint64_t* row_ptr, size_t lo, size_t nl, uint32_t base, bool close, uint32_t* seg_ptr, cudaStream_t stream) {
    SB_LAUNCH(k_idx_row_segments, grid_for(nl + 1, IDX_T, 8), IDX_T, 0, stream, row_ptr, lo, nl, base, close, seg_ptr);
}
void launch_idx_flags(const uint32_t* col, const Fr* val, uint32_t* idx, size_t cnt, cudaStream_t stream) {
    if (cnt) SB_LAUNCH(k_idx_flags, grid_for(cnt, IDX_T, 8), IDX_T, 0, stream, col, val, idx, cnt);
}
void launch_idx_col_count(const uint32_t* col, size_t nnz, uint32_t lo, uint32_t hi, uint32_t* cnt, cudaStream_t stream) {
    if (nnz) SB_LAUNCH(k_idx_col_count, grid_for(nnz, IDX_T, 8), IDX_T, 0, stream, col, nnz, lo, hi, cnt);
}
void launch_idx_exscan(uint32_t* data, size_t n, uint32_t* ws, cudaStream_t stream) {
    const size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    SB_LAUNCH(k_scan_tile_sums, (int)ntiles, IDX_T, 0, stream, data, n, ws);
    SB_LAUNCH(k_scan_tile_offsets, 1, IDX_T, 0, stream, ws, ntiles);
    SB_LAUNCH(k_scan_apply, (int)ntiles, IDX_T, 0, stream, data, n, ws, ntiles);
}
void launch_idx_col_scatter(const uint64_t* row_ptr, size_t n_rows, const uint32_t* col, const Fr* val, size_t nnz, uint32_t lo, uint32_t hi,
                            uint32_t tag_base, uint32_t* cursor, uint32_t* idx_out, Fr* val_out, cudaStream_t stream) {
    if (nnz) SB_LAUNCH(k_idx_col_scatter, grid_for(nnz, IDX_T, 8), IDX_T, 0, stream, row_ptr, n_rows, col, val, nnz, lo, hi, tag_base, cursor, idx_out, val_out);
}
void launch_seg_hist(const uint32_t* seg_ptr, size_t nseg, PlanCounts* counts, cudaStream_t stream) {
    SB_LAUNCH(k_seg_hist, grid_for(nseg, IDX_T, 8), IDX_T, 0, stream, seg_ptr, nseg, counts);
}
void launch_seg_emit(const uint32_t* seg_ptr, size_t nseg, PlanCursors* cur, SegItem* items, SegFixup* fix, uint32_t n_fix, cudaStream_t stream) {
    SB_LAUNCH(k_seg_emit, grid_for(nseg, IDX_T, 8), IDX_T, 0, stream, seg_ptr, nseg, cur, items, fix);
    if (n_fix) SB_LAUNCH(k_seg_emit_long, (int)n_fix, IDX_T, 0, stream, seg_ptr, (const SegFixup*)fix, cur, items);
}
