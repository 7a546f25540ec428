// indexer.cuh -- device-side construction of the sparse-product plans.
//
// The reference indexes by cloning the three constraint matrices (src/ahp/indexer.rs:41-64) and wrapping each in a
// MatrixExtension (src/data_structures/r1cs_reader.rs:36-70) that later walks its rows for sum_over_y (:75-85) and
// eval_on_x (:91-117).  Here the index holds two segmented plans -- by row for Az/Bz/Cz, by column (the transpose of
// all three matrices at once) for M(r_x, .) -- and they are BUILT ON THE DEVICE from the caller's CSR arrays: the host
// uploads the arrays, reads back a handful of counters to size the allocations, and hashes the transcript prefix.
#pragma once
#include "kernels_fr.cuh"

// counters of one plan, filled by launch_seg_hist: items per length (longest SEG_LMAX), split segments, partial slots
struct PlanCounts {
    uint32_t hist[SEG_LMAX + 1];
    uint32_t n_fix, n_partials, pad;
};
// cursors of the emit pass (device): next free item position per length, next fixup / partial slot
struct PlanCursors {
    uint32_t item[SEG_LMAX + 1];
    uint32_t fix, partial, pad;
};

enum { IDX_ERR_ROWPTR = 1, IDX_ERR_COL = 2 };
// *err |= IDX_ERR_ROWPTR when row_ptr[0..n] is not non-decreasing from 0 up to at most `nnz_cap`
void launch_idx_validate_rows(const uint64_t* row_ptr, size_t n, uint64_t nnz_cap, uint32_t* err, cudaStream_t stream);
// *err |= IDX_ERR_COL when some col[e] >= ncols  (r1cs_reader.rs:55-62)
void launch_idx_validate_cols(const uint32_t* col, size_t nnz, uint32_t ncols, uint32_t* err, cudaStream_t stream);
// seg_ptr[r] = base + row_ptr[lo + r] - row_ptr[lo] for r < nl (and seg_ptr[nl] when `close`)
void launch_idx_row_segments(const uint64_t* row_ptr, size_t lo, size_t nl, uint32_t base, bool close, uint32_t* seg_ptr, cudaStream_t stream);
// idx[e] = col[e] | (val[e] == 1 ? SEG_UNIT_FLAG : 0)
void launch_idx_flags(const uint32_t* col, const Fr* val, uint32_t* idx, size_t cnt, cudaStream_t stream);
// cnt[y - lo] += 1 for every entry with lo <= col[e] < hi
void launch_idx_col_count(const uint32_t* col, size_t nnz, uint32_t lo, uint32_t hi, uint32_t* cnt, cudaStream_t stream);
// in-place exclusive scan of data[0..n); data[n] receives the total (data has n + 1 entries); ws: >= n / 1024 + 2 words
void launch_idx_exscan(uint32_t* data, size_t n, uint32_t* ws, cudaStream_t stream);
// transpose scatter: entry e of a matrix (row found by binary search in row_ptr) goes to position cursor[col - lo]++ of the
// column plan with gather index tag_base + row
void launch_idx_col_scatter(const uint64_t* row_ptr, size_t n_rows, const uint32_t* col, const Fr* val, size_t nnz, uint32_t lo, uint32_t hi,
                            uint32_t tag_base, uint32_t* cursor, uint32_t* idx_out, Fr* val_out, cudaStream_t stream);
// item statistics of a segmentation (seg_ptr has nseg + 1 entries)
void launch_seg_hist(const uint32_t* seg_ptr, size_t nseg, PlanCounts* counts, cudaStream_t stream);
// emit the items (sorted by decreasing length through the per-length cursors), the fixups of split segments
void launch_seg_emit(const uint32_t* seg_ptr, size_t nseg, PlanCursors* cur, SegItem* items, SegFixup* fix, uint32_t n_fix, cudaStream_t stream);
