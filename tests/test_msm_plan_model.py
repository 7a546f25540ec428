"""Index arithmetic of the MSM accumulation plan, restated in numpy and checked for its invariants (CPU only).

This is a model of `k_scan_plan` / `k_seg_accum` / `k_bucket_reduce1` in r1cs-spartan_b200/csrc/msm.cu -- the same
formulas, executed for every thread index -- with integers standing in for curve points (a "point" is the set of
entry indices it has absorbed).  It pins down what the CUDA code relies on:
  * level 0 covers every entry of every bucket run exactly once, in chunks of at most S0;
  * every later level sums groups of at most S1 partial sums of ONE bucket, until one point per bucket is left;
  * with the length-sorted accumulation order the chunks owned by 32 consecutive threads have (almost) equal length,
    and the bucket reduction finds bucket b's point through the inverse permutation.
The GPU tests check the real kernels against the oracle; this file documents and guards the arithmetic itself.
"""
import numpy as np
import pytest


def msm_levels(maxrun, s0, s1, max_levels=16):
    levels, cover = 1, s0
    while cover < maxrun and levels < max_levels:
        cover *= s1
        levels += 1
    return levels


def plan_model(counts, s0, s1, sort):
    """k_scan_plan: offsets, accumulation order (perm / invperm) and the chunk plan of every level."""
    B = len(counts)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    levels = msm_levels(int(counts.max(initial=0)), s0, s1)
    if sort:
        nch = -(-counts // s0)
        key = np.where(nch > 0, -(-counts // np.maximum(nch, 1)), 0)
        perm = np.argsort(-key, kind="stable")              # any order inside a key bin is allowed
        invperm = np.empty(B, dtype=np.int64); invperm[perm] = np.arange(B)
    else:
        perm = invperm = np.arange(B)
    plans, div = [], s0
    for _ in range(levels):
        per = -(-counts[perm] // div)
        plans.append(np.concatenate([[0], np.cumsum(per)]))
        div *= s1
    return offsets, perm, invperm, plans, levels


def accumulate_model(counts, s0, s1, sort):
    """k_seg_accum for every thread of every level, then the lookup k_bucket_reduce1 does."""
    offsets, perm, invperm, plans, levels = plan_model(counts, s0, s1, sort)
    B = len(counts)
    pts, lens0 = None, None
    for l in range(levels):
        chunk_start = plans[l]
        seg_off = offsets if l == 0 else plans[l - 1]
        out = []
        lens = []
        for p in range(int(chunk_start[B])):
            lo = int(np.searchsorted(chunk_start, p, side="right")) - 1      # last k with chunk_start[k] <= p
            j, nch = p - int(chunk_start[lo]), int(chunk_start[lo + 1] - chunk_start[lo])
            sb = int(perm[lo]) if l == 0 else lo
            off, cnt = int(seg_off[sb]), int(seg_off[sb + 1] - seg_off[sb])
            beg, end = off + j * cnt // nch, off + (j + 1) * cnt // nch
            assert 0 < end - beg <= (s0 if l == 0 else s1)
            lens.append(end - beg)
            if l == 0:
                out.append(frozenset(range(beg, end)))
            else:
                acc = frozenset()
                for e in range(beg, end):
                    assert not (acc & pts[e])
                    acc |= pts[e]
                out.append(acc)
        pts = out
        if l == 0:
            lens0 = lens
    last = plans[levels - 1]
    for b in range(B):                                        # bucket -> its single point (or none)
        k = int(invperm[b])
        o = int(last[k])
        if last[k + 1] > o:
            assert last[k + 1] == o + 1
            assert pts[o] == frozenset(range(int(offsets[b]), int(offsets[b + 1])))
        else:
            assert counts[b] == 0
    return lens0


@pytest.mark.parametrize("sort", [False, True])
@pytest.mark.parametrize("s0,s1", [(48, 3), (24, 3), (3, 2), (2, 2), (7, 5)])
def test_plan_covers_every_bucket_exactly_once(s0, s1, sort):
    rng = np.random.default_rng(s0 * 100 + s1)
    for counts in (rng.poisson(60, 40), rng.integers(0, 4, 50), np.array([0, 0, 500, 1, 0, 49, 48, 47]), np.array([1]),
                   np.zeros(5, dtype=np.int64), np.array([1000] * 3)):
        accumulate_model(np.asarray(counts, dtype=np.int64), s0, s1, sort)


def test_sorted_order_gives_warps_equal_length_chunks():
    rng = np.random.default_rng(1)
    counts = rng.poisson(256, 4096).astype(np.int64)          # the largest ladder level: runs of 256 +- 16
    eff = {}
    for sort in (False, True):
        lens = np.array(accumulate_model(counts, 48, 3, sort))
        pad = (-len(lens)) % 32
        warps = np.concatenate([lens, np.zeros(pad, dtype=lens.dtype)]).reshape(-1, 32)
        eff[sort] = lens.sum() / (warps.max(axis=1).sum() * 32)      # active lanes / issued lanes
        if sort:
            assert (np.diff(lens) <= 1).all() and lens[:32].max() == lens.max() and lens[-32:].min() == lens.min()   # longest first (a run's own chunks differ by at most 1)
    assert eff[True] > 0.985 and eff[False] < 0.95, eff
