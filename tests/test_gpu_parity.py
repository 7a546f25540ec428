"""GPU parity tests: every stage of the CUDA path against the CPU oracle, through the C ABI.

Bar: bit-exact (all arithmetic is exact field / group arithmetic).  The oracle follows the reference
literally (log_n separate eq tables, 2 log_n + 3 sumcheck tables, duplicated-scalar G2 MSMs), so these
tests also validate the algebraic rewrites the CUDA path uses.
"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FR_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FQ_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


@pytest.fixture(scope="module")
def sb():
    import r1cs_spartan_b200 as m
    return m


def _edge_and_random(mod, n, seed):
    rnd = random.Random(seed)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 255 if mod > (1 << 255) else 1 << 200]
    edge = [e % mod for e in edge]
    vals = edge + [rnd.randrange(mod) for _ in range(n - len(edge))]
    return vals


# ---------------------------------------------------------------- field arithmetic (K0)
@pytest.mark.parametrize("field", ["fr", "fq"])
def test_field_ops_match_oracle(sb, oracle, gpu_ctx, field):
    mod = FR_MOD if field == "fr" else FQ_MOD
    n = 4096
    a = _edge_and_random(mod, n, 1); b = list(reversed(_edge_and_random(mod, n, 2)))
    to_m = oracle.fr_from_ints if field == "fr" else oracle.fq_from_ints
    binop = oracle.fr_binop if field == "fr" else oracle.fq_binop
    A, B = to_m(a), to_m(b)
    for op in ("add", "sub", "mul"):
        got = gpu_ctx.field_binop(field, op, A, B)
        assert np.array_equal(got, binop(op, A, B)), (field, op)
    # the PTX carry-chain product against the portable 64-bit-accumulator product on the device itself
    assert np.array_equal(gpu_ctx.field_binop(field, "mul", A, B), gpu_ctx.field_binop(field, "mul_portable", A, B))


# ---------------------------------------------------------------- eq table (K2)
@pytest.mark.parametrize("dim", [1, 2, 5, 9, 12])
def test_eq_table_is_product_of_reference_tables(sb, oracle, gpu_ctx, dim):
    t = oracle.fr_rand(1000 + dim, dim)
    got = sb.eq_extension(t, ctx=gpu_ctx)
    ref = oracle.eq_extension(t)          # (dim, 2^dim, 4): the reference's dim separate tables
    prod = ref[0]
    for i in range(1, dim):
        prod = oracle.fr_binop("mul", prod, ref[i])
    assert np.array_equal(got, prod)


def test_eq_table_boolean_point_is_indicator(sb, oracle, gpu_ctx):
    # reference test eq::functionality_test (eq.rs:30-46): t = 0b101101001 over 9 variables
    tbits = 0b101101001
    t = oracle.fr_from_ints([(tbits >> i) & 1 for i in range(9)])
    got = sb.eq_extension(t, ctx=gpu_ctx)
    ints = oracle.fr_to_ints(got)
    assert ints[tbits] == 1 and sum(ints) == 1


# ---------------------------------------------------------------- sparse products (K1, K5)
def _random_csr(oracle, log_n, nnz, seed, unit_fraction=0.5):
    rnd = random.Random(seed)
    n = 1 << log_n
    cells = set()
    while len(cells) < nnz:
        cells.add((rnd.randrange(n), rnd.randrange(n)))
    rows = [[] for _ in range(n)]
    for (x, y) in sorted(cells):
        rows[x].append(y)
    row_ptr = np.zeros(n + 1, dtype=np.uint64)
    col = []
    for x in range(n):
        row_ptr[x + 1] = row_ptr[x] + len(rows[x])
        col += rows[x]
    vals = oracle.fr_rand(seed + 7, nnz)
    one = oracle.fr_from_ints([1])[0]
    for e in range(nnz):
        if rnd.random() < unit_fraction:
            vals[e] = one
    return row_ptr, np.array(col, dtype=np.uint32), vals


@pytest.mark.parametrize("log_n,nnz", [(3, 10), (6, 512), (8, 3000)])
def test_sum_over_y_and_eval_on_x_random_matrices(sb, oracle, gpu_ctx, log_n, nnz):
    mats = [_random_csr(oracle, log_n, nnz, 10 * log_n + k) for k in range(3)]
    pk = sb.MLProofForR1CS.index(*mats, ctx=gpu_ctx)
    ocs = oracle.R1CS.from_csr(log_n, mats)
    z = oracle.fr_rand(5, 1 << log_n)
    got = pk.sum_over_y(z)
    for k in range(3):
        assert np.array_equal(got[k], ocs.sum_over_y(k, z)), k
    r_x = oracle.fr_rand(6, log_n)
    for k in range(3):
        assert np.array_equal(pk.eval_on_x(r_x, which=k), ocs.eval_on_x(k, r_x)), k
    r_abc = oracle.fr_rand(8, 3)
    comb = None
    for k in range(3):
        term = oracle.fr_binop("mul", ocs.eval_on_x(k, r_x), np.repeat(r_abc[k:k + 1], 1 << log_n, axis=0))
        comb = term if comb is None else oracle.fr_binop("add", comb, term)
    assert np.array_equal(pk.eval_on_x(r_x, r_abc=r_abc), comb)


def test_eval_on_x_boolean_point_returns_matrix_row(sb, oracle, gpu_ctx):
    # reference test r1cs_reader::test_eval_on_x_sanity (r1cs_reader.rs:128-145): x = 0b110010, LSB-first point
    log_n = 6
    mats = [_random_csr(oracle, log_n, 1 << 9, 77 + k, unit_fraction=0.0) for k in range(3)]
    pk = sb.MLProofForR1CS.index(*mats, ctx=gpu_ctx)
    x = 0b110010
    point = oracle.fr_from_ints([(x >> i) & 1 for i in range(log_n)])
    got = pk.eval_on_x(point, which=0)
    row_ptr, col, val = mats[0]
    expect = np.zeros((1 << log_n, 4), dtype=np.uint64)
    for e in range(int(row_ptr[x]), int(row_ptr[x + 1])):
        expect[col[e]] = val[e]
    assert np.array_equal(got, expect)


def test_synthetic_circuit_long_row_and_column(sb, oracle, gpu_ctx):
    # the benchmark circuit has one ~n-entry row (A, B) and an ~n/2-entry column 0 (B): exercises the split path
    log_n = 10
    cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 4242)
    pk = sb.MLProofForR1CS.index(*cs.mats, ctx=gpu_ctx)
    ocs = oracle.R1CS.from_csr(log_n, cs.mats)
    z = np.concatenate([cs.v, cs.w])
    assert ocs.is_satisfied(z)
    got = pk.sum_over_y(z)
    for k in range(3):
        assert np.array_equal(got[k], ocs.sum_over_y(k, z))
    r_x = oracle.fr_rand(9, log_n)
    for k in range(3):
        assert np.array_equal(pk.eval_on_x(r_x, which=k), ocs.eval_on_x(k, r_x))


def test_device_indexer_edge_cases(sb, oracle, gpu_ctx):
    # the plans are built by kernels (csrc/indexer.cu): an empty matrix, a row and a column longer than SEG_LMAX = 32 with a
    # remainder chunk (100 = 3 * 32 + 4 and 70 = 2 * 32 + 6 entries), a row that is exactly one full chunk, unit and non-unit
    # coefficients side by side, unsorted columns inside a row, and empty rows at both ends.  (Columns are distinct within a
    # row, as upstream `to_matrices` produces them: with duplicates the reference's own sum_over_y adds both terms,
    # r1cs_reader.rs:79-83, while its eval_on_x keeps one of them, :100-108 -- the library adds them in both.)
    log_n = 7
    n = 1 << log_n
    rnd = random.Random(99)
    one = oracle.fr_from_ints([1])[0]

    def build(rows):
        row_ptr = np.zeros(n + 1, dtype=np.uint64); col = []
        for x in range(n):
            row_ptr[x + 1] = row_ptr[x] + len(rows[x]); col += rows[x]
        vals = oracle.fr_rand(len(col) + 3, max(len(col), 1))[:len(col)]
        for e in range(len(col)):
            if rnd.random() < 0.4:
                vals[e] = one
        return row_ptr, np.array(col, dtype=np.uint32), vals

    rows_a = [[] for _ in range(n)]
    rows_a[5] = rnd.sample(range(n), 100)                              # long row, columns in random order
    rows_a[6] = list(range(32))                                        # exactly one full chunk
    for x in range(10, 80):
        rows_a[x] = [3] + rnd.sample(range(4, n), rnd.randrange(3))                # column 3 holds 70 entries
    rows_b = [rnd.sample(range(n), rnd.randrange(4)) for _ in range(n)]
    rows_b[0] = []; rows_b[n - 1] = []
    rows_c = [[] for _ in range(n)]                                    # an empty matrix
    mats = [build(rows_a), build(rows_b), build(rows_c)]
    pk = sb.MLProofForR1CS.index(*mats, ctx=gpu_ctx)
    ocs = oracle.R1CS.from_csr(log_n, mats)
    z = oracle.fr_rand(17, n)
    got = pk.sum_over_y(z)
    for k in range(3):
        assert np.array_equal(got[k], ocs.sum_over_y(k, z)), k
    r_x = oracle.fr_rand(18, log_n)
    for k in range(3):
        assert np.array_equal(pk.eval_on_x(r_x, which=k), ocs.eval_on_x(k, r_x)), k
    plan_ms, hash_ms = pk.timing()
    assert plan_ms >= 0 and hash_ms >= 0


# ---------------------------------------------------------------- keygen / MSM / commitment (K7-K9)
@pytest.fixture(scope="module")
def small_pp(sb, oracle, gpu_ctx):
    nv = 6
    g, h = oracle.generators()
    t = oracle.fr_rand(31337, nv)
    return nv, g, h, t, sb.MLPolyCommit.keygen(nv, g, h, t, keep_all_levels=True, ctx=gpu_ctx), oracle.PP.keygen_with(nv, g, h, t)


def test_keygen_matches_oracle(sb, oracle, small_pp):
    nv, g, h, t, pp, opp = small_pp
    for level in range(nv):
        assert np.array_equal(pp.export(1, level), opp.g1(level)), ("g1", level)
        assert np.array_equal(pp.export(2, level), opp.g2(level)), ("g2", level)
    assert np.array_equal(pp.g_mask_random(), opp.g_mask())


def test_pp_load_equals_keygen(sb, oracle, gpu_ctx, small_pp):
    nv, g, h, t, pp, opp = small_pp
    pp2 = sb.MLPolyCommit.load(nv, opp.g1(0), [opp.g2(i) for i in range(nv)], h, ctx=gpu_ctx)
    z = oracle.fr_rand(3, 1 << nv)
    point = oracle.fr_rand(4, nv)
    assert np.array_equal(sb.MLPolyCommit.commit(pp2, z)[1], sb.MLPolyCommit.commit(pp, z)[1])
    e1, (_, p1) = sb.MLPolyCommit.open(pp2, z, point)
    e2, (_, p2) = sb.MLPolyCommit.open(pp, z, point)
    assert np.array_equal(e1, e2) and np.array_equal(p1, p2)


@pytest.mark.parametrize("n", [1, 2, 3, 33, 1000])
def test_msm_matches_oracle(sb, oracle, gpu_ctx, n):
    g, h = oracle.generators()
    ks = oracle.fr_rand(50 + n, n)
    b1 = np.stack([oracle.g1_mul(g, ks[i]) for i in range(n)])
    b2 = np.stack([oracle.g2_mul(h, ks[i]) for i in range(n)])
    s = oracle.fr_rand(60 + n, n)
    assert np.array_equal(sb.multi_scalar_mul(1, b1, s, ctx=gpu_ctx), oracle.msm_g1(b1, s))
    assert np.array_equal(sb.multi_scalar_mul(2, b2, s, ctx=gpu_ctx), oracle.msm_g2(b2, s))


def test_msm_structured_scalars(sb, oracle, gpu_ctx):
    # zero / one / minus-one / repeated scalars and repeated bases: the exceptional cases of the group law
    g, _ = oracle.generators()
    n = 64
    ks = oracle.fr_rand(5, n)
    bases = np.stack([oracle.g1_mul(g, ks[i % 4]) for i in range(n)])       # only 4 distinct bases
    vals = [0, 1, FR_MOD - 1, 2, 5, 5, 5, 5] * (n // 8)
    s = oracle.fr_from_ints(vals)
    assert np.array_equal(sb.multi_scalar_mul(1, bases, s, ctx=gpu_ctx), oracle.msm_g1(bases, s))
    zeros = oracle.fr_from_ints([0] * n)
    assert not sb.multi_scalar_mul(1, bases, zeros, ctx=gpu_ctx).any()       # infinity = all-zero encoding


def test_commit_and_open_match_oracle(sb, oracle, small_pp):
    nv, g, h, t, pp, opp = small_pp
    z = oracle.fr_rand(71, 1 << nv)
    point = oracle.fr_rand(72, nv)
    got_nv, com = sb.MLPolyCommit.commit(pp, z)
    assert got_nv == nv and np.array_equal(com, opp.commit(z))
    # reference test commit::commit_test (commit.rs:54-66): commit == g * f(t)
    assert np.array_equal(com, oracle.g1_mul(g, oracle.mle_eval(z, t)))
    ev, (hh, proofs) = sb.MLPolyCommit.open(pp, z, point)
    oev, oproofs = opp.open(z, point)
    assert np.array_equal(ev, oev) and np.array_equal(proofs, oproofs) and np.array_equal(hh, h)


def test_open_at_padded_public_point(sb, oracle, small_pp):
    # prover_second_round opens at (r_v, 0, .., 0) (prover.rs:152): zero coordinates make r' = even
    nv, g, h, t, pp, opp = small_pp
    z = oracle.fr_rand(81, 1 << nv)
    point = np.zeros((nv, 4), dtype=np.uint64)
    point[:2] = oracle.fr_rand(82, 2)
    ev, (_, proofs) = sb.MLPolyCommit.open(pp, z, point)
    oev, oproofs = opp.open(z, point)
    assert np.array_equal(ev, oev) and np.array_equal(proofs, oproofs)


# ---------------------------------------------------------------- AHP rounds with explicit verifier messages
def test_interactive_rounds_match_literal_sumcheck(sb, oracle, gpu_ctx, small_pp):
    # the shape of reference test ahp::tests::test_small (ahp/tests.rs:8-75) with caller-chosen challenges
    nv, g, h, t, pp, opp = small_pp
    log_n, log_v = nv, 2
    cs = sb.SyntheticR1CS(1 << log_v, (1 << log_n) - (1 << log_v), 1, 555)
    pk = sb.MLProofForR1CS.index(*cs.mats, ctx=gpu_ctx)
    ocs = oracle.R1CS.from_csr(log_n, cs.mats)
    z = np.concatenate([cs.v, cs.w])
    P = sb.MLProofForR1CS
    st = P.prover_init(pk, cs.v, cs.w)
    st, m1 = P.prover_first_round(st, pp)
    assert np.array_equal(m1["commitment"][1], opp.commit(z))
    r_v = oracle.fr_rand(1, log_v)
    st, m2 = P.prover_second_round(st, r_v, pp)
    pt = np.zeros((log_n, 4), dtype=np.uint64); pt[:log_v] = r_v
    oev, oproofs = opp.open(z, pt)
    assert np.array_equal(m2["z_rv_0"], oev) and np.array_equal(m2["proof_for_z_rv_0"][1], oproofs)
    tor = oracle.fr_rand(2, log_n)
    st, m3 = P.prover_third_round(st, tor)
    assert m3["ml_index_info"] == (log_n + 2, log_n)
    az, bz, cz = st.export_abc()
    oabc = [ocs.sum_over_y(k, z) for k in range(3)]
    assert all(np.array_equal(x, y) for x, y in zip((az, bz, cz), oabc))
    # literal first sumcheck: products [Az, Bz, eq_0..] and [-Cz, eq_0..]
    chal = oracle.fr_rand(3, log_n)
    eq = oracle.eq_extension(tor)
    ncz = oracle.fr_binop("sub", np.zeros_like(cz), oabc[2])
    tables = np.stack([oabc[0], oabc[1], ncz] + [eq[i] for i in range(log_n)])
    eq_idx = list(range(3, 3 + log_n))
    lit = oracle.sumcheck_prove(tables, [[0, 1] + eq_idx, [2] + eq_idx], chal)
    vm = None
    for j in range(log_n):
        st, evals = P.prove_first_sumcheck_round(st, vm)
        assert np.array_equal(evals, lit[j]), j
        vm = chal[j]
    st, m4 = P.prove_fourth_round(st, vm)
    for k, name in enumerate(("va", "vb", "vc")):
        assert np.array_equal(m4[name], oracle.mle_eval(oabc[k], chal))
    r_abc = oracle.fr_rand(4, 3)
    st, m5 = P.prove_fifth_round(st, r_abc[0], r_abc[1], r_abc[2])
    assert m5["index_info"] == (2, log_n)
    tabs = []
    for k in range(3):
        tabs.append(oracle.fr_binop("mul", ocs.eval_on_x(k, chal), np.repeat(r_abc[k:k + 1], 1 << log_n, axis=0)))
    tables2 = np.stack(tabs + [z])
    chal2 = oracle.fr_rand(5, log_n)
    lit2 = oracle.sumcheck_prove(tables2, [[0, 3], [1, 3], [2, 3]], chal2)
    vm = None
    for j in range(log_n):
        st, evals = P.prove_second_sumcheck_round(st, vm)
        assert np.array_equal(evals, lit2[j]), j
        vm = chal2[j]
    m6 = P.prove_sixth_round(st, vm, pp)
    oev, oproofs = opp.open(z, chal2)
    assert np.array_equal(m6["z_ry"], oev) and np.array_equal(m6["proof_for_z_ry"][1], oproofs)


# ---------------------------------------------------------------- the non-interactive argument
def _prove_both(sb, oracle, ctx, log_n, num_public, density, seed):
    cs = sb.SyntheticR1CS(num_public, (1 << log_n) - num_public, density, seed)
    g, h = oracle.generators()
    t = oracle.fr_rand(seed ^ 0xABCDEF, log_n)
    pp = sb.MLPolyCommit.keygen(log_n, g, h, t, ctx=ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
    proof, tr = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp, trace=True)
    ocs = oracle.R1CS.from_csr(log_n, cs.mats)
    opp = oracle.PP.keygen_with(log_n, g, h, t)
    oproof, otr = oracle.prove(ocs, opp, cs.v, cs.w)
    return proof, tr, oproof, otr


@pytest.mark.parametrize("log_n,num_public,density", [(3, 4, 0), (4, 4, 0), (6, 8, 30), (8, 32, 0), (10, 32, 0), (11, 32, 127)])
def test_prove_bytes_and_trace_match_oracle(sb, oracle, gpu_ctx, log_n, num_public, density):
    proof, tr, oproof, otr = _prove_both(sb, oracle, gpu_ctx, log_n, num_public, density, 0x5EED0000 + log_n)
    assert len(proof) == len(oproof)
    _assert_trace_equal(tr, otr)     # intermediates first, so that a mismatch is reported at the earliest stage
    assert proof == oproof


def test_prove_is_deterministic_and_handles_are_reusable(sb, oracle, gpu_ctx):
    log_n = 7
    cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 1)
    g, h = oracle.generators()
    pp = sb.MLPolyCommit.keygen(log_n, g, h, oracle.fr_rand(2, log_n), ctx=gpu_ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=gpu_ctx)
    p1 = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp)
    p2 = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp)
    assert p1 == p2
    # a different witness gives a different proof with the same handles
    w2 = cs.w.copy(); w2[0] = oracle.fr_from_ints([12345])[0]
    assert sb.MLArgumentForR1CS.prove(pk, cs.v, w2, pp) != p1


def test_context_may_be_closed_before_its_handles(sb, oracle):
    # handles use their context's device and streams when they are destroyed; the context counts them, so closing it
    # first only marks it and the last handle to go completes the destruction (use-after-free otherwise)
    ctx = sb.Context(0)
    cs = sb.SyntheticR1CS(4, 12, 0, 9)
    g, h = oracle.generators()
    pp = sb.MLPolyCommit.keygen(4, g, h, oracle.fr_rand(2, 4), ctx=ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
    wit = sb.Witness(pk, cs.v, cs.w)
    proof = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp)
    ctx.close()                       # deferred: three handles are alive
    wit.close(); pk.close(); pp.close()
    ctx2 = sb.Context(0)              # a fresh context works and gives the same proof
    pp2 = sb.MLPolyCommit.keygen(4, g, h, oracle.fr_rand(2, 4), ctx=ctx2)
    pk2 = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx2)
    assert sb.MLArgumentForR1CS.prove(pk2, cs.v, cs.w, pp2) == proof
    pk2.close(); pp2.close(); ctx2.close()


def test_proof_size_formula(sb):
    L = sb.load_library()
    # proof sizes derived from the reference's struct layout
    assert [L.sb_proof_size(l) for l in (8, 12, 16, 20, 24)] == [5720, 9880, 15064, 21272, 28504]


# ---------------------------------------------------------------- error behaviour (reference: Error::InvalidArgument)
def test_invalid_arguments(sb, oracle, gpu_ctx):
    log_n = 4
    cs = sb.SyntheticR1CS(4, 12, 0, 9)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=gpu_ctx)
    g, h = oracle.generators()
    pp = sb.MLPolyCommit.keygen(log_n, g, h, oracle.fr_rand(2, log_n), ctx=gpu_ctx)
    with pytest.raises(sb.InvalidArgument):      # |v| + |w| != n  (prover.rs:117-119)
        sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w[:-1], pp)
    with pytest.raises(sb.InvalidArgument):      # |v| not a power of two (prover.rs:114-116)
        sb.MLArgumentForR1CS.prove(pk, np.concatenate([cs.v, cs.w[:1]])[:3], np.concatenate([cs.v[3:], cs.w]), pp)
    rp, col, val = cs.mats[0]
    bad_col = col.copy(); bad_col[0] = 1 << log_n
    with pytest.raises(sb.InvalidArgument):      # sparse index out of bound (r1cs_reader.rs:55-62)
        sb.MLArgumentForR1CS.index((rp, bad_col, val), cs.mats[1], cs.mats[2], ctx=gpu_ctx)
    with pytest.raises(sb.InvalidArgument):      # not a power of two (indexer.rs:49)
        m3 = (rp[:4].copy(), col[:int(rp[3])], val[:int(rp[3])])
        sb.MLArgumentForR1CS.index(m3, m3, m3, ctx=gpu_ctx)
    pp5 = sb.MLPolyCommit.keygen(5, g, h, oracle.fr_rand(3, 5), ctx=gpu_ctx)
    with pytest.raises(sb.InvalidArgument):      # parameter size mismatch
        sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp5)
    st = sb.MLProofForR1CS.prover_init(pk, cs.v, cs.w)
    with pytest.raises(sb.InvalidArgument):      # rounds out of order
        sb.MLProofForR1CS.prover_third_round(st, oracle.fr_rand(1, log_n))


# ---------------------------------------------------------------- acceptance: the CPU verifier accepts GPU proofs
def test_gpu_proof_is_accepted_by_the_verifier(sb, oracle, gpu_ctx):
    # benchmark.rs:35-47: prove -> serialize -> deserialize -> verify, with the prover on the GPU
    log_n, log_v = 8, 2
    cs = sb.SyntheticR1CS(1 << log_v, (1 << log_n) - (1 << log_v), 1, 12345)
    g, h = oracle.generators()
    t = oracle.fr_rand(777, log_n)
    pp = sb.MLPolyCommit.keygen(log_n, g, h, t, ctx=gpu_ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=gpu_ctx)
    proof = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp)
    ocs = oracle.R1CS.from_csr(log_n, cs.mats)
    vp = oracle.PP.keygen_with(log_n, g, h, t)
    assert oracle.verify(ocs, vp, cs.v, proof) == 1
    w_bad = cs.w.copy(); w_bad[7] = oracle.fr_rand(3, 1)[0]
    assert oracle.verify(ocs, vp, cs.v, sb.MLArgumentForR1CS.prove(pk, cs.v, w_bad, pp)) < 0
    # the commitment scheme alone (commitment::mod.rs commit_open_verify_bench)
    z = np.concatenate([cs.v, cs.w]); point = oracle.fr_rand(4, log_n)
    _, com = sb.MLPolyCommit.commit(pp, z)
    ev, (_, proofs) = sb.MLPolyCommit.open(pp, z, point)
    assert oracle.pc_verify(vp, com, point, ev, proofs)


def test_resident_witness_gives_the_same_proof(sb, oracle, gpu_ctx):
    # bench.py's device-resident arm (sb_prove_resident) against the host-buffer call (sb_prove)
    log_n = 9
    cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 31)
    g, h = oracle.generators()
    pp = sb.MLPolyCommit.keygen(log_n, g, h, oracle.fr_rand(32, log_n), ctx=gpu_ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=gpu_ctx)
    wit = sb.Witness(pk, cs.v, cs.w)
    p_host = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp)
    p_res, phases = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases")
    assert p_host == p_res and set(phases) >= {"prove1_commit", "sumcheck1", "total"}
    l0 = gpu_ctx.launch_count(); h0, d0 = gpu_ctx.copy_counters()
    sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp)
    h1, d1 = gpu_ctx.copy_counters()
    assert gpu_ctx.launch_count() > l0                      # the CUDA path ran
    assert h1 - h0 >= 32 << log_n and d1 - d0 > 0           # v, w went in; messages came out


def test_adversarial_scalars_do_not_serialise(sb, oracle, gpu_ctx):
    # every scalar equal: all digits of a window land in one bucket; the chunked accumulation must stay exact
    g, _ = oracle.generators()
    n = 512
    ks = oracle.fr_rand(91, n)
    bases = np.stack([oracle.g1_mul(g, ks[i]) for i in range(n)])
    s = np.repeat(oracle.fr_rand(92, 1), n, axis=0)
    assert np.array_equal(sb.multi_scalar_mul(1, bases, s, ctx=gpu_ctx), oracle.msm_g1(bases, s))


def test_wire_formats_carry_keys_and_proofs(sb, oracle, gpu_ctx):
    # keys made on the GPU are written in the reference's key-cache layout
    # (commitment/mod.rs:48-62, uncompressed), read back, uploaded with sb_pp_load, and must give the same proof;
    # the proof itself parses into the reference's Proof fields and re-serializes to the same bytes.
    from r1cs_spartan_b200 import wire
    log_n = 6
    cs = sb.SyntheticR1CS(8, (1 << log_n) - 8, 5, 4242)
    g, h = oracle.generators()
    pp = sb.MLPolyCommit.keygen(log_n, g, h, oracle.fr_rand(17, log_n), keep_all_levels=True, ctx=gpu_ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=gpu_ctx)
    proof = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp)
    pg = [pp.export(1, i) for i in range(log_n)]; ph = [pp.export(2, i) for i in range(log_n)]
    cache = wire.key_cache_to_bytes([(dict(nv=log_n, powers_of_g=pg, powers_of_h=ph, g=g, h=h),
                                      dict(nv=log_n, g=g, h=h, g_mask_random=pp.g_mask_random()))])
    (kp, kv), = wire.key_cache_from_bytes(cache, check=True)
    pp2 = sb.MLPolyCommit.load(log_n, kp["powers_of_g"][0], kp["powers_of_h"], kp["h"], ctx=gpu_ctx)
    mats, ln = wire.index_from_bytes(wire.index_to_bytes(cs.mats, log_n))
    pk2 = sb.MLArgumentForR1CS.index(*mats, ctx=gpu_ctx)
    assert sb.MLArgumentForR1CS.prove(pk2, cs.v, cs.w, pp2) == proof
    parsed = wire.Proof.from_bytes(proof)
    assert parsed.to_bytes() == proof and parsed.commitment_nv == log_n
    _, com = sb.MLPolyCommit.commit(pp, np.concatenate([cs.v, cs.w]))
    assert np.array_equal(parsed.commitment, com)
    vp = oracle.PP.verifier_only(log_n, kv["g"], kv["h"], kv["g_mask_random"])
    assert oracle.verify(oracle.R1CS.from_csr(log_n, mats), vp, cs.v, proof) == 1


def _assert_trace_equal(tr, otr):
    """every intermediate the reference's prover produces, earliest stage first (BASELINE.json configs 2 and 3)"""
    assert np.array_equal(tr.commitment, np.frombuffer(otr.blob("commitment"), dtype=np.uint64)), "commitment"
    assert np.array_equal(tr.r_v, otr.fr("r_v")), "r_v"
    assert np.array_equal(tr.z_rv_0, otr.fr("z_rv_0")[0]), "z(r_v, 0)"
    assert np.array_equal(tr.open1_proofs.reshape(-1), np.frombuffer(otr.blob("open1_proofs"), dtype=np.uint64)), "first opening"
    assert np.array_equal(tr.tor, otr.fr("tor")), "tor"
    assert np.array_equal(tr.az, otr.fr("az")), "Az"
    assert np.array_equal(tr.bz, otr.fr("bz")), "Bz"
    assert np.array_equal(tr.cz, otr.fr("cz")), "Cz"
    assert np.array_equal(tr.sc1_evals.reshape(-1, 4), otr.fr("sc1_evals")), "first sumcheck messages"
    assert np.array_equal(tr.r_x, otr.fr("r_x")), "r_x"
    assert np.array_equal(tr.vabc, otr.fr("vabc")), "va vb vc"
    assert np.array_equal(tr.r_abc, otr.fr("r_abc")), "r_a r_b r_c"
    assert np.array_equal(tr.sc2_evals.reshape(-1, 4), otr.fr("sc2_evals")), "second sumcheck messages"
    assert np.array_equal(tr.r_y, otr.fr("r_y")), "r_y"
    assert np.array_equal(tr.z_ry, otr.fr("z_ry")[0]), "z(r_y)"
    assert np.array_equal(tr.open2_proofs.reshape(-1), np.frombuffer(otr.blob("open2_proofs"), dtype=np.uint64)), "second opening"


@pytest.mark.parametrize("log_n", [16, 20])
def test_full_size_prove_is_bit_exact_and_accepted(sb, oracle, gpu_ctx, log_n):
    # BASELINE.json configs 2 (2^16) and 3 (2^20, the north-star size): the benchmark circuit of benchmark.rs:63-65
    # proved on the GPU and by the literal CPU restatement on THE SAME public parameters, compared bit for bit on
    # Az/Bz/Cz, every message of both sumchecks, va/vb/vc, the commitment, both openings, every challenge and the
    # proof bytes.  The oracle is given the GPU-made parameters (the GPU keygen itself is checked level by level
    # against the oracle's literal setup.rs in test_keygen_matches_oracle) and runs its heavy loops on all host
    # cores -- test-only threading, the bytes do not depend on it (tests/test_cpu_oracle.py checks that).
    import os
    cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)
    g, h = oracle.generators()
    t = oracle.fr_rand(2024 + log_n, log_n)
    pp = sb.MLPolyCommit.keygen(log_n, g, h, t, keep_all_levels=True, ctx=gpu_ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=gpu_ctx)
    proof, tr = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp, trace=True)
    assert len(proof) == sb.load_library().sb_proof_size(log_n)
    ocs = oracle.R1CS.from_csr(log_n, cs.mats)
    # prove -> verify (benchmark.rs:35-47) with the CPU pairing verifier on the g^{t_i} of the GPU keygen
    vp = oracle.PP.verifier_only(log_n, g, h, pp.g_mask_random())
    assert oracle.verify(ocs, vp, cs.v, proof) == 1
    if log_n <= 16:
        w_bad = cs.w.copy(); w_bad[12345 % len(w_bad)] = oracle.fr_rand(1, 1)[0]
        assert oracle.verify(ocs, vp, cs.v, sb.MLArgumentForR1CS.prove(pk, cs.v, w_bad, pp)) < 0
    # the literal prover on the same parameters
    opp = oracle.PP.from_arrays(log_n, pp.export(1, 0), np.concatenate([pp.export(2, i) for i in range(log_n)], axis=0), h)
    old = oracle.set_threads(int(os.environ.get("SB_ORACLE_THREADS", os.cpu_count() or 1)))
    try:
        oproof, otr = oracle.prove(ocs, opp, cs.v, cs.w)
    finally:
        oracle.set_threads(old)
    _assert_trace_equal(tr, otr)
    assert proof == oproof
    # the witness kept in HBM (bench.py's `value` arm) gives the same bytes
    wit = sb.Witness(pk, cs.v, cs.w)
    assert sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit) == proof
    # BASELINE.json config 4 on one GPU: the commitment scheme alone on a random table of the same size
    # (commitment/mod.rs:66-83 commit_open_verify_bench): commit and open equal the oracle's, the pairing check
    # accepts, a wrong value fails
    table = oracle.fr_rand(5 + log_n, 1 << log_n); point = oracle.fr_rand(6 + log_n, log_n)
    _, com = sb.MLPolyCommit.commit(pp, table)
    ev, (_, proofs) = sb.MLPolyCommit.open(pp, table, point)
    assert oracle.pc_verify(vp, com, point, ev, proofs)
    assert not oracle.pc_verify(vp, com, point, oracle.fr_rand(7, 1)[0], proofs)
    if log_n <= 16:
        old = oracle.set_threads(int(os.environ.get("SB_ORACLE_THREADS", os.cpu_count() or 1)))
        try:
            assert np.array_equal(com, opp.commit(table))
            oev, oproofs = opp.open(table, point)
        finally:
            oracle.set_threads(old)
        assert np.array_equal(ev, oev) and np.array_equal(proofs, oproofs)


def test_msm_knobs_forced_on_small_inputs():
    # The batched-affine pairwise rounds only switch on for groups of >= 2^21 (window, point) pairs and the deep
    # accumulation levels only appear with long bucket runs; force both on the small parity cases (tiny chunk
    # sizes = many levels; 1..6 rounds, shares of 1..256 additions per inversion), in a fresh process because the
    # knobs are read once.
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ({"SB_MSM_AFFINE_LOG2": "5"}, {"SB_MSM_AFFINE_LOG2": "3", "SB_MSM_S0": "3", "SB_MSM_S1": "2", "SB_MSM_AFFINE_ROUNDS": "2", "SB_MSM_AFFINE_K": "3"},
                  {"SB_MSM_AFFINE_LOG2": "0", "SB_MSM_AFFINE_ROUNDS": "6", "SB_MSM_AFFINE_K": "1"}, {"SB_MSM_AFFINE_ROUNDS": "0"},
                  {"SB_MSM_S0": "2", "SB_MSM_S1": "2", "SB_MSM_RED_L": "8", "SB_MSM_ORDER": "0", "SB_MSM_SORTED": "0"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                            "-k", "msm_matches or structured or commit_and_open or prove_bytes or adversarial"],
                           env=env, cwd=root, capture_output=True, text=True, timeout=1200)
        assert r.returncode == 0, (extra, r.stdout[-3000:], r.stderr[-2000:])
