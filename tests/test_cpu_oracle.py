"""CPU tests (-m "not gpu"): the C++ oracle against the golden vectors and the Python big-int model,
the reference's own algebraic identities on the oracle, the generated PTX streams under emulation,
the workload generator, and the C-ABI library's export table.  No compute call touches a GPU here."""
import hashlib
import json
import os
import random
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
FR_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _sha_fr(oracle, arr):
    h = hashlib.sha256()
    for v in oracle.fr_to_ints(arr):
        h.update(v.to_bytes(32, "little"))
    return h.hexdigest()


# ---------------------------------------------------------------- golden vectors
def test_oracle_known_answers(oracle):
    k = GOLDEN["kat"]
    fs = oracle.FsRng()
    fs.feed(b"r1cs-spartan golden"); fs.feed(bytes(range(200)))
    assert fs.fill(77).hex() == k["fs_fill_77"]
    fs.feed(b"\x00" * 64)
    assert [hex(v) for v in oracle.fr_to_ints(fs.fr_rand(5))] == k["fs_fr_rand_5"]
    assert [hex(v) for v in oracle.fr_to_ints(oracle.fr_rand(7, 4))] == k["splitmix_fr_rand_seed7"]
    g, h = oracle.generators()
    assert [hex(c) for c in oracle.g1_to_py(g)] == k["g1_generator"]
    assert oracle.ser_g1(g).hex() == k["ser_g1_gen"] and oracle.ser_g2(h).hex() == k["ser_g2_gen"]
    assert oracle.ser_g1(np.zeros(12, dtype=np.uint64)).hex() == k["ser_g1_inf"]
    sc = oracle.fr_from_ints([int(k["scalar"], 16)])[0]
    assert oracle.ser_g1(oracle.g1_mul(g, sc)).hex() == k["ser_g1_k_gen"]
    assert oracle.ser_g2(oracle.g2_mul(h, sc)).hex() == k["ser_g2_k_gen"]
    minus1 = oracle.fr_from_ints([FR_MOD - 1])[0]
    assert oracle.ser_g1(oracle.g1_mul(g, minus1)).hex() == k["ser_g1_neg_gen"]
    assert oracle.ser_g2(oracle.g2_mul(h, minus1)).hex() == k["ser_g2_neg_gen"]
    one = oracle.fr_from_ints([1])[0]
    assert hex(oracle.limbs_to_ints(one.reshape(1, 4))[0]) == k["fr_montgomery_one"]


@pytest.mark.parametrize("case", GOLDEN["prove"], ids=lambda c: "l%d" % c["log_n"])
def test_oracle_prove_matches_golden(oracle, case):
    import r1cs_spartan_b200 as sb
    log_n = case["log_n"]
    cs = sb.SyntheticR1CS(case["num_public"], (1 << log_n) - case["num_public"], case["density"], case["seed"])
    assert cs.nnz == case["nnz"]
    ocs = oracle.R1CS.from_csr(log_n, cs.mats)
    pp = oracle.PP.keygen(log_n, case["trapdoor_seed"])
    proof, tr = oracle.prove(ocs, pp, cs.v, cs.w)
    assert len(proof) == case["proof_len"]
    for k in ("az", "bz", "cz", "r_v", "tor", "r_x", "r_y"):
        assert _sha_fr(oracle, tr.fr(k)) == case["sha256"][k], k
    assert [hex(v) for v in oracle.fr_to_ints(tr.fr("vabc"))] == case["va_vb_vc"]
    assert oracle.ser_g1(np.frombuffer(tr.blob("commitment"), dtype=np.uint64)).hex() == case["commitment_compressed_hex"]
    assert proof.hex() == case["proof_hex"]


# ---------------------------------------------------------------- oracle vs the python model on fresh inputs
def test_oracle_field_ops_vs_bigint(oracle):
    from oracle import pymodel as pm
    rnd = random.Random(3)
    for mod, to_m, from_m, binop in ((pm.R, oracle.fr_from_ints, oracle.fr_to_ints, oracle.fr_binop),
                                     (pm.P, oracle.fq_from_ints, oracle.fq_to_ints, oracle.fq_binop)):
        a = [0, 1, mod - 1] + [rnd.randrange(mod) for _ in range(500)]
        b = [mod - 1, mod - 1, mod - 1] + [rnd.randrange(mod) for _ in range(500)]
        A, B = to_m(a), to_m(b)
        assert from_m(binop("mul", A, B)) == [x * y % mod for x, y in zip(a, b)]
        assert from_m(binop("add", A, B)) == [(x + y) % mod for x, y in zip(a, b)]
        assert from_m(binop("sub", A, B)) == [(x - y) % mod for x, y in zip(a, b)]


def test_oracle_workload_equals_product_generator(oracle):
    # the oracle restates constraints.rs independently of r1cs-spartan_b200/workload.py
    import r1cs_spartan_b200 as sb
    for (npub, log_n, dens) in [(4, 4, 0), (32, 8, 0), (32, 9, 100), (8, 6, 255)]:
        cs = sb.SyntheticR1CS(npub, (1 << log_n) - npub, dens, 0x5EED0000 + log_n)
        oc = oracle.R1CS.synth(npub, (1 << log_n) - npub, dens, 0x5EED0000 + log_n)
        ov, ow = oc.vw()
        assert np.array_equal(cs.v, ov) and np.array_equal(cs.w, ow)
        for k in range(3):
            rp, col, val = oc.csr(k)
            assert np.array_equal(rp, cs.mats[k][0]) and np.array_equal(col, cs.mats[k][1]) and np.array_equal(val, cs.mats[k][2])
        assert oc.is_satisfied(np.concatenate([cs.v, cs.w]))
        assert cs.n == 1 << log_n


# ---------------------------------------------------------------- the reference's own identities, on the oracle
def test_eq_extension_indicator(oracle):            # eq.rs:30-46
    tbits = 0b101101001
    t = oracle.fr_from_ints([(tbits >> i) & 1 for i in range(9)])
    tabs = oracle.eq_extension(t)
    prod = tabs[0]
    for i in range(1, 9):
        prod = oracle.fr_binop("mul", prod, tabs[i])
    ints = oracle.fr_to_ints(prod)
    assert ints[tbits] == 1 and sum(ints) == 1


def test_commit_is_g_to_f_of_trapdoor(oracle):      # commit.rs:54-66
    nv = 4
    pp = oracle.PP.keygen(nv, 5)
    z = oracle.fr_rand(6, 1 << nv)
    g, _ = pp.gh()
    assert np.array_equal(pp.commit(z), oracle.g1_mul(g, oracle.mle_eval(z, pp.trapdoor())))


def test_open_quotient_identity(oracle):            # verify.rs:61-95 (the field part and pi_i = h * q_i(s[i..]))
    nv = 5
    pp = oracle.PP.keygen(nv, 8)
    s = pp.trapdoor()
    z = oracle.fr_rand(9, 1 << nv)
    point = oracle.fr_rand(10, nv)
    ev, proofs, q = pp.open(z, point, want_q=True)
    _, h = pp.gh()
    fx = oracle.fr_to_ints(oracle.mle_eval(z, s).reshape(1, 4))[0]
    ft = oracle.fr_to_ints(ev.reshape(1, 4))[0]
    si = oracle.fr_to_ints(s); pi = oracle.fr_to_ints(point)
    rhs, off = 0, 0
    for i in range(nv):
        k = nv - i
        qk = q[off:off + (1 << (k - 1))]; off += 1 << (k - 1)
        q_i = np.repeat(qk, 2, axis=0)               # q_i[a] = q[k][a >> 1]
        qv = oracle.mle_eval(q_i, s[i:])
        assert np.array_equal(oracle.g2_mul(h, qv), proofs[i]), "open error"
        rhs = (rhs + (si[i] - pi[i]) * oracle.fr_to_ints(qv.reshape(1, 4))[0]) % FR_MOD
    assert (fx - ft) % FR_MOD == rhs
    # D3 (halved MSMs) rests on powers_of_h[i][2x] + powers_of_h[i][2x+1] == powers_of_h[i+1][x]
    for i in range(nv - 1):
        lo, hi = pp.g2(i), pp.g2(i + 1)
        for x in range(hi.shape[0]):
            assert np.array_equal(oracle.g2_add(lo[2 * x], lo[2 * x + 1]), hi[x])
    last = pp.g2(nv - 1)
    assert np.array_equal(oracle.g2_add(last[0], last[1]), h)


def test_msm_pippenger_vs_double_and_add(oracle):
    from oracle import pymodel as pm
    g, h = oracle.generators()
    n = 40
    ks = oracle.fr_rand(1, n); s = oracle.fr_rand(2, n)
    b1 = np.stack([oracle.g1_mul(g, ks[i]) for i in range(n)])
    got = oracle.g1_to_py(oracle.msm_g1(b1, s))
    exp = pm.msm(pm.FQ, [oracle.g1_to_py(x) for x in b1], oracle.fr_to_ints(s))
    assert got == exp


# ---------------------------------------------------------------- generated PTX streams, emulated
def test_generated_field_streams_under_emulation():
    sys.path.insert(0, os.path.join(ROOT, "r1cs-spartan_b200", "tools"))
    import gen_field as g
    rnd = random.Random(11)
    for fld in ("fr", "fq"):
        f = g.FIELDS[fld]; n, p = f["N"], f["p"]
        rinv = pow(1 << (32 * n), -1, p)
        cases = [(0, 0), (p - 1, p - 1), (1, p - 1), (0, 1)] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(300)]
        for a, b in cases:
            assert g.run_emulated("mul", fld, a, b) == a * b * rinv % p
            assert g.run_emulated("add", fld, a, b) == (a + b) % p
            assert g.run_emulated("sub", fld, a, b) == (a - b) % p


def test_lazy_reduction_sequences_under_emulation():
    # The Fq2 product (Fq2::mul, device path) is a
    # composition of generated streams (wide product, wide add/sub, Montgomery reduction).  Execute the same
    # compositions, stream by stream, on the emulator and compare with big-integer arithmetic; also check the
    # range facts the comments in field.cuh rely on (inputs of every reduction below p * 2^384, no 768-bit wrap).
    sys.path.insert(0, os.path.join(ROOT, "r1cs-spartan_b200", "tools"))
    import gen_field as g
    f = g.FIELDS["fq"]; p = f["p"]; R = 1 << 384; rinv = pow(R, -1, p)
    W = 1 << 768
    mw = lambda a, b: g.run_emulated_wide("fq", a, b)
    op = lambda kind, a, b: g.run_emulated_wide_op("fq", kind, a, b)

    def redc(t):
        assert 0 <= t < p * R, "reduction input out of range"
        return g.run_emulated_redc("fq", t)

    def fq2_mul(a, b):                                    # Fq2::mul, __CUDA_ARCH__ branch
        t0, t1 = mw(a[0], b[0]), mw(a[1], b[1])
        t2 = mw(op("addn", a[0], a[1]), op("addn", b[0], b[1]))
        u = op("sub", t2, t0); im = op("sub", u, t1)
        re = op("subp2", t0, t1)
        return redc(re), redc(im)

    def ref_mul(a, b):                                    # Montgomery-domain Fq2 product
        return ((a[0] * b[0] - a[1] * b[1]) * rinv % p, (a[0] * b[1] + a[1] * b[0]) * rinv % p)

    rnd = random.Random(5)
    edge = [0, 1, p - 1, p - 2, (p - 1) // 2]
    pick = lambda: rnd.choice(edge) if rnd.random() < 0.3 else rnd.randrange(p)
    for _ in range(120):
        a, b, c, d = [(pick(), pick()) for _ in range(4)]
        assert fq2_mul(a, b) == ref_mul(a, b)
        assert fq2_mul(c, d) == ref_mul(c, d)
    # extremes of the ranges
    hi = (p - 1, p - 1); lo = (0, 0)
    for a, b, c, d in ((hi, hi, hi, hi), (hi, hi, lo, lo), (lo, lo, hi, hi), ((p - 1, 0), (p - 1, 0), (0, p - 1), (0, p - 1)),
                       ((0, p - 1), (0, p - 1), (p - 1, 0), (p - 1, 0))):
        assert fq2_mul(a, b) == ref_mul(a, b)
        assert fq2_mul(c, d) == ref_mul(c, d)


def test_generated_header_is_up_to_date(tmp_path):
    gen = os.path.join(ROOT, "r1cs-spartan_b200", "tools", "gen_field.py")
    out = tmp_path / "fp_gen.cuh"
    subprocess.check_call([sys.executable, gen, str(out)])
    assert out.read_text() == open(os.path.join(ROOT, "r1cs-spartan_b200", "csrc", "fp_gen.cuh")).read()


# ---------------------------------------------------------------- C ABI surface
def test_library_exports_every_declared_symbol():
    import re
    import r1cs_spartan_b200 as sb
    if not os.path.exists(sb.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = sb.load_library()
    header = open(os.path.join(ROOT, "include", "spartan_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libspartan_b200.so does not export %s" % name
    assert declared == set(sb.EXPORTS)
    assert lib.sb_proof_size(20) == 21272


def test_no_cpu_fallback_without_gpu():
    import torch
    import r1cs_spartan_b200 as sb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sb.CudaError):
        sb.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "r1cs-spartan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in src.replace("oracle/spartan_oracle.cpp synth_r1cs", "") or f == "workload.py", f
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


# ---------------------------------------------------------------- verifier (pairing) on the oracle
def test_pairing_is_bilinear_and_nondegenerate(oracle):
    assert oracle.pairing_check(oracle.fr_rand(1, 1)[0], oracle.fr_rand(2, 1)[0])
    assert oracle.pairing_check(oracle.fr_from_ints([1])[0], oracle.fr_from_ints([FR_MOD - 1])[0])


def test_point_decompression_round_trip(oracle):
    g, h = oracle.generators()
    for k in oracle.fr_rand(3, 6):
        p1, p2 = oracle.g1_mul(g, k), oracle.g2_mul(h, k)
        assert np.array_equal(oracle.deser_g1(oracle.ser_g1(p1)), p1)
        assert np.array_equal(oracle.deser_g2(oracle.ser_g2(p2)), p2)
    assert not oracle.deser_g1(oracle.ser_g1(np.zeros(12, dtype=np.uint64))).any()


def test_commit_open_verify(oracle):
    # reference: commitment::verify::sanity (verify.rs:61-95), nv = 6
    nv = 6
    pp = oracle.PP.keygen(nv, 21)
    z = oracle.fr_rand(22, 1 << nv)
    point = oracle.fr_rand(23, nv)
    com = pp.commit(z)
    ev, proofs = pp.open(z, point)
    assert oracle.pc_verify(pp, com, point, ev, proofs)
    assert not oracle.pc_verify(pp, com, point, oracle.fr_rand(24, 1)[0], proofs)        # wrong evaluation
    bad = proofs.copy(); bad[[0, 1]] = bad[[1, 0]]
    assert not oracle.pc_verify(pp, com, point, ev, bad)                                  # proofs out of order


def test_prove_then_verify_accepts_and_rejects(oracle):
    # reference: ahp::tests::test_small (ahp/tests.rs:73-75: log_n = 8, log_v = 2, density 1) through the
    # non-interactive driver (benchmark.rs:35-47: prove -> serialize -> deserialize -> verify)
    log_n, log_v = 8, 2
    cs = oracle.R1CS.synth(1 << log_v, (1 << log_n) - (1 << log_v), 1, 12345)
    v, w = cs.vw()
    pp = oracle.PP.keygen(log_n, 54321)
    proof, _ = oracle.prove(cs, pp, v, w)
    assert oracle.verify(cs, pp, v, proof) == 1
    w_bad = w.copy(); w_bad[5] = oracle.fr_rand(9, 1)[0]
    bad_proof, _ = oracle.prove(cs, pp, v, w_bad)
    assert oracle.verify(cs, pp, v, bad_proof) < 0                # WrongWitness
    v_bad = v.copy(); v_bad[1] = oracle.fr_rand(10, 1)[0]
    assert oracle.verify(cs, pp, v_bad, proof) < 0                # public input does not match the commitment
    assert oracle.verify(cs, pp, v, proof[:-1]) == 0              # truncated proof
    flipped = bytearray(proof); flipped[8 + 48 + 5] ^= 1          # z_rv_0
    assert oracle.verify(cs, pp, v, bytes(flipped)) <= 0


def test_oracle_threading_does_not_change_a_byte(oracle):
    # the full-size GPU parity tests (2^16, 2^20) run the literal prover's heavy loops on all host cores
    # (oracle.set_threads, test-only); exact arithmetic makes the result independent of the thread count --
    # checked here on proof bytes, every traced intermediate, keygen, commit and open
    import r1cs_spartan_b200 as sb
    log_n = 11
    cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)
    ocs = oracle.R1CS.from_csr(log_n, cs.mats)
    g, h = oracle.generators()
    t = oracle.fr_rand(77, log_n)
    assert oracle.set_threads(1) == 1          # the default is the reference's configuration: one thread
    pp1 = oracle.PP.keygen_with(log_n, g, h, t)
    p1, tr1 = oracle.prove(ocs, pp1, cs.v, cs.w)
    table = oracle.fr_rand(5, 1 << log_n); point = oracle.fr_rand(6, log_n)
    c1, o1 = pp1.commit(table), pp1.open(table, point)
    oracle.set_threads(5)
    try:
        pp5 = oracle.PP.keygen_with(log_n, g, h, t)
        p5, tr5 = oracle.prove(ocs, pp5, cs.v, cs.w)
        c5, o5 = pp5.commit(table), pp5.open(table, point)
    finally:
        oracle.set_threads(1)
    assert p1 == p5
    for k in ("az", "bz", "cz", "sc1_evals", "sc2_evals", "vabc", "m_comb", "commitment", "open1_proofs", "open2_proofs", "r_x", "r_y"):
        assert tr1.blob(k) == tr5.blob(k), k
    assert np.array_equal(pp1.g2_all(), pp5.g2_all()) and np.array_equal(pp1.g1(0), pp5.g1(0)) and np.array_equal(pp1.g_mask(), pp5.g_mask())
    assert np.array_equal(c1, c5) and np.array_equal(o1[0], o5[0]) and np.array_equal(o1[1], o5[1])
