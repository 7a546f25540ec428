"""Wire formats (r1cs-spartan_b200/wire.py) against the oracle's independent serializers and parsers.

CPU only.  The oracle (C++ restatement + Python big-int model) is the checker: the product module never
imports it.  What is pinned: the byte layouts of Proof, IndexPK/IndexVK, PublicParameter, VerifierParameter
and the key cache as the reference's derive(CanonicalSerialize) structs define them (file:line in wire.py).
"""
import json
import os

import numpy as np
import pytest

from r1cs_spartan_b200 import wire
from r1cs_spartan_b200.workload import SyntheticR1CS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def small(oracle):
    log_n = 5
    cs = oracle.R1CS.synth(4, (1 << log_n) - 4, 1, 77)
    pp = oracle.PP.keygen(log_n, 5)
    v, w = cs.vw()
    proof, trace = oracle.prove(cs, pp, v, w)
    return dict(log_n=log_n, cs=cs, pp=pp, v=v, w=w, proof=bytes(proof), trace=trace)


def test_field_and_point_encodings_match_the_oracle(oracle):
    g, h = oracle.generators()
    ks = oracle.fr_rand(11, 6)
    for k in ks:
        p1, p2 = oracle.g1_mul(g, k), oracle.g2_mul(h, k)
        b1, b2 = wire.g1_to_bytes(p1), wire.g2_to_bytes(p2)
        assert b1 == oracle.ser_g1(p1) and b2 == oracle.ser_g2(p2)
        assert np.array_equal(wire.g1_from_bytes(b1), p1) and np.array_equal(wire.g2_from_bytes(b2), p2)
        assert np.array_equal(wire.g1_from_bytes(b1), oracle.deser_g1(b1))
        assert np.array_equal(wire.g2_from_bytes(b2), oracle.deser_g2(b2))
        u1, u2 = wire.g1_to_bytes(p1, compressed=False), wire.g2_to_bytes(p2, compressed=False)
        assert len(u1) == 96 and len(u2) == 192 and u1[:47] == b1[:47] and u2[:95] == b2[:95]
        assert np.array_equal(wire.g1_from_bytes(u1, compressed=False), p1)
        assert np.array_equal(wire.g2_from_bytes(u2, compressed=False), p2)
    # infinity, both forms
    z1, z2 = np.zeros(12, np.uint64), np.zeros(24, np.uint64)
    assert wire.g1_to_bytes(z1) == oracle.ser_g1(z1) and wire.g2_to_bytes(z2) == oracle.ser_g2(z2)
    for comp in (True, False):
        assert not wire.g1_from_bytes(wire.g1_to_bytes(z1, comp), comp).any()
        assert not wire.g2_from_bytes(wire.g2_to_bytes(z2, comp), comp).any()
    # Fr: canonical little-endian
    fr = oracle.fr_rand(3, 5)
    ints = oracle.fr_to_ints(fr)
    assert wire.fr_array_to_bytes(fr) == b"".join(int(v).to_bytes(32, "little") for v in ints)
    assert np.array_equal(wire.fr_array_from_bytes(wire.fr_array_to_bytes(fr), 5), fr)


def test_golden_vectors():
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["kat"]
    from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
    assert wire.g1_to_bytes(G1_GENERATOR).hex() == kat["ser_g1_gen"]
    assert wire.g2_to_bytes(G2_GENERATOR).hex() == kat["ser_g2_gen"]
    assert wire.g1_to_bytes(np.zeros(12, np.uint64)).hex() == kat["ser_g1_inf"]
    neg = wire.g1_from_bytes(bytes.fromhex(kat["ser_g1_neg_gen"]))
    assert np.array_equal(neg[:6], G1_GENERATOR[:6]) and not np.array_equal(neg[6:], G1_GENERATOR[6:])
    neg2 = wire.g2_from_bytes(bytes.fromhex(kat["ser_g2_neg_gen"]))
    assert np.array_equal(neg2[:12], G2_GENERATOR[:12]) and not np.array_equal(neg2[12:], G2_GENERATOR[12:])


def test_malformed_points_are_rejected():
    with pytest.raises(wire.SerializationError):
        wire.g1_from_bytes(b"\x00" * 47)                                   # truncated
    with pytest.raises(wire.SerializationError):
        wire.g1_from_bytes((wire.FQ_MOD).to_bytes(48, "little"))           # x not reduced
    bad = None
    for x in range(1, 50):                                                  # an x with no point above it
        if wire._fq_sqrt((x ** 3 + 4) % wire.FQ_MOD) is None:
            bad = x
            break
    with pytest.raises(wire.SerializationError):
        wire.g1_from_bytes(bad.to_bytes(48, "little"))
    with pytest.raises(wire.SerializationError):
        wire.g1_from_bytes((1).to_bytes(48, "little") + (1).to_bytes(48, "little"), compressed=False)


def test_proof_round_trip_and_fields(oracle, small):
    p = wire.Proof.from_bytes(small["proof"])
    assert p.to_bytes() == small["proof"]
    log_n, tr = small["log_n"], small["trace"]
    assert p.commitment_nv == log_n and len(p.first_sumcheck_messages) == log_n and len(p.second_sumcheck_messages) == log_n
    assert all(len(m) == log_n + 3 for m in p.first_sumcheck_messages)     # max_multiplicands + 1 evaluations
    assert all(len(m) == 3 for m in p.second_sumcheck_messages)
    assert sorted(p.third_index_info) == sorted((log_n + 2, log_n)) and sorted(p.fifth_index_info) == sorted((2, log_n))
    assert p.proof_for_z_rv_0.proofs.shape == (log_n, 24) and p.proof_for_z_ry.proofs.shape == (log_n, 24)
    g, h = small["pp"].gh()
    assert np.array_equal(p.proof_for_z_rv_0.h, h) and np.array_equal(p.proof_for_z_ry.h, h)
    assert np.array_equal(tr.fr("vabc"), np.stack([p.va, p.vb, p.vc]))
    assert np.array_equal(tr.fr("z_rv_0")[0], p.z_rv_0) and np.array_equal(tr.fr("z_ry")[0], p.z_ry)
    # the two openings inside the proof verify against the commitment inside the proof
    r_v = tr.fr("r_v").reshape(-1, 4); r_y = tr.fr("r_y").reshape(-1, 4)
    point0 = np.concatenate([r_v, np.zeros((log_n - len(r_v), 4), np.uint64)])
    assert oracle.pc_verify(small["pp"], p.commitment, point0, p.z_rv_0, p.proof_for_z_rv_0.proofs)
    assert oracle.pc_verify(small["pp"], p.commitment, r_y, p.z_ry, p.proof_for_z_ry.proofs)
    # truncation and trailing bytes are errors, not silent acceptance
    with pytest.raises(wire.SerializationError):
        wire.Proof.from_bytes(small["proof"][:-1])
    with pytest.raises(wire.SerializationError):
        wire.Proof.from_bytes(small["proof"] + b"\x00")


def test_index_layout_matches_the_transcript_bytes(oracle):
    # lib.rs:62-64 feeds the three serialized matrices to the transcript; the Python model's ser_matrix is that
    # byte string.  IndexPK = the three matrices then log_n (indexer.rs:11-17).
    from oracle import pymodel as pm
    log_n = 4
    cs = SyntheticR1CS(4, (1 << log_n) - 4, 3, 99)
    blob = wire.index_to_bytes(cs.mats, log_n)
    expect = b""
    for (row_ptr, col, val) in cs.mats:
        ints = oracle.fr_to_ints(val)
        rows = [[(int(ints[e]), int(col[e])) for e in range(int(row_ptr[i]), int(row_ptr[i + 1]))] for i in range(1 << log_n)]
        expect += pm.ser_matrix(rows, 1 << log_n)
    assert blob == expect + (log_n).to_bytes(8, "little")
    mats, ln = wire.index_from_bytes(blob)
    assert ln == log_n
    for a, b in zip(mats, cs.mats):
        assert np.array_equal(a[0], np.asarray(b[0], np.uint64)) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    with pytest.raises(wire.SerializationError):
        wire.index_from_bytes(blob[:-8] + (log_n + 1).to_bytes(8, "little"))


@pytest.mark.parametrize("compressed", [True, False])
def test_commitment_keys_round_trip(oracle, compressed):
    nv = 3
    pp = oracle.PP.keygen(nv, 21)
    g, h = pp.gh()
    pg = [pp.g1(i) for i in range(nv)]; ph = [pp.g2(i) for i in range(nv)]
    blob = wire.public_parameter_to_bytes(nv, pg, ph, g, h, compressed)
    per = (48, 96) if compressed else (96, 192)
    npts = sum(1 << (nv - i) for i in range(nv))
    assert len(blob) == 8 + 8 + 8 * nv + per[0] * npts + 8 + 8 * nv + per[1] * npts + per[0] + per[1]
    back = wire.public_parameter_from_bytes(blob, compressed)
    assert back["nv"] == nv and np.array_equal(back["g"], g) and np.array_equal(back["h"], h)
    for i in range(nv):
        assert np.array_equal(back["powers_of_g"][i], pg[i]) and np.array_equal(back["powers_of_h"][i], ph[i])
    vblob = wire.verifier_parameter_to_bytes(nv, g, h, pp.g_mask(), compressed)
    vback = wire.verifier_parameter_from_bytes(vblob, compressed)
    assert np.array_equal(vback["g_mask_random"], pp.g_mask())
    # the arrays that come back drive the oracle's commitment scheme exactly like the originals
    pp2 = oracle.PP.from_arrays(nv, back["powers_of_g"][0], np.concatenate(back["powers_of_h"]), back["h"])
    z = oracle.fr_rand(8, 1 << nv)
    assert np.array_equal(pp2.commit(z), pp.commit(z))
    if not compressed:
        pairs = wire.key_cache_from_bytes(wire.key_cache_to_bytes([(back, vback), (back, vback)]))
        assert len(pairs) == 2 and np.array_equal(pairs[1][0]["powers_of_h"][1], ph[1])


def test_rust_shim_lists_every_export():
    """rust/src/ffi.rs declares exactly the symbols include/spartan_b200.h exports, and rust/build.rs compiles exactly the
    Makefile's source list (the Rust side cannot be compiled here: this keeps it from drifting)."""
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "spartan_b200.h")).read()
    ffi = open(os.path.join(root, "rust", "src", "ffi.rs")).read()
    c_syms = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", hdr))
    rs_syms = set(re.findall(r"pub fn (sb_[a-z0-9_]+)\s*\(", ffi))
    assert c_syms == rs_syms, (sorted(c_syms - rs_syms), sorted(rs_syms - c_syms))
    import r1cs_spartan_b200 as sb
    assert set(sb.EXPORTS) <= c_syms, sorted(set(sb.EXPORTS) - c_syms)
    mk = open(os.path.join(root, "r1cs-spartan_b200", "Makefile")).read()
    mk_src = set(re.findall(r"csrc/([a-z_]+\.cu)", re.search(r"^SRC := (.*)$", mk, re.M).group(1)))
    rs_src = set(re.findall(r'"([a-z_]+\.cu)"', open(os.path.join(root, "rust", "build.rs")).read()))
    assert mk_src == rs_src, (mk_src, rs_src)


def test_host_field_arithmetic_selftest():
    """the host's 64-bit Montgomery product equals the portable 32-bit loop, and the binary-GCD inversion equals the Fermat
    ladder, on random and edge-case operands of both fields (pure host code of the product library: no device needed)"""
    import ctypes as C
    import r1cs_spartan_b200 as sb
    L = sb.load_library()
    L.sb_selftest_host_field.restype = C.c_int
    assert L.sb_selftest_host_field(C.c_int(4000), C.c_uint64(12345)) == 0
