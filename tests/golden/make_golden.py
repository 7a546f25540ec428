#!/usr/bin/env python3
"""Generates tests/golden/*.json from the independent Python big-integer model (oracle/pymodel.py).

PARITY UNPINNED: the reference cannot be executed in this environment (Rust, un-vendored arkworks) and
ships no golden vectors, so these fixtures pin the C++ oracle and the CUDA path to the Python model,
i.e. to a second, independently written restatement -- not to arkworks output.  If a real arkworks
build ever becomes available, regenerate this file from it: every consumer only reads the JSON.

Usage: python tests/golden/make_golden.py        (takes a few seconds)
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pymodel as pm                      # noqa: E402
import r1cs_spartan_b200.workload as wl               # noqa: E402  (pure python, no GPU)

HERE = os.path.dirname(os.path.abspath(__file__))


def rows_py(mat):
    row_ptr, col, val = mat
    vals = [v * pm.R_MONT_INV % pm.R for v in wl.limbs_to_mont(val)] if len(col) else []
    return [[(vals[e], int(col[e])) for e in range(int(row_ptr[r]), int(row_ptr[r + 1]))] for r in range(len(row_ptr) - 1)]


def sha(ints, nbytes=32):
    h = hashlib.sha256()
    for v in ints:
        h.update(int(v).to_bytes(nbytes, "little"))
    return h.hexdigest()


def prove_case(log_n, num_public, density, seed, tseed):
    cs = wl.SyntheticR1CS(num_public, (1 << log_n) - num_public, density, seed)
    g, h = pm.derive_generators()
    rng = pm.SplitMix64(tseed)
    t = [pm.fr_rand(rng) for _ in range(log_n)]
    pp, _ = pm.keygen(log_n, g, h, t)
    v = [x * pm.R_MONT_INV % pm.R for x in wl.limbs_to_mont(cs.v)]
    w = [x * pm.R_MONT_INV % pm.R for x in wl.limbs_to_mont(cs.w)]
    trace = {}
    proof = pm.prove(rows_py(cs.mats[0]), rows_py(cs.mats[1]), rows_py(cs.mats[2]), v, w, pp, trace)
    return {
        "log_n": log_n, "num_public": num_public, "density": density, "seed": seed, "trapdoor_seed": tseed,
        "nnz": cs.nnz, "proof_hex": proof.hex(), "proof_len": len(proof),
        "sha256": {k: sha(trace[k]) for k in ("az", "bz", "cz", "r_v", "tor", "r_x", "r_y")},
        "va_vb_vc": [hex(trace[k]) for k in ("va", "vb", "vc")],
        "commitment_compressed_hex": pm.ser_g1(trace["com"]).hex(),
        "z_rv_0": hex(trace["z_rv_0"]), "z_ry": hex(trace["z_ry"]),
    }


def kat():
    out = {}
    # transcript: feed / squeeze / Fr::rand
    fs = pm.Blake2sRng()
    fs.feed(b"r1cs-spartan golden"); fs.feed(bytes(range(200)))
    out["fs_fill_77"] = fs.fill_bytes(77).hex()
    fs.feed(b"\x00" * 64)
    out["fs_fr_rand_5"] = [hex(pm.fr_rand(fs)) for _ in range(5)]
    rng = pm.SplitMix64(7)
    out["splitmix_fr_rand_seed7"] = [hex(pm.fr_rand(rng)) for _ in range(4)]
    g, h = pm.derive_generators()
    out["g1_generator"] = [hex(g[0]), hex(g[1])]
    out["g2_generator"] = [[hex(c) for c in h[0]], [hex(c) for c in h[1]]]
    out["ser_g1_gen"] = pm.ser_g1(g).hex(); out["ser_g1_neg_gen"] = pm.ser_g1(pm.pt_neg(pm.FQ, g)).hex()
    out["ser_g2_gen"] = pm.ser_g2(h).hex(); out["ser_g2_neg_gen"] = pm.ser_g2(pm.pt_neg(pm.FQ2, h)).hex()
    out["ser_g1_inf"] = pm.ser_g1(None).hex()
    k = 0x1234567890abcdef1234567890abcdef
    out["scalar"] = hex(k)
    out["ser_g1_k_gen"] = pm.ser_g1(pm.pt_mul(pm.FQ, g, k)).hex()
    out["ser_g2_k_gen"] = pm.ser_g2(pm.pt_mul(pm.FQ2, h, k)).hex()
    out["fr_montgomery_one"] = hex(pm.R_MONT)
    return out


if __name__ == "__main__":
    data = {"kat": kat(), "prove": [prove_case(3, 4, 0, 0x5EED0003, 99), prove_case(5, 8, 0, 0x5EED0005, 99),
                                    prove_case(6, 32, 60, 0x5EED0006, 7)]}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(data, f, indent=1)
    print("wrote golden.json:", [c["proof_len"] for c in data["prove"]])
