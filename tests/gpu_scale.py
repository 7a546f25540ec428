"""Scale check (run on the GPU box): prove at 2^LOG_N constraints on one GPU, check the proof with the CPU pairing
verifier (prove -> verify is the size-independent property), print timings and device memory.
    python tests/gpu_scale.py 22 [verify|noverify] [oracle]
The optional third argument builds the circuit with the oracle's C++ generator (same circuit family, 4x faster than
the Python one) -- the oracle stays test infrastructure: it makes the input and checks the output, nothing else.
"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import r1cs_spartan_b200 as sb
from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
from r1cs_spartan_b200 import workload as wl

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
verify = len(sys.argv) <= 2 or sys.argv[2] != "noverify"
t0 = time.time()
ocs = None
if len(sys.argv) > 3 and sys.argv[3] == "oracle":
    from oracle import binding as ob
    ob.build(); ob.lib()
    ocs = ob.R1CS.synth(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)

    class _CS:
        pass
    cs = _CS(); cs.mats = [ocs.csr(k) for k in range(3)]; cs.v, cs.w = ocs.vw(); cs.nnz = [int(m[0][-1]) for m in cs.mats]
else:
    cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)
print("workload %.1fs nnz %s" % (time.time() - t0, cs.nnz), flush=True)
ctx = sb.Context(0)
trap = np.stack([wl.mont_to_limbs([wl.fr_rand_mont(wl.SplitMix64(99 + i))])[0] for i in range(log_n)])
t0 = time.time(); pp = sb.MLPolyCommit.keygen(log_n, G1_GENERATOR, G2_GENERATOR, trap, ctx=ctx); print("keygen %.1fs" % (time.time() - t0), flush=True)
t0 = time.time(); pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx); print("index %.2fs (device-side plans incl. upload %.0f ms, transcript hash added %.0f ms)" % ((time.time() - t0,) + pk.timing()), flush=True)
wit = sb.Witness(pk, cs.v, cs.w)
for i in range(4):
    t0 = time.time()
    proof, ph = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases")
    print("prove %.1f ms" % (1e3 * (time.time() - t0)), {k: round(v, 1) for k, v in ph.items() if v > 0.5}, flush=True)
free, total = torch.cuda.mem_get_info(0)
print("device memory in use: %.1f GB of %.1f GB" % ((total - free) / 2**30, total / 2**30), flush=True)
assert len(proof) == sb.load_library().sb_proof_size(log_n)
if verify:
    from oracle import binding as ob
    t0 = time.time()
    if ocs is None:
        ocs = ob.R1CS.from_csr(log_n, cs.mats)
    vp = ob.PP.verifier_only(log_n, G1_GENERATOR, G2_GENERATOR, pp.g_mask_random())
    r = ob.verify(ocs, vp, cs.v, proof)
    print("CPU verifier -> %d (%.1fs)" % (r, time.time() - t0), flush=True)
    assert r == 1
print("ok")
