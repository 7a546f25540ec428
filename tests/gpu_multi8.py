"""Multi-GPU configs of BASELINE.json on ONE process / ONE context over N GPUs (sb_ctx_create_multi); run on the GPU box:
    python tests/gpu_multi8.py N [c4] [prove20] [c5]
  c4       config 4: commit / open / verify of a random 2^20-entry table, MSMs sharded over N GPUs, checked by the CPU pairing
           verifier (oracle pc_verify; a wrong value must fail)
  prove20  the 2^20 benchmark proof on the multi context: median time, sha256 (bench.py's single-GPU proof has the same)
  c5       config 5: 2^24 constraints: keygen / index / prove times, per-GPU memory, sha256 of the proof
The oracle makes inputs and checks outputs only (test infrastructure)."""
import hashlib, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import r1cs_spartan_b200 as sb
from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
from r1cs_spartan_b200 import workload as wl
from oracle import binding as ob

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else sb.device_count()
what = set(sys.argv[2:]) or {"c4", "prove20"}
ob.build(); ob.lib()
ctx = sb.Context(devices=list(range(ndev)))
print("one process, one context over %d GPUs" % ndev, flush=True)


def trapdoor(log_n):
    return np.stack([wl.mont_to_limbs([wl.fr_rand_mont(wl.SplitMix64(99 + i))])[0] for i in range(log_n)])


def mem():
    return ["%.1f" % ((t - f) / 2**30) for f, t in (torch.cuda.mem_get_info(i) for i in range(ndev))]


if "c4" in what or "prove20" in what:
    log_n = 20
    t0 = time.time(); pp = sb.MLPolyCommit.keygen(log_n, G1_GENERATOR, G2_GENERATOR, trapdoor(log_n), ctx=ctx); print("2^20 keygen %.1fs" % (time.time() - t0), flush=True)
if "c4" in what:
    vp = ob.PP.verifier_only(log_n, G1_GENERATOR, G2_GENERATOR, pp.g_mask_random())
    table = ob.fr_rand(5 + log_n, 1 << log_n); point = ob.fr_rand(6 + log_n, log_n)
    for rep in range(3):
        t0 = time.perf_counter(); _, com = sb.MLPolyCommit.commit(pp, table); t1 = time.perf_counter()
        ev, (_, proofs) = sb.MLPolyCommit.open(pp, table, point); t2 = time.perf_counter()
        print("C4 commit %.1f ms  open %.1f ms (host table in, %d GPUs)" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), ndev), flush=True)
    t0 = time.time()
    ok = ob.pc_verify(vp, com, point, ev, proofs); bad = ob.pc_verify(vp, com, point, ob.fr_rand(7, 1)[0], proofs)
    print("C4 CPU pairing verifier: accepts the opening = %s, accepts a wrong value = %s (%.1fs)" % (bool(ok), bool(bad), time.time() - t0), flush=True)
    assert ok and not bad
if "prove20" in what:
    cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
    wit = sb.Witness(pk, cs.v, cs.w)
    for _ in range(3):
        proof = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); proof, ph = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases"); ts.append(1e3 * (time.perf_counter() - t0))
    print("2^20 prove on %d GPUs, one process: median %.2f ms min %.2f  sha256 %s" % (ndev, sorted(ts)[len(ts) // 2], min(ts), hashlib.sha256(proof).hexdigest()), flush=True)
    print("   phases (rank 0):", {k: round(v, 2) for k, v in ph.items()}, flush=True)
    wit.close(); pk.close()
if "c4" in what or "prove20" in what:
    pp.close()
if "c5" in what:
    log_n = 24
    t0 = time.time()
    ocs = ob.R1CS.synth(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)
    mats = [ocs.csr(k) for k in range(3)]; v, w = ocs.vw()
    print("C5 workload %.1fs nnz %s" % (time.time() - t0, [int(m[0][-1]) for m in mats]), flush=True)
    t0 = time.time(); pp = sb.MLPolyCommit.keygen(log_n, G1_GENERATOR, G2_GENERATOR, trapdoor(log_n), ctx=ctx); print("C5 keygen %.1fs" % (time.time() - t0), flush=True)
    t0 = time.time(); pk = sb.MLArgumentForR1CS.index(*mats, ctx=ctx)
    print("C5 index %.2fs (device-side plans incl. upload %.0f ms, transcript hash added %.0f ms)" % ((time.time() - t0,) + pk.timing()), flush=True)
    wit = sb.Witness(pk, v, w)
    for i in range(4):
        t0 = time.perf_counter()
        proof, ph = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases")
        print("C5 prove 2^24 on %d GPUs: %.1f ms" % (ndev, 1e3 * (time.perf_counter() - t0)), {k: round(x, 1) for k, x in ph.items() if x > 0.5}, flush=True)
    print("C5 device memory in use per GPU (GB):", mem(), flush=True)
    print("C5 proof %d bytes sha256 %s" % (len(proof), hashlib.sha256(proof).hexdigest()), flush=True)
    assert len(proof) == sb.load_library().sb_proof_size(log_n)
print("ok")
