"""Per-ladder-level durations of the G2 MSM kernels in serialised mode (run on the GPU box)."""
import json, os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import r1cs_spartan_b200 as sb
from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
from r1cs_spartan_b200 import workload as wl

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = sb.Context(0)
cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)
trap = np.stack([wl.mont_to_limbs([wl.fr_rand_mont(wl.SplitMix64(99 + i))])[0] for i in range(log_n)])
pp = sb.MLPolyCommit.keygen(log_n, G1_GENERATOR, G2_GENERATOR, trap, ctx=ctx)
pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
wit = sb.Witness(pk, cs.v, cs.w)
for _ in range(3):
    sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit)
ctx.set_serial_msm(True)
ctx.prof_enable(True); ctx.prof_report()
sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit)
tl = ctx.prof_timeline()
ctx.prof_enable(False)
# group the kernels of the first opening by MSM: each MSM = digits, scan_plan, scatter, then accum levels, reduce1, reduce2
seq = [t for t in tl]
msms, cur = [], None
for name, t0, t1 in seq:
    if name == "k_msm_digits":
        cur = {"digits": t1 - t0, "k": []}; msms.append(cur)
    elif cur is not None and name.startswith(("k_scan_plan", "k_msm_scatter", "k_seg_accum", "k_bucket_reduce")):
        cur["k"].append((name, t1 - t0))
for i, m in enumerate(msms[:22]):
    print(i, " ".join("%s=%.3f" % (n.replace("k_seg_accum_", "").replace("k_bucket_", "").replace("k_msm_", ""), d) for n, d in m["k"]))
