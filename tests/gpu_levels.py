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
# first opening: kernels between the first k_open_fold and the first k_segsum, grouped by the job tag (log2 of its size)
fold0 = [t for t in tl if t[0] == "k_open_fold"][0][1]
seg0 = [t for t in tl if t[0] == "k_segsum"][0][1]
commit = [t for t in tl if t[1] < fold0]
print("commit: " + " ".join("%s=%.3f" % (t[0].replace("k_seg_accum_", "").replace("k_bucket_", "").replace("k_msm_", ""), t[2] - t[1]) for t in commit))
rows = {}
for name, t0, t1, tag in tl:
    if fold0 < t0 < seg0 and tag >= 0:
        r = rows.setdefault(tag, {"mixed": 0.0, "full": 0.0, "nfull": 0, "reduce1": 0.0, "reduce2": 0.0, "sort": 0.0})
        if "mixed" in name: r["mixed"] += t1 - t0
        elif "full" in name: r["full"] += t1 - t0; r["nfull"] += 1
        elif "reduce1" in name: r["reduce1"] += t1 - t0
        elif "reduce2" in name: r["reduce2"] += t1 - t0
        else: r["sort"] += t1 - t0
for tag in sorted(rows, reverse=True):
    r = rows[tag]
    print("2^%-2d sort %.3f mixed %.3f full %.3f (%d) reduce1 %.3f reduce2 %.3f" % (tag, r["sort"], r["mixed"], r["full"], r["nfull"], r["reduce1"], r["reduce2"]))
print("SUM knobs=%s mixed %.2f full %.2f reduce1 %.2f reduce2 %.2f sort %.2f" % (
    {k: v for k, v in os.environ.items() if k.startswith("SB_")},
    sum(r["mixed"] for r in rows.values()), sum(r["full"] for r in rows.values()), sum(r["reduce1"] for r in rows.values()),
    sum(r["reduce2"] for r in rows.values()), sum(r["sort"] for r in rows.values())))
