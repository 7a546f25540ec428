"""Dump the CUDA-event timeline of one prove (run on the GPU box): python tests/gpu_timeline.py [log_n]"""
import json, os, sys
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
import time
for _ in range(4):
    sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit)
ts = []
for _ in range(8):
    t0 = time.perf_counter(); _, ph0 = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases"); ts.append((time.perf_counter() - t0) * 1e3)
knobs = {k: v for k, v in os.environ.items() if k.startswith("SB_")}
print("STEADY log_n=%d %s median %.2f ms min %.2f  %s" % (log_n, knobs, sorted(ts)[len(ts) // 2], min(ts), {k: round(v, 2) for k, v in ph0.items()}))
if os.environ.get("SB_NO_TIMELINE"):
    sys.exit(0)
ctx.prof_enable(True); ctx.prof_report()
_, ph = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases")
tl = ctx.prof_timeline()
ctx.prof_enable(False)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"phases": ph, "timeline": tl}, open("gpurun_out/timeline_%d%s.json" % (log_n, os.environ.get("SB_TAG", "")), "w"))
print(ph)
# coarse view: per 1-ms bucket, which kernels are running
end = max(t[2] for t in tl)
print("span %.2f ms, %d launches" % (end, len(tl)))
big = [t for t in tl if t[2] - t[1] > 0.5]
for t in sorted(big, key=lambda t: t[1]):
    print("%8.2f -> %8.2f  (%6.2f)  %s #%d" % (t[1], t[2], t[2] - t[1], t[0], t[3]))
