"""Integer-pipe microbenchmark (run on the GPU box): Montgomery products per second for Fr and Fq."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r1cs_spartan_b200 as sb

ctx = sb.Context(0)
out = {}
for field, imad in (("fr", 139), ("fq", 303)):
    for threads_per_sm in (256, 512, 1024, 2048):
        n = 148 * threads_per_sm
        iters = 2000
        ms = ctx.mul_bench(field, n, iters)
        muls = n * iters * 2
        out["%s_%d" % (field, threads_per_sm)] = {"ms": ms, "Gmul_s": muls / ms / 1e6, "imad_T_s": muls * imad / ms / 1e9}
for which, name in ((0, "sc1_fused"), (1, "sc1_first"), (2, "sc2_fused"), (3, "open_fold")):
    for log_m in (16, 20, 22):
        out["%s_2^%d_ms" % (name, log_m)] = ctx.kernel_bench(which, log_m, reps=10, flush_l2=True)
print(json.dumps(out, indent=1))
