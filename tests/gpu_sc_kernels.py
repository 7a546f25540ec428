import sys, os, json
sys.path.insert(0, os.getcwd())
import r1cs_spartan_b200 as sb
ctx = sb.Context(0)
fr_ms = ctx.mul_bench("fr", 148 * 1024, 1000); peak = 148 * 1024 * 1000 * 2 / fr_ms / 1e6
for which, nm, mults in ((0, "sc1_fused", 3.0), (1, "sc1_first", 3.0), (2, "sc2_fused", 1.75), (3, "open_fold", 0.5)):
    for lg in (18, 20, 22):
        ms = ctx.kernel_bench(which, lg, reps=10, flush_l2=True)
        print("%-10s 2^%d %.4f ms  imad %.3f" % (nm, lg, ms, mults * (1 << lg) / ms / 1e6 / peak))
