# round-2 final (GPU box): full GPU test suite, the default bench line (with one real full-size CPU proof), launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu_final.log
python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 600 gpurun_out/r02_bench_n1_final.json; tail -3 gpurun_out/r02_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
python tests/gpu_sc_kernels.py 2>&1 | tail -12
# C5 cross-check: the same 2^24 instance on ONE GPU must give the proof the 8-GPU context gave (sha256 in profiles/r02_multi8_one_process.txt)
python - <<'PY' 2>&1 | tee gpurun_out/r02_scale_2e24_one_gpu.log
import hashlib, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import r1cs_spartan_b200 as sb
from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
from r1cs_spartan_b200 import workload as wl
from oracle import binding as ob
ob.build(); ob.lib()
log_n = 24
t0 = time.time(); ocs = ob.R1CS.synth(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n); mats = [ocs.csr(k) for k in range(3)]; v, w = ocs.vw(); print("workload %.1fs" % (time.time() - t0), flush=True)
ctx = sb.Context(0)
trap = np.stack([wl.mont_to_limbs([wl.fr_rand_mont(wl.SplitMix64(99 + i))])[0] for i in range(log_n)])
t0 = time.time(); pp = sb.MLPolyCommit.keygen(log_n, G1_GENERATOR, G2_GENERATOR, trap, ctx=ctx); print("keygen %.1fs" % (time.time() - t0), flush=True)
t0 = time.time(); pk = sb.MLArgumentForR1CS.index(*mats, ctx=ctx); print("index %.2fs (device-side plans incl. upload %.0f ms, transcript hash added %.0f ms)" % ((time.time() - t0,) + pk.timing()), flush=True)
wit = sb.Witness(pk, v, w)
for i in range(3):
    t0 = time.perf_counter(); proof, ph = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases")
    print("prove 2^24 on 1 GPU: %.1f ms" % (1e3 * (time.perf_counter() - t0)), {k: round(x, 1) for k, x in ph.items() if x > 0.5}, flush=True)
print("proof %d bytes sha256 %s" % (len(proof), hashlib.sha256(proof).hexdigest()), flush=True)
PY
