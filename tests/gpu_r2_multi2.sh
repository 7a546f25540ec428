# round-2 (GPU box, --gpus 2): sharded prover tests (one process per GPU, and ONE process over 2 GPUs), bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2gpu" 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_multi_n2.log
if [ -z "$SB_SKIP_BENCH" ]; then python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 2500 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err; fi
python - <<'PY'
# one process, one context over both GPUs: steady-state prove time at 2^20 and equality with the single-GPU proof
import os, sys, time, hashlib
sys.path.insert(0, os.getcwd())
import numpy as np
import r1cs_spartan_b200 as sb
from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
from r1cs_spartan_b200 import workload as wl
log_n = 20
cs = sb.SyntheticR1CS(32, (1 << log_n) - 32, 0, 0x5EED0000 + log_n)
trap = np.stack([wl.mont_to_limbs([wl.fr_rand_mont(wl.SplitMix64(99 + i))])[0] for i in range(log_n)])
res = {}
for devs in ([0], [0, 1]):
    ctx = sb.Context(devices=devs) if len(devs) > 1 else sb.Context(0)
    pp = sb.MLPolyCommit.keygen(log_n, G1_GENERATOR, G2_GENERATOR, trap, ctx=ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
    wit = sb.Witness(pk, cs.v, cs.w)
    for _ in range(3):
        proof = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); proof = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit); ts.append(1e3 * (time.perf_counter() - t0))
    res[len(devs)] = (sorted(ts)[len(ts) // 2], hashlib.sha256(proof).hexdigest())
    print("ONE PROCESS, %d GPU(s): median %.2f ms  sha256 %s" % (len(devs), res[len(devs)][0], res[len(devs)][1]), flush=True)
    wit.close(); pk.close(); pp.close(); ctx.close()
assert res[1][1] == res[2][1], "multi-GPU context proof differs from the single-GPU proof"
print("multi-context proof == single-GPU proof")
PY
