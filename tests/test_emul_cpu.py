"""CPU tests of the PRODUCT's own CUDA sources under the CUDA-on-CPU test shim (tests/emul/).

tests/emul/build.sh compiles r1cs-spartan_b200/csrc/*.cu, unchanged, with g++ against tests/emul/cuda_runtime.h: every
kernel launch runs its CUDA threads as host threads.  That executes the index logic of every kernel (MSM digits /
plan / scatter / chunked accumulation / bucket reduction over multi-slot groups, sparse-product plans, sumcheck folds
and reductions) and the complete host driver -- single-GPU and hypercube-sharded -- against the CPU oracle, here, without
a GPU.  What it does NOT cover is the PTX field arithmetic (the portable C++ path of field.cuh runs instead); the -m gpu
tests do that on the device.  The emulation library is test infrastructure: it lives under tests/, refuses to create a
context unless SB_EMUL_TESTS=1, and the product never builds or loads it.
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")
LIB = os.path.join(EMUL, "libsb_emul_TESTONLY.so")


@pytest.fixture(scope="module")
def emul_env():
    if not (os.path.exists("/usr/bin/g++") or shutil.which("g++")):
        pytest.skip("no host compiler")
    subprocess.check_call([os.path.join(EMUL, "build.sh")])
    return dict(os.environ, SB_EMUL_TESTS="1", SB_LIB_PATH=LIB)


def _pytest_gpu_subset(env, k, extra=None, timeout=1500):
    e = dict(env, **(extra or {}))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k", k],
                       env=e, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (extra, r.stdout[-3000:], r.stderr[-2000:])
    assert " passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], r.stdout[-500:]


def test_emulated_library_refuses_to_run_outside_the_tests(emul_env):
    env = {k: v for k, v in emul_env.items() if k != "SB_EMUL_TESTS"}
    r = subprocess.run([sys.executable, "-c", "import r1cs_spartan_b200 as sb; sb.Context(0)"], env=env, cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def test_product_kernels_match_oracle_under_emulation(emul_env):
    # the stage tests and small complete proofs of tests/test_gpu_parity.py, through the C ABI of the emulated build
    _pytest_gpu_subset(emul_env, "eq_table or sum_over_y or eval_on_x or synthetic_circuit or device_indexer or keygen or pp_load or msm_matches or structured "
                                 "or adversarial or commit_and_open or padded or interactive or invalid_arguments or resident or closed_before "
                                 "or prove_bytes_and_trace_match_oracle[3 or prove_bytes_and_trace_match_oracle[6 or prove_bytes_and_trace_match_oracle[8")


@pytest.mark.parametrize("knobs", [
    {"SB_MSM_S0": "2", "SB_MSM_S1": "2", "SB_MSM_LEVELS": "2"},                       # two launched levels: the device must raise S1
    {"SB_MSM_S0": "3", "SB_MSM_S1": "2", "SB_MSM_LEVELS": "8", "SB_MSM_SORTED": "0", "SB_MSM_RED_L": "8"},
    {"SB_MSM_SPLIT": "3", "SB_MSM_SPLIT_MIN_NV": "2", "SB_MSM_S0": "5"},                # opening ladders split into three pipelines
    {"SB_MSM_AFFINE_LOG2": "3", "SB_MSM_AFFINE_ROUNDS": "3", "SB_MSM_AFFINE_K": "3", "SB_MSM_S0": "4"},   # pairwise affine rounds in front of the accumulation (G1 and G2)
])
def test_msm_pipeline_knobs_under_emulation(emul_env, knobs):
    _pytest_gpu_subset(emul_env, "structured or adversarial or commit_and_open or prove_bytes_and_trace_match_oracle[6", knobs)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_prover_under_emulation(emul_env, world, oracle):
    # the hypercube-sharded prover, one process per (emulated) GPU, shared-memory exchange: proof bytes and every traced
    # intermediate equal the oracle's on every rank; sharded keygen and sliced sb_pp_load both
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_multi as tm
    old = {k: os.environ.get(k) for k in ("SB_EMUL_TESTS", "SB_LIB_PATH", "SB_EMUL_DEVICES")}
    os.environ.update(SB_EMUL_TESTS="1", SB_LIB_PATH=LIB, SB_EMUL_DEVICES="8")
    try:
        tm._run(world, [(3, 4, 0, False), (5, 8, 30, True)] if world == 2 else [(4, 4, 0, True)], backend="gloo")
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_multi_context_one_process_under_emulation(emul_env):
    # sb_ctx_create_multi: one context over 2 / 4 (emulated) devices of ONE process, worker thread per device, in-process
    # exchange; proof bytes and traced intermediates equal the oracle's, error mapping preserved
    e = dict(emul_env, SB_EMUL_DEVICES="8")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_multi.py"), "-m", "gpu", "-x", "-q", "-k",
                        "multi_context_one_process_2gpu or multi_context_one_process_4gpu"], env=e, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert "2 passed" in r.stdout, r.stdout[-500:]


def test_multi_context_argument_checks_under_emulation(emul_env):
    # sb_ctx_create_multi: a power-of-two number of devices, every index in range; one device gives a plain context;
    # the single-matrix helpers refuse a multi context with InvalidArgument instead of touching a shard behind the exchange
    code = """
import numpy as np, r1cs_spartan_b200 as sb
assert sb.device_count() == 8
for bad in ([0, 1, 2], [], [0, 99]):
    try:
        sb.Context(devices=bad)
    except (sb.InvalidArgument, sb.CudaError, ValueError):
        pass
    else:
        raise SystemExit("accepted devices=%r" % (bad,))
one = sb.Context(devices=[3])
assert np.array_equal(sb.eq_extension(np.zeros((2, 4), dtype=np.uint64), ctx=one).shape, (4, 4))
one.close()
m = sb.Context(devices=[0, 1])
try:
    sb.eq_extension(np.zeros((2, 4), dtype=np.uint64), ctx=m)
except sb.InvalidArgument:
    pass
else:
    raise SystemExit("eq_extension accepted a multi context")
cs = sb.SyntheticR1CS(4, 12, 0, 9)
pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=m)
try:
    pk.sum_over_y(np.concatenate([cs.v, cs.w]))
except sb.InvalidArgument:
    pass
else:
    raise SystemExit("sum_over_y accepted a multi context")
m.close(); pk.close()          # context first: its destruction completes with the last handle
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], env=dict(emul_env, SB_EMUL_DEVICES="8"), cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
