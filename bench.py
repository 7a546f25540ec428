#!/usr/bin/env python3
"""bench.py -- prove time of the reference's synthetic R1CS benchmark on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
    python bench.py --impl reference --steps K --warmup W    the CPU restatement of the reference

One "step" = one non-interactive proof (MLArgumentForR1CS::prove, /root/reference/src/lib.rs:58-146) of the
benchmark circuit (/root/reference/src/benchmark.rs:63-65: 32 public inputs, density 0) at 2^LOG_N
constraints, with the index (IndexPK) and the public parameters (PublicParameter) pre-built -- exactly
what the reference's own harness times under "Prove" (benchmark.rs:34-41).

  value : ms per proof with the witness z already resident in HBM (sb_prove_resident)
  e2e   : ms per proof through the reference-facing call with HOST buffers (sb_prove): H2D of v, w and
          D2H of every prover message inside the timed region; the proof comes back as bytes
Timing: CUDA events on the library's stream are used for the per-kernel split; the step time itself is
host wall-clock bracketed by device synchronisation (every step ends with the proof bytes on the host,
so the device is idle at both ends), max over ranks.
"""
import os as _os
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before any CUDA initialisation (see api.load_library)
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = int(os.environ.get("SB_BENCH_LOG_N", "20"))
NUM_PUBLIC = 32
METRIC = "prove_time_2^%d_constraints" % LOG_N
UNIT = "ms"


def workload_name(log_n):
    return "synthetic R1CS 2^%d constraints (reference benchmark circuit, 32 public inputs, density 0, BLS12-381)" % log_n


# ---------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            # nvidia-smi takes a few hundred milliseconds to initialise NVML, and while it does, driver calls of other
            # processes can stall: seen in round 2 as ONE 0.6-1.3 s step somewhere after the sampler was started (once in
            # the e2e region: 160 ms instead of 40 ms per step).  Wait for its first sample before any timed region starts.
            t0 = time.perf_counter()
            while not self.rows and self.proc.poll() is None and time.perf_counter() - t0 < 8.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------- CPU baseline (oracle; the one place bench may run it)
def cpu_prove_sample(log_n_sample):
    """Times the literal single-threaded CPU restatement of the reference on the same circuit family."""
    from oracle import binding as ob
    import r1cs_spartan_b200 as sb
    cs = sb.SyntheticR1CS(NUM_PUBLIC, (1 << log_n_sample) - NUM_PUBLIC, 0, 0x5EED0000 + log_n_sample)
    ocs = ob.R1CS.from_csr(log_n_sample, cs.mats)
    pp = ob.PP.keygen(log_n_sample, 99)
    t0 = time.perf_counter()
    proof, tr = ob.prove(ocs, pp, cs.v, cs.w)
    dt = time.perf_counter() - t0
    phases = {k: tr.time(k) for k in ("prove1_commit", "prove2_open", "sumcheck1", "prove5_eval_on_x", "sumcheck2", "prove6_open")}
    return dt, phases, len(proof)


def _ark_msm_cost(n):
    """additions of arkworks' serial bucket method (window c = ln(n) + 2): ceil(255 / c) * (n + 2^(c+1))"""
    if n < 32:
        c = 3
    else:
        c = (max(n - 1, 1).bit_length() * 69) // 100 + 2
    return -(-255 // c) * (n + (2 << c))


def extrapolate_cpu(phases, total_s, ls, lt):
    """Scale the phase times of a CPU proof at 2^ls constraints to 2^lt with the operation counts of the literal
    algorithm instead of a blanket factor 2^(lt - ls): MSM work per point FALLS with n (wider windows), the
    reference's first sumcheck grows like n * (l + 3)(2l + 3), everything else is linear."""
    lin = float(1 << (lt - ls))
    r_commit = _ark_msm_cost(1 << lt) / _ark_msm_cost(1 << ls)
    r_open = sum(_ark_msm_cost(1 << k) for k in range(1, lt + 1)) / sum(_ark_msm_cost(1 << k) for k in range(1, ls + 1))
    r_sc1 = lin * ((lt + 3) * (2 * lt + 3)) / ((ls + 3) * (2 * ls + 3))
    known = ("prove1_commit", "prove2_open", "prove6_open", "sumcheck1", "prove5_eval_on_x", "sumcheck2")
    rest = max(total_s - sum(phases[k] for k in known), 0.0)
    est = (phases["prove1_commit"] * r_commit + (phases["prove2_open"] + phases["prove6_open"]) * r_open + phases["sumcheck1"] * r_sc1 +
           (phases["prove5_eval_on_x"] + phases["sumcheck2"] + rest) * lin)
    return est, {"commit_x": r_commit, "open_x": r_open, "sumcheck1_x": r_sc1, "linear_x": lin}


def cpu_model():
    import platform
    try:
        model = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        model = platform.processor()
    return model


def bench_trapdoor(log_n):
    """the trapdoor both arms use for their public parameters (SplitMix64 seeds 99 + i, arkworks' Fr sampling rule)"""
    import numpy as np
    import r1cs_spartan_b200 as sb
    return np.stack([sb.workload.mont_to_limbs([sb.workload.fr_rand_mont(sb.workload.SplitMix64(99 + i))])[0] for i in range(log_n)])


def cpu_prove_full(log_n, opp=None, setup_threads=None):
    """ONE real proof of the bench workload by the literal CPU restatement, single-threaded (the reference's
    configuration, Cargo.toml:26).  opp: oracle parameters made elsewhere (our arm passes the GPU-made ones so that
    the proof bytes can be compared); otherwise the oracle's own keygen runs first, untimed, on all host cores."""
    from oracle import binding as ob
    import r1cs_spartan_b200 as sb
    from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
    cs = sb.SyntheticR1CS(NUM_PUBLIC, (1 << log_n) - NUM_PUBLIC, 0, 0x5EED0000 + log_n)
    ocs = ob.R1CS.from_csr(log_n, cs.mats)
    t_setup = time.perf_counter()
    if opp is None:
        ob.set_threads(setup_threads or os.cpu_count() or 1)
        opp = ob.PP.keygen_with(log_n, G1_GENERATOR, G2_GENERATOR, bench_trapdoor(log_n))
    ob.set_threads(1)                               # the timed part is one thread, always
    setup_s = time.perf_counter() - t_setup
    t0 = time.perf_counter()
    proof, tr = ob.prove(ocs, opp, cs.v, cs.w)
    dt = time.perf_counter() - t0
    phases = {k: tr.time(k) for k in ("transcript_init", "prove1_commit", "prove2_open", "prove3_setup", "sumcheck1", "prove4",
                                       "prove5_eval_on_x", "sumcheck2", "prove6_open")}
    return dt, phases, proof, setup_s


def run_reference(args):
    """The reference arm: the reference's own prover cannot be built here (Rust over un-vendored arkworks git
    dependencies, no cargo), so this times oracle/ -- its literal single-threaded C++ restatement -- on ONE REAL proof
    of the very workload our arm proves (2^LOG_N constraints; about 4.5 minutes at 2^20).  Nothing is extrapolated:
    `steps` is the number of proofs actually timed (1 unless SB_REF_STEPS says otherwise), whatever --steps asked."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import hashlib
    steps = max(int(os.environ.get("SB_REF_STEPS", "1")), 1)
    times, phases, proof, setup_s = [], None, None, 0.0
    opp = None
    for _ in range(steps):
        dt, phases, proof, s_s = cpu_prove_full(LOG_N, opp)
        times.append(dt); setup_s += s_s
    value = 1e3 * sum(times) / len(times)
    sample = ("%d full single-threaded proof(s) of the same workload (2^%d constraints) by the literal C++ restatement of the "
              "reference (2 log n + 3 table sumcheck, duplicated-scalar G2 MSMs, arkworks' window rule); nothing extrapolated; "
              "public parameters made by the restatement's own keygen on all host cores, untimed (%.1f s); --steps %d / --warmup %d "
              "were requested, %d step(s) timed, no warm-up" % (steps, LOG_N, setup_s, args.steps, args.warmup, steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 0, "requested_steps": args.steps, "requested_warmup": args.warmup, "timed_steps": steps,
        "ms_per_step": value, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 limbs (4 x u64 Montgomery Fr, 6 x u64 Fq; exact integer arithmetic)", "data": "synthetic",
        "same_config": True,
        "config": {"workload": workload_name(LOG_N), "parallelism": "1 CPU thread (the reference never enables arkworks' `parallel`)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cpu": cpu_model(), "host_cores_available": os.cpu_count(), "phases_s": phases},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "proof_sha256": hashlib.sha256(proof).hexdigest(), "proof_bytes": len(proof),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import r1cs_spartan_b200 as sb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    if world > 1:
        from r1cs_spartan_b200 import dist as sbdist
        ctx = sbdist.sharded_context(local_rank)
    else:
        ctx = sb.Context(local_rank)
    t_setup = time.perf_counter()
    cs = sb.SyntheticR1CS(NUM_PUBLIC, (1 << LOG_N) - NUM_PUBLIC, 0, 0x5EED0000 + LOG_N)
    # generators: any r-torsion generators are a valid pp (the reference samples them); these are the ones the
    # tests use (G1: the standard generator; G2: derived in oracle/pymodel.py), written as Montgomery limbs
    from r1cs_spartan_b200.generators import G1_GENERATOR, G2_GENERATOR
    trap = bench_trapdoor(LOG_N)
    cpu_log = LOG_N if args.cpu_sample_log_n is None else args.cpu_sample_log_n
    want_cpu = world == 1 and not args.no_cpu_baseline
    # one GPU + CPU baseline at the full size: keep every parameter level so that the oracle can prove on the very same ones
    pp = sb.MLPolyCommit.keygen(LOG_N, G1_GENERATOR, G2_GENERATOR, trap, keep_all_levels=(want_cpu and cpu_log == LOG_N), ctx=ctx)
    pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
    wit = sb.Witness(pk, cs.v, cs.w)
    setup_s = time.perf_counter() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_log = []

    event_s = {}

    def timed(fn, steps, tag=None):
        """K steps bracketed by barrier + synchronize on both sides.  Every step ends with the proof bytes on the host, so
        the device is idle at both brackets; the region is timed with CUDA events (recorded on torch's current stream while
        nothing is queued: they stamp the device timeline) and with the host clock -- the two agree to microseconds, the
        event time is the one reported.  Max over ranks."""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            ts = time.perf_counter()
            fn()
            step_log.append(round(1e3 * (time.perf_counter() - ts), 3))
        e1.record()
        torch.cuda.synchronize()
        dt_host = time.perf_counter() - t0
        dt = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([dt, dt_host], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_host = float(t[0].item()), float(t[1].item())
        if tag:
            event_s[tag] = {"cuda_events_s": dt, "host_clock_s": dt_host}
        return dt

    proofs, phase_log = [], []

    def resident():
        pr, ph = sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit, trace="phases")
        proofs.append(pr); phase_log.append(ph)

    # the end-to-end arm reads its inputs from PINNED host memory (the contract's "host->device copy of that step's
    # inputs from pinned host memory"), as a caller that cares about transfer time would provide them
    v_pin = torch.from_numpy(cs.v.view(np.int64)).pin_memory().numpy().view(np.uint64)
    w_pin = torch.from_numpy(cs.w.view(np.int64)).pin_memory().numpy().view(np.uint64)

    def e2e():
        pr, ph = sb.MLArgumentForR1CS.prove(pk, v_pin, w_pin, pp, trace="phases")
        proofs.append(pr); phase_log.append(ph)
    for _ in range(max(args.warmup, 3)):
        resident()
    for _ in range(max(args.warmup, 3)):       # the end-to-end arm gets its own W warm-up steps (first-touch of the pinned staging path)
        e2e()
    # ---- correctness evidence the driver can see: one hash per line, identical on every rank, and -- when the prover is
    # sharded -- identical to the proof a plain single-GPU context makes for the same instance on rank 0's GPU
    import hashlib
    proof_sha = hashlib.sha256(proofs[0]).hexdigest()
    sharded_check = None
    if world > 1:
        hashes = [None] * world
        dist.all_gather_object(hashes, proof_sha)
        assert all(h == hashes[0] for h in hashes), "ranks disagree on the proof: %s" % hashes
        if rank == 0:
            ctx1 = sb.Context(local_rank)
            pp1 = sb.MLPolyCommit.keygen(LOG_N, G1_GENERATOR, G2_GENERATOR, trap, ctx=ctx1)
            pk1 = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx1)
            single_proof = sb.MLArgumentForR1CS.prove(pk1, cs.v, cs.w, pp1)
            assert single_proof == proofs[0], "the sharded proof differs from the single-GPU proof of the same instance"
            sharded_check = {"ranks_agree": True, "equals_single_gpu_proof": True, "single_gpu_proof_sha256": hashlib.sha256(single_proof).hexdigest()}
            pk1.close(); pp1.close(); del pk1, pp1
        dist.barrier()
    # ---- timed region: K proofs with the witness resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("SB_NO_CLOCKS"):
        sampler.start()
    del phase_log[:]
    l0 = ctx.launch_count()
    dt = timed(resident, args.steps, "resident")
    launches = (ctx.launch_count() - l0) // args.steps
    steps_resident = list(step_log); del step_log[:]
    phases_resident = list(phase_log); del phase_log[:]
    # ---- e2e: host buffers in, proof bytes out
    h0, d0 = ctx.copy_counters()
    dt_e2e = timed(e2e, args.steps, "e2e")
    h1, d1 = ctx.copy_counters()
    steps_e2e = list(step_log); del step_log[:]
    clocks = sampler.stop() if rank == 0 else None
    # ---- per-kernel split: the same K steps again with CUDA events around every launch and the MSMs of an
    # opening serialised on one stream (concurrent streams would charge queueing time to the kernels)
    ctx.set_serial_msm(True)
    ctx.prof_enable(True)
    ctx.prof_report()
    dt_serial = timed(resident, args.steps)
    prof = ctx.prof_report()
    # one more serialised proof for the per-level split of the G2 accumulation (records carry log2 of the job size)
    resident()
    timeline = ctx.prof_timeline()
    ctx.prof_enable(False)
    ctx.set_serial_msm(False)
    assert all(p == proofs[0] for p in proofs), "proofs differ between steps"
    _, tr = sb.MLArgumentForR1CS.prove(pk, None, None, pp, trace=True, witness=wit)

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    ms = 1e3 * dt / args.steps
    ms_e2e = 1e3 * dt_e2e / args.steps
    # ---- roofline of the dominant kernel, from the CUDA-event split of the serialised steps
    n = 1 << LOG_N
    kernels = {k: {"launches_per_step": v["launches"] // args.steps, "ms_per_step": v["ms"] / args.steps} for k, v in prof.items()}
    # integer-pipe ceiling measured in the same run: dependent-free Montgomery products (2 chains per thread)
    single = sb.Context(local_rank) if world > 1 else ctx
    fq_ms = single.mul_bench("fq", 148 * 1024, 1000)
    fr_ms = single.mul_bench("fr", 148 * 1024, 1000)
    fq_peak = 148 * 1024 * 1000 * 2 / fq_ms / 1e6     # G Fq-mul/s
    fr_peak = 148 * 1024 * 1000 * 2 / fr_ms / 1e6
    work = msm_nominal_work(LOG_N - (world.bit_length() - 1), world)
    extra = {"roofline_imad": imad_roofline(kernels, work, fq_peak)}
    extra["roofline_imad"]["peak_fq_gmul_s"] = fq_peak
    extra["roofline_imad"]["peak_fr_gmul_s"] = fr_peak
    roofline = build_roofline(kernels, timeline, work, fq_peak, hbm_peak, peak_src, LOG_N, world)
    # sumcheck kernels alone, L2 flushed between launches
    sc = {}
    for which, nm, mults, bts in ((0, "sc1_fused_round", 12 / 4.0, 152.0), (1, "sc1_first_round", 6 / 2.0, (6 * 32 + 32) / 2.0), (2, "sc2_fused_round", 7 / 4.0, (8 * 32 + 4 * 32) / 4.0)):
        kms = single.kernel_bench(which, LOG_N, reps=10, flush_l2=True)
        sc[nm] = {"ms": kms, "fr_gmul_s": mults * n / kms / 1e6, "imad_frac": mults * n / kms / 1e6 / fr_peak,
                  "hbm_gb_s": bts * n / kms / 1e6, "hbm_frac": bts * n / kms / 1e6 / hbm_peak}
    extra["sumcheck_kernels"] = sc

    cpu = None
    if want_cpu:
        from oracle import binding as ob
        if cpu_log == LOG_N:
            # ONE REAL single-threaded proof of the same instance on the same (GPU-made, exported) public parameters; its
            # bytes must equal the GPU's.  About 4.5 minutes at 2^20; `--cpu-sample-log-n K` trades it for a 2^K sample + model.
            opp = ob.PP.from_arrays(LOG_N, pp.export(1, 0), np.concatenate([pp.export(2, i) for i in range(LOG_N)], axis=0), G2_GENERATOR)
            dtc, phases, cproof, _ = cpu_prove_full(LOG_N, opp)
            assert cproof == proofs[0], "the CPU restatement's proof differs from the GPU proof"
            cpu = {"value": 1e3 * dtc, "unit": UNIT, "cores": 1, "kind": "port", "same_config": True, "proof_bytes_equal_gpu": True,
                   "sample": "one full proof of the same instance (2^%d constraints, same public parameters: the GPU-made ones, exported) by the "
                             "literal C++ restatement of the reference on 1 core (the reference is single-threaded, Cargo.toml:26); nothing "
                             "extrapolated; its proof bytes equal the GPU's" % LOG_N,
                   "host_cpu": cpu_model(), "host_cores_available": os.cpu_count(), "phases_s": phases}
        else:
            dtc, phases, _ = cpu_prove_sample(cpu_log)
            scale = 1 << (LOG_N - cpu_log)
            est_s, factors = extrapolate_cpu(phases, dtc, cpu_log, LOG_N)
            cpu = {"value": 1e3 * est_s, "unit": UNIT, "cores": 1, "kind": "port", "same_config": False, "extrapolated": True,
                   "measured_sample_ms": 1e3 * dtc,
                   "sample": "full prove at 2^%d constraints took %.2f s on 1 core (the reference is single-threaded, Cargo.toml:26); scaled to 2^%d "
                             "per phase with the literal algorithm's operation counts (a blanket x%d would give %.0f ms)"
                             % (cpu_log, dtc, LOG_N, scale, 1e3 * dtc * scale),
                   "scaling_factors": factors,
                   "host_cpu": cpu_model(), "host_cores_available": os.cpu_count(), "sample_phases_s": phases}

    line = {
        "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 limbs (8 x u32 Montgomery Fr, 12 x u32 Fq; exact integer arithmetic)", "data": "synthetic",
        "config": {"workload": workload_name(LOG_N), "parallelism": "single GPU" if world == 1 else "hypercube sharded on the top %d variables" % (world.bit_length() - 1),
                   "l2": "working set (z, tables, 16x pre-shifted bases: > 4 GB) exceeds the 126 MB L2; no flush between steps",
                   "collective": ("none (single GPU)" if world == 1 else
                                  "data plane: per-round allgather of 96 B (and of one partial group element per MSM level) between the ranks' "
                                  "HOST threads through a POSIX shared-memory mailbox (%s); the round results are Fiat-Shamir material that goes "
                                  "through the host transcript anyway, so no device-side collective is on the path.  NCCL (torch.distributed) "
                                  "only carries this script's barrier and its max-over-ranks timing reduction" % os.environ.get("SB_COMM", "shm")),
                   "timing": "CUDA events around the K steps (device idle at both brackets), max over ranks; host clock alongside in timing_check",
                   "nnz": cs.nnz, "proof_bytes": len(proofs[0]), "setup_seconds_untimed": setup_s},
        "proof_sha256": proof_sha, "sharded_check": sharded_check, "timing_check": event_s,
        "e2e": {"value": ms_e2e, "unit": UNIT, "h2d_bytes_per_step": (h1 - h0) // args.steps, "d2h_bytes_per_step": (d1 - d0) // args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "phases_ms": tr.phase_ms,
        "step_ms": {"resident": steps_resident, "e2e": steps_e2e, "serialised_profiled_avg": 1e3 * dt_serial / args.steps},
        "step_phases_ms": [{k: round(v, 2) for k, v in ph.items()} for ph in phases_resident],
        "kernels": kernels,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def msm_layout(m, group_max=None):
    """window bits c and window count W of an MSM over m points in a group whose largest slot has group_max points
    (mirror of msm_layout in csrc/msm.cu)"""
    lg = max(m - 1, 0).bit_length()
    off = 0 if max(m, group_max or m) <= (1 << 17) else -3
    c = min(16, max(4, lg + off))
    rest = 255 - (c - 1)
    return c, (rest + c - 1) // c + 1


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the `ncu --set full` capture under profiles/
# The launch captured with --set full is the LARGEST one of the kernel (first pairwise round of the rest-of-ladder group of an
# opening: 5.0 M additions, 4.9 GB algorithmic); bench.py reports it as `roofline.traffic` with that caveat and repeats it, next
# to the launch's own algorithmic bytes, in `roofline.largest_launch`.
NCU_TRAFFIC = {("k_affine_round<Fq2>", 20): (4.194155e9 + 1.879138e9, "profiles/r02_ncu_affine_round_v3.txt")}


def build_roofline(kernels, timeline, work, fq_peak, hbm_peak, peak_src, log_n, world):
    """The `roofline` object of the bench line: the MSM kernel with the most serialised CUDA-event time against the
    integer-pipe ceiling (nominal Fq products per second), its largest launch, and HBM as the secondary figure.
    kernels: {name: {launches_per_step, ms_per_step}}; timeline: [(name, t0_ms, t1_ms, tag)]; work: msm_nominal_work()."""
    roofline = None
    msm_kernels = {k: v for k, v in kernels.items() if k in work}
    if msm_kernels:
        name = max(msm_kernels, key=lambda k: msm_kernels[k]["ms_per_step"])
        rec, w = kernels[name], work[name]
        per_launch_ms = rec["ms_per_step"] / rec["launches_per_step"]
        ach = w["fq_products"] / rec["ms_per_step"] / 1e6                     # G Fq-mul/s over all its launches of a step
        alg_bytes = w["bytes"] / rec["launches_per_step"]
        hbm_ach = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        ntr = NCU_TRAFFIC.get((name, log_n)) if world == 1 else None
        roofline = {"kernel": name, "bound": "imad", "achieved": ach, "peak": fq_peak, "unit": "G Fq-mul/s", "frac": ach / fq_peak,
                    "traffic": ntr[0] if ntr else None, "traffic_source": ntr[1] if ntr else None,
                    "launches_per_step": rec["launches_per_step"], "avg_launch_ms": per_launch_ms,
                    "algorithmic_fq_products_per_launch": w["fq_products"] / rec["launches_per_step"],
                    "algorithmic_unit": w["unit"],
                    "peak_source": "sb_mul_bench in this run (IMAD.WIDE issues at half rate: 148 SM x 32 lanes x clock; MEASURED_PEAKS.json holds no integer peak)",
                    "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                            "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                            "note": "secondary: the kernel is bound by instruction issue on the integer pipe, not by HBM"}}
        # the largest launch of that kernel (first pairwise round / first level of the largest group), from the serialised timeline
        big = [(t1 - t0, tag) for (nm, t0, t1, tag) in timeline if nm == name]
        if big and w.get("largest_launch_fq_products"):
            d_ms = max(big)[0]
            g = w["largest_launch_fq_products"] / d_ms / 1e6
            roofline["largest_launch"] = {"ms": d_ms, "achieved": g, "frac": g / fq_peak, "fq_products": w["largest_launch_fq_products"],
                                          "algorithmic_bytes": w.get("largest_launch_bytes"), "traffic": ntr[0] if ntr else None}
            if ntr:
                roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the LARGEST launch (ncu --set full); compare it with "
                                            "largest_launch.algorithmic_bytes, not with the per-launch average above")
    return roofline


def msm_nominal_work(ell, world=1):
    """Nominal (algorithmic) work of one proof's MSM kernels on one rank, by profiler kernel name: Fq products and HBM bytes.
    Mirrors the pipeline of csrc/prover.cu / csrc/msm.cu at its defaults: the commitment is one G1 group of 2^ell points; an
    opening ladder has slots of 2^(ell-1) ... 1 points, split into {slot 0} (the shared first proof element, computed ONCE per
    proof) and {the rest} (once per opening) when ell >= 12; a group of >= 2^21 (window, point) entries E runs R = 4 pairwise
    affine rounds (E (1 - 2^-R) affine additions: 5M + 1S = 17 Fq products over Fq2, 6 over Fq; the shared inversion, one per
    ~150 additions, is not counted) and then E / 2^R mixed XYZZ additions (8M + 2S = 28 / 10 Fq products); smaller groups
    run E mixed additions.  Bytes per affine addition: two x coordinates + prefix out in the first pass, two points + prefix
    in + result out in the second: 10 field elements (+ 16 bytes of entry indices in the first round); per mixed addition
    one gathered affine point + a 4-byte entry.  (With world > 1 the tail ladders over the top variables are ignored.)"""
    R = 4

    def entries(ms):
        return sum(msm_layout(m, max(ms))[1] * m for m in ms)
    out = {}

    def add(name, fq_products, nbytes, largest=0.0, largest_bytes=0.0):
        d = out.setdefault(name, {"fq_products": 0.0, "bytes": 0.0, "largest_launch_fq_products": 0.0, "largest_launch_bytes": 0.0})
        d["fq_products"] += fq_products; d["bytes"] += nbytes
        if largest > d["largest_launch_fq_products"]:
            d["largest_launch_fq_products"] = largest; d["largest_launch_bytes"] = largest_bytes

    def group(ms, g2, times):
        e = entries(ms)
        fsz = 96 if g2 else 48
        aff_p, mix_p = (17, 28) if g2 else (6, 10)
        suffix = "<Fq2>" if g2 else "<Fq>"
        if e >= (1 << 21):
            add("k_affine_round" + suffix, times * e * (1 - 2.0 ** -R) * aff_p, times * (e * (1 - 2.0 ** -R) * 10 * fsz + e / 2 * 16), e / 2 * aff_p,
                e / 2 * (10 * fsz + 16))
            add("k_seg_accum_mixed" + suffix, times * (e / 2 ** R) * mix_p, times * (e / 2 ** R) * 2 * fsz, (e / 2 ** R) * mix_p, (e / 2 ** R) * 2 * fsz)
        else:
            add("k_seg_accum_mixed" + suffix, times * e * mix_p, times * e * (2 * fsz + 4), e * mix_p, e * (2 * fsz + 4))
    group([1 << ell], False, 1)
    slots = [1 << k for k in range(ell - 1, -1, -1)]
    if ell >= 12:
        group(slots[:1], True, 1); group(slots[1:], True, 2)
    else:
        group(slots, True, 2)
    for k, d in out.items():
        d["unit"] = "Fq products: 17 per affine G2 addition (5M + 1S over Fq2), 6 over G1; 28 / 10 per mixed XYZZ addition"
    return out


def imad_roofline(kernels, work, fq_peak):
    """Achieved nominal Fq products per second of every MSM kernel (serialised CUDA-event time) against the measured ceiling."""
    out = {}
    for name, w in work.items():
        k = kernels.get(name)
        if k and k["ms_per_step"] > 0:
            g = w["fq_products"] / k["ms_per_step"] / 1e6
            out[name] = {"ms_per_step": k["ms_per_step"], "launches_per_step": k["launches_per_step"], "fq_gmul_s": g, "frac": g / fq_peak}
    tot_ms = sum(v["ms_per_step"] for v in out.values())
    if tot_ms:
        g = sum(work[k]["fq_products"] for k in out) / tot_ms / 1e6
        out["all_msm_accumulation"] = {"ms_per_step": tot_ms, "fq_gmul_s": g, "frac": g / fq_peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-log-n", type=int, default=None,
                    help="CPU baseline on a 2^K sample scaled by the operation-count model instead of one real full-size proof")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
