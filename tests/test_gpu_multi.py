"""Multi-GPU parity: the hypercube-sharded prover (one process per GPU, NCCL) must produce the same proof
bytes as the single-threaded CPU oracle, on every rank.  Needs >= 2 GPUs (gpurun --gpus 2/4/8)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, cases, q, backend="nccl"):
    # backend "gloo" = the CUDA-on-CPU emulation build (tests/test_emul_cpu.py): same product code, host threads for kernels
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        if backend == "nccl":
            torch.cuda.set_device(rank)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        import r1cs_spartan_b200 as sb
        from r1cs_spartan_b200 import dist as sbdist
        from oracle import binding as ob
        ctx = sbdist.sharded_context(rank)
        g, h = ob.generators()
        for (log_n, num_public, density, use_load) in cases:
            cs = sb.SyntheticR1CS(num_public, (1 << log_n) - num_public, density, 0x5EED0000 + log_n)
            t = ob.fr_rand(1234 + log_n, log_n)
            opp = ob.PP.keygen_with(log_n, g, h, t)
            if use_load:      # slice an existing reference PublicParameter
                pp = sb.MLPolyCommit.load(log_n, opp.g1(0), [opp.g2(i) for i in range(log_n)], h, ctx=ctx)
            else:             # sharded keygen: every rank builds its own slice
                pp = sb.MLPolyCommit.keygen(log_n, g, h, t, ctx=ctx)
            pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
            proof, tr = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp, trace=True)
            ocs = ob.R1CS.from_csr(log_n, cs.mats)
            oproof, otr = ob.prove(ocs, opp, cs.v, cs.w)
            nl = (1 << log_n) // world
            sl = slice(rank * nl, (rank + 1) * nl)
            assert np.array_equal(tr.commitment, np.frombuffer(otr.blob("commitment"), dtype=np.uint64)), "commitment"
            assert np.array_equal(tr.open1_proofs.reshape(-1), np.frombuffer(otr.blob("open1_proofs"), dtype=np.uint64)), "open1"
            assert np.array_equal(tr.az[sl], otr.fr("az")[sl]) and np.array_equal(tr.cz[sl], otr.fr("cz")[sl]), "Az/Cz slice"
            assert np.array_equal(tr.sc1_evals.reshape(-1, 4), otr.fr("sc1_evals")), "sumcheck 1"
            assert np.array_equal(tr.vabc, otr.fr("vabc")), "va vb vc"
            assert np.array_equal(tr.sc2_evals.reshape(-1, 4), otr.fr("sc2_evals")), "sumcheck 2"
            assert proof == oproof, "proof bytes (log_n=%d)" % log_n
            # commit / open entry points on the sharded context
            z = np.concatenate([cs.v, cs.w])
            assert np.array_equal(sb.MLPolyCommit.commit(pp, z)[1], opp.commit(z))
            point = ob.fr_rand(5, log_n)
            ev, (_, proofs) = sb.MLPolyCommit.open(pp, z, point)
            oev, oproofs = opp.open(z, point)
            assert np.array_equal(ev, oev) and np.array_equal(proofs, oproofs)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


def _run(world, cases, backend="nccl"):
    import torch
    if backend == "nccl" and torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q, backend)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_sharded_prove_matches_oracle_2gpu(oracle):
    _run(2, [(3, 4, 0, False), (6, 8, 30, True), (10, 32, 0, False)])


def test_sharded_prove_matches_oracle_4gpu(oracle):
    _run(4, [(4, 4, 0, True), (10, 32, 0, False)])


def test_sharded_prove_matches_oracle_8gpu(oracle):
    _run(8, [(5, 8, 0, False), (10, 32, 0, False)])


# ---------------------------------------------------------------------------------------------------------------------
# ONE process, ONE context over several GPUs (sb_ctx_create_multi): MLArgumentForR1CS::prove stays one call (lib.rs:58).
def _multi_context_cases(ndev, cases):
    import r1cs_spartan_b200 as sb
    from oracle import binding as ob
    if sb.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    ctx = sb.Context(devices=list(range(ndev)))
    g, h = ob.generators()
    if os.environ.get("SB_EMUL_TESTS"):          # CUDA-on-CPU shim: host threads for CUDA threads -- keep the instances tiny
        cases = [c for c in cases if c[0] <= 6][:1]
    for (log_n, num_public, density, use_load) in cases:
        cs = sb.SyntheticR1CS(num_public, (1 << log_n) - num_public, density, 0x5EED0000 + log_n)
        t = ob.fr_rand(1234 + log_n, log_n)
        opp = ob.PP.keygen_with(log_n, g, h, t)
        if use_load:
            pp = sb.MLPolyCommit.load(log_n, opp.g1(0), [opp.g2(i) for i in range(log_n)], h, ctx=ctx)
        else:
            pp = sb.MLPolyCommit.keygen(log_n, g, h, t, ctx=ctx)
            assert np.array_equal(pp.g_mask_random(), opp.g_mask())        # vp.g_mask_random = g^{t_i} (setup.rs:92-94), host data of the multi context
        pk = sb.MLArgumentForR1CS.index(*cs.mats, ctx=ctx)
        ocs = ob.R1CS.from_csr(log_n, cs.mats)
        oproof, otr = ob.prove(ocs, opp, cs.v, cs.w)
        proof, tr = sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp, trace=True)
        assert proof == oproof, "proof bytes (log_n=%d)" % log_n
        assert np.array_equal(tr.commitment, np.frombuffer(otr.blob("commitment"), dtype=np.uint64))
        assert np.array_equal(tr.az, otr.fr("az")) and np.array_equal(tr.bz, otr.fr("bz")) and np.array_equal(tr.cz, otr.fr("cz"))
        assert np.array_equal(tr.sc1_evals.reshape(-1, 4), otr.fr("sc1_evals"))
        assert np.array_equal(tr.sc2_evals.reshape(-1, 4), otr.fr("sc2_evals"))
        assert np.array_equal(tr.vabc, otr.fr("vabc"))
        # a resident witness and a second proof on the same handles
        wit = sb.Witness(pk, cs.v, cs.w)
        assert sb.MLArgumentForR1CS.prove(pk, None, None, pp, witness=wit) == oproof
        # commit / open
        z = np.concatenate([cs.v, cs.w])
        assert np.array_equal(sb.MLPolyCommit.commit(pp, z)[1], opp.commit(z))
        point = ob.fr_rand(5, log_n)
        ev, (_, proofs) = sb.MLPolyCommit.open(pp, z, point)
        oev, oproofs = opp.open(z, point)
        assert np.array_equal(ev, oev) and np.array_equal(proofs, oproofs)
        # invalid arguments keep the reference's error mapping on a multi context (prover.rs:117-119, r1cs_reader.rs:55-62)
        with pytest.raises(sb.InvalidArgument):
            sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w[:-1], pp)
        rp, col, val = cs.mats[0]
        bad_col = col.copy(); bad_col[0] = 1 << log_n
        with pytest.raises(sb.InvalidArgument):
            sb.MLArgumentForR1CS.index((rp, bad_col, val), cs.mats[1], cs.mats[2], ctx=ctx)
        # and the context still works after a failed call (no rank is left behind in the exchange)
        assert sb.MLArgumentForR1CS.prove(pk, cs.v, cs.w, pp) == oproof
        wit.close(); pk.close(); pp.close()
    ctx.close()


def test_multi_context_one_process_2gpu(oracle):
    _multi_context_cases(2, [(3, 4, 0, False), (6, 8, 30, True), (10, 32, 0, False)])


def test_multi_context_one_process_4gpu(oracle):
    _multi_context_cases(4, [(4, 4, 0, True), (9, 32, 0, False)])


def test_multi_context_one_process_8gpu(oracle):
    _multi_context_cases(8, [(5, 8, 0, False), (10, 32, 0, False)])
