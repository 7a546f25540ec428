"""CPU tests of the multi-rank host logic (world size 2 and 4, gloo): the allgather hook the C library calls,
and the algebraic identities the hypercube sharding rests on, checked with the oracle per rank."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import r1cs_spartan_b200 as sb
        from r1cs_spartan_b200 import dist as sbdist
        from r1cs_spartan_b200.workload import limbs_to_mont, mont_to_limbs
        from oracle import binding as ob

        # 1. the hooks, exactly as the C library calls them: torch.distributed (gloo here) and shared memory
        shm = sbdist.ShmComm()
        for rep in range(50):
            nb = 96 + 8 * (rep % 5)
            send = (C.c_uint8 * nb)(*[(rank * 17 + rep + i) % 251 for i in range(nb)])
            recv = (C.c_uint8 * (nb * world))()
            assert shm.cb(shm.struct.user, C.addressof(send), C.addressof(recv), nb) == 0
            got = bytes(recv)
            for r in range(world):
                assert got[r * nb:(r + 1) * nb] == bytes((r * 17 + rep + i) % 251 for i in range(nb)), (rep, r)
        # a rank that failed locally releases its peers: they fail the exchange at once (not after the time-out), and the
        # next epoch (the barrier every library call starts with) re-aligns everybody
        import time
        L = sb.load_library()
        assert shm.struct.barrier(shm.struct.user) == 0
        nb = 96
        send = (C.c_uint8 * nb)(*[rank] * nb); recv = (C.c_uint8 * (nb * world))()
        t0 = time.perf_counter()
        if rank == world - 1:
            L.sb_comm_shm_abort(C.byref(shm.struct))
        else:
            assert shm.cb(shm.struct.user, C.addressof(send), C.addressof(recv), nb) != 0
        assert time.perf_counter() - t0 < 30.0
        dist.barrier()
        assert shm.struct.barrier(shm.struct.user) == 0
        assert shm.cb(shm.struct.user, C.addressof(send), C.addressof(recv), nb) == 0
        assert bytes(recv) == b"".join(bytes([r]) * nb for r in range(world))
        shm.close()
        comm = sbdist.TorchComm()
        assert (comm.rank, comm.world) == (rank, world)
        for nbytes in (96, 8 * 384 + 32):
            send = (C.c_uint8 * nbytes)(*[(rank * 31 + i) % 251 for i in range(nbytes)])
            recv = (C.c_uint8 * (nbytes * world))()
            assert comm.cb(None, C.addressof(send), C.addressof(recv), nbytes) == 0
            got = bytes(recv)
            for r in range(world):
                assert got[r * nbytes:(r + 1) * nbytes] == bytes((r * 31 + i) % 251 for i in range(nbytes))

        # 2. slice identities (every rank checks its own slice against the global objects)
        glog = world.bit_length() - 1
        nv = 5
        loc = nv - glog
        g, h = ob.generators()
        t = ob.fr_rand(4321, nv)
        gpp = ob.PP.keygen_with(nv, g, h, t)
        w = mont_to_limbs([sbdist.slice_weight_mont(limbs_to_mont(t[loc:]), rank)])[0]
        lpp = ob.PP.keygen_with(loc, ob.g1_mul(g, w), ob.g2_mul(h, w), t[:loc])
        nl = 1 << loc
        #   the rank's slice of a PublicParameter is the PublicParameter of the low variables with scaled generators
        assert np.array_equal(lpp.g1(0), gpp.g1(0)[rank * nl:(rank + 1) * nl])
        for L in range(1, loc):
            sz = 1 << (loc - L)
            assert np.array_equal(lpp.g2(L), gpp.g2(L)[rank * sz:(rank + 1) * sz]), L
        assert np.array_equal(ob.g2_mul(h, w), gpp.g2(loc)[rank])          # last local base
        #   eq(tau, x) on the slice = eq(tau_hi, rho) * eq(tau_low, x_low)
        tabs = ob.eq_extension(t)
        full = tabs[0]
        for i in range(1, nv):
            full = ob.fr_binop("mul", full, tabs[i])
        ltabs = ob.eq_extension(t[:loc])
        low = ltabs[0]
        for i in range(1, loc):
            low = ob.fr_binop("mul", low, ltabs[i])
        assert np.array_equal(ob.fr_binop("mul", low, np.repeat(w.reshape(1, 4), nl, axis=0)), full[rank * nl:(rank + 1) * nl])
        #   commitment = sum of the per-slice commitments (exchanged through the hook)
        z = ob.fr_rand(77, 1 << nv)
        part = lpp.commit(z[rank * nl:(rank + 1) * nl])
        send = np.ascontiguousarray(part)
        recv = np.zeros((world, 12), dtype=np.uint64)
        assert comm.cb(None, send.ctypes.data, recv.ctypes.data, 96) == 0
        acc = recv[0]
        for r in range(1, world):
            acc = ob.g1_add(acc, recv[r])
        assert np.array_equal(acc, gpp.commit(z))
        #   opening: local proofs sum to the global proofs for the low levels; folded values form the tail table
        point = ob.fr_rand(78, nv)
        gev, gproofs = gpp.open(z, point)
        lev, lproofs = lpp.open(z[rank * nl:(rank + 1) * nl], point[:loc])
        rec = np.concatenate([lproofs.reshape(-1), lev])
        allr = np.zeros((world, rec.size), dtype=np.uint64)
        assert comm.cb(None, np.ascontiguousarray(rec).ctypes.data, allr.ctypes.data, rec.size * 8) == 0
        for i in range(loc):
            acc = allr[0, 24 * i:24 * (i + 1)]
            for r in range(1, world):
                acc = ob.g2_add(acc, allr[r, 24 * i:24 * (i + 1)])
            assert np.array_equal(acc, gproofs[i]), i
        tail_table = allr[:, 24 * loc:]
        tpp = ob.PP.keygen_with(glog, g, h, t[loc:])
        tev, tproofs = tpp.open(tail_table, point[loc:])
        assert np.array_equal(tev, gev) and np.array_equal(tproofs, gproofs[loc:])
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:          # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 4])
def test_sharding_identities_and_allgather_hook_gloo(world, oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results
