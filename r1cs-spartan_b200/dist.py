"""One-process-per-GPU plumbing for the hypercube-sharded prover.

The library needs exactly one collective -- an allgather of a few hundred bytes (three field elements per
sumcheck round, one partial group element per MSM) -- and takes it as a C callback (`sb_comm` in
include/spartan_b200.h).  This module backs that callback with `torch.distributed`: NCCL over
NVLink/NVSwitch on the GPU box (one rank per GPU, launched by torchrun), gloo in the CPU tests.
The payloads are latency-bound (96 bytes per round), so the link bandwidth is irrelevant here.
"""
import ctypes as C

import numpy as np

from .api import COMM_ALLGATHER, COMM_BARRIER, CommStruct, Context


class TorchComm:
    """sb_comm backed by the default torch.distributed process group."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world & (self.world - 1):
            raise ValueError("the hypercube is split on its top variables: world size must be a power of two")
        self.backend = dist.get_backend()
        self.torch, self.dist = torch, dist
        self.device = torch.device("cuda", device if device is not None else torch.cuda.current_device()) \
            if self.backend == "nccl" else torch.device("cpu")
        self.calls = 0
        self.bytes = 0
        self.cb = COMM_ALLGATHER(self._allgather)
        self.struct = CommStruct(self.rank, self.world, self.cb, COMM_BARRIER(), None)   # no barrier hook: the collective itself orders the ranks

    def _allgather(self, user, send, recv, nbytes):
        try:
            torch, dist = self.torch, self.dist
            src = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(send)).copy()
            t = torch.from_numpy(src).to(self.device)
            if self.backend == "nccl":
                out = torch.empty(self.world * nbytes, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(out, t)
                host = out.cpu().numpy()
            else:
                parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(self.world)]
                dist.all_gather(parts, t)
                host = torch.cat(parts).numpy()
            C.memmove(recv, host.ctypes.data, self.world * nbytes)
            self.calls += 1
            self.bytes += nbytes
            return 0
        except Exception as e:      # never let an exception cross the C boundary
            import sys
            print("allgather hook failed: %r" % (e,), file=sys.stderr)
            return 1


class ShmComm:
    """sb_comm over a POSIX shared-memory mailbox (csrc/comm_shm.cu): all ranks are on one node, the payloads
    are tiny and already on the host, so this is ~50x lower latency than a device collective.  The process
    group is only used to agree on the mailbox name and to order creation before attachment."""

    def __init__(self):
        import os
        import torch.distributed as dist
        from .api import load_library
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world & (self.world - 1):
            raise ValueError("the hypercube is split on its top variables: world size must be a power of two")
        names = [None]
        if self.rank == 0:
            names[0] = "/sb_b200_%d_%s" % (os.getpid(), os.environ.get("MASTER_PORT", "0"))
        dist.broadcast_object_list(names, src=0)
        self.name = names[0]
        self.struct = CommStruct()
        L = load_library()
        if self.rank == 0:
            st = L.sb_comm_shm_open(self.name.encode(), self.rank, self.world, 1, C.byref(self.struct))
            if st != 0:
                raise RuntimeError("cannot create the shared-memory mailbox %s" % self.name)
        dist.barrier()
        if self.rank != 0:
            st = L.sb_comm_shm_open(self.name.encode(), self.rank, self.world, 0, C.byref(self.struct))
            if st != 0:
                raise RuntimeError("cannot attach to the shared-memory mailbox %s" % self.name)
        dist.barrier()

    @property
    def cb(self):
        return self.struct.allgather

    def close(self):
        from .api import load_library
        load_library().sb_comm_shm_close(C.byref(self.struct))


def sharded_context(device=0, comm=None):
    """Context for this rank of the default process group (a plain Context when world size is 1).
    comm: "shm" (default; single node) or "nccl"/"torch" (torch.distributed allgather); env SB_COMM overrides."""
    import os
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return Context(device)
    kind = os.environ.get("SB_COMM") or comm or "shm"
    return Context(device, comm=ShmComm() if kind == "shm" else TorchComm(device))


def slice_weight_mont(t_hi_mont, rho):
    """eq(t_hi, rho) as a Montgomery residue: the weight of slice rho in any eq table of t (host mirror of
    `top_weight` in csrc/prover.cu, used by the CPU tests of the sharding identities)."""
    from .workload import FR_MOD
    R = 1 << 256
    rinv = pow(R, -1, FR_MOD)
    w = R % FR_MOD
    for k, t in enumerate(t_hi_mont):
        f = t if (rho >> k) & 1 else (R % FR_MOD - t) % FR_MOD
        w = w * f % FR_MOD * rinv % FR_MOD
    return w
