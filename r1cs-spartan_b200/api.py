"""Host-side mirror of the reference's prover-facing API over the C ABI (include/spartan_b200.h).

Class and method names, argument meaning and error behaviour follow the reference:

    MLArgumentForR1CS.index / prove            /root/reference/src/lib.rs:45-146
    MLProofForR1CS.prover_init ... prove_sixth_round   /root/reference/src/ahp/prover.rs:109-281
    MLPolyCommit.keygen / commit / open         /root/reference/src/commitment/{setup,commit,open}.rs
    MatrixExtension.sum_over_y / eval_on_x      /root/reference/src/data_structures/r1cs_reader.rs:75-117
    eq_extension                                /root/reference/src/data_structures/eq.rs:5-20

Everything here calls the CUDA library; there is no CPU fallback -- importing works without a GPU
(so the ABI can be inspected), but creating a Context raises.  Field elements are numpy uint64 arrays
of shape (..., 4) holding arkworks' Montgomery limbs; G1/G2 affine points are (..., 12) / (..., 24).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB_LIB_PATH") or os.path.join(_HERE, "libspartan_b200.so")   # SB_LIB_PATH: A/B builds
_lib = None

SB_OK, SB_EINVAL, SB_ECUDA, SB_ENOMEM, SB_ECOMM, SB_EINTERNAL = range(6)

EXPORTS = [
    "sb_ctx_create", "sb_ctx_create_sharded", "sb_ctx_create_multi", "sb_ctx_destroy", "sb_last_error", "sb_launch_count", "sb_device_count",
    "sb_index_create", "sb_index_destroy", "sb_index_timing", "sb_pp_load", "sb_pp_keygen", "sb_pp_export", "sb_pp_export_g_mask",
    "sb_pp_destroy", "sb_commit", "sb_open", "sb_msm", "sb_eq_table", "sb_sum_over_y", "sb_eval_on_x",
    "sb_prover_init", "sb_prover_destroy", "sb_prover_first_round", "sb_prover_second_round", "sb_prover_third_round",
    "sb_prover_first_sumcheck_round", "sb_prover_fourth_round", "sb_prover_fifth_round",
    "sb_prover_second_sumcheck_round", "sb_prover_sixth_round", "sb_prover_export_abc", "sb_phase_name", "sb_phase_span", "sb_prove",
    "sb_proof_size", "sb_field_binop", "sb_mul_bench", "sb_kernel_bench",
    "sb_witness_upload", "sb_witness_destroy", "sb_prove_resident", "sb_copy_counters", "sb_prof_enable", "sb_prof_report",
    "sb_set_serial_msm", "sb_selftest_host_field", "sb_prof_timeline", "sb_comm_shm_open", "sb_comm_shm_close", "sb_comm_shm_abort", "sb_comm_local_open",
]


class InvalidArgument(ValueError):
    """reference: Error::InvalidArgument (src/error.rs:5-14)"""


class CudaError(RuntimeError):
    pass


class CsrStruct(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p)]


COMM_ALLGATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
COMM_BARRIER = C.CFUNCTYPE(C.c_int, C.c_void_p)


class CommStruct(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("allgather", COMM_ALLGATHER), ("barrier", COMM_BARRIER), ("user", C.c_void_p)]


class TraceStruct(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("az", "bz", "cz", "sc1_evals", "sc2_evals", "r_v", "tor", "r_x", "r_abc", "r_y",
                                          "vabc", "commitment", "z_rv_0", "z_ry", "open1_proofs", "open2_proofs")] + \
               [("phase_ms", C.c_double * 16)]


def load_library():
    """Load libspartan_b200.so; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                              "There is no CPU fallback." % LIB_PATH)
        # one hardware work queue per ladder stream (the default of 8 aliases the 20+ streams of an opening onto
        # shared queues and serialises them); only effective if set before the CUDA context is created
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
        L = C.CDLL(LIB_PATH)
        L.sb_last_error.restype = C.c_char_p
        L.sb_last_error.argtypes = [C.c_void_p]
        L.sb_phase_name.restype = C.c_char_p
        L.sb_launch_count.restype = C.c_uint64
        L.sb_proof_size.restype = C.c_size_t
        L.sb_proof_size.argtypes = [C.c_uint32]
        L.sb_prof_report.restype = C.c_size_t
        L.sb_copy_counters.restype = None
        L.sb_index_timing.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _fr(a, count=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1 and a.shape[0] == 4:
        a = a.reshape(1, 4)
    if a.ndim != 2 or a.shape[1] != 4:
        raise InvalidArgument("expected Fr array of shape (k, 4)")
    if count is not None and a.shape[0] != count:
        raise InvalidArgument("expected %d field elements, got %d" % (count, a.shape[0]))
    return a


class Context:
    def __init__(self, device=0, comm=None, devices=None):
        """device: one GPU.  devices=[...]: ONE context over several GPUs of this process (sb_ctx_create_multi)."""
        L = load_library()
        self.h = C.c_void_p()
        self._comm = None
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            st = L.sb_ctx_create_multi(arr, C.c_int(len(devices)), C.byref(self.h))
        elif comm is None:
            st = L.sb_ctx_create(C.c_int(device), C.byref(self.h))
        else:
            self._comm = comm             # keeps the callback alive
            st = L.sb_ctx_create_sharded(C.c_int(device), C.byref(comm.struct), C.byref(self.h))
        if st != SB_OK:
            msg = L.sb_last_error(None).decode()
            self.h = C.c_void_p()
            raise (InvalidArgument if st == SB_EINVAL else CudaError)(msg)

    def check(self, st):
        if st == SB_OK:
            return
        msg = load_library().sb_last_error(self.h).decode()
        if st == SB_EINVAL:
            raise InvalidArgument(msg)
        if st == SB_ENOMEM:
            raise MemoryError(msg)
        raise CudaError("status %d: %s" % (st, msg))

    def close(self):
        if self.h:
            load_library().sb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- measurement hooks
    def launch_count(self):
        return int(load_library().sb_launch_count())

    def copy_counters(self):
        a, b = C.c_uint64(), C.c_uint64()
        load_library().sb_copy_counters(C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def prof_enable(self, on=True):
        load_library().sb_prof_enable(C.c_int(1 if on else 0))

    def prof_timeline(self):
        """[[kernel, start_ms, end_ms], ...] of the launches recorded since prof_enable (CUDA events)."""
        import json
        L = load_library()
        L.sb_prof_timeline.restype = C.c_size_t
        buf = C.create_string_buffer(1 << 20)
        L.sb_prof_timeline(buf, C.c_size_t(len(buf)))
        return json.loads(buf.value.decode() or "[]")

    def set_serial_msm(self, on=True):
        load_library().sb_set_serial_msm(self.h, C.c_int(1 if on else 0))

    def prof_report(self):
        """{"kernel": {"launches": n, "ms": t}} measured with CUDA events on the launching stream."""
        import json
        L = load_library()
        buf = C.create_string_buffer(1 << 16)
        L.sb_prof_report(buf, C.c_size_t(len(buf)))
        return json.loads(buf.value.decode() or "{}")

    def field_binop(self, field, op, a, b):
        words = 4 if field == "fr" else 6
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
        assert a.shape == b.shape and a.shape[-1] == words
        out = np.empty_like(a)
        opc = {"add": 0, "sub": 1, "mul": 2, "mul_portable": 3}[op]
        self.check(load_library().sb_field_binop(self.h, C.c_int(0 if field == "fr" else 1), C.c_int(opc), _p(a), _p(b), _p(out),
                                                 C.c_size_t(a.size // words)))
        return out

    def mul_bench(self, field, n_threads, iters):
        ms = C.c_double()
        self.check(load_library().sb_mul_bench(self.h, C.c_int(0 if field == "fr" else 1), C.c_size_t(n_threads), C.c_int(iters), C.byref(ms)))
        return ms.value

    def kernel_bench(self, which, log_m, reps=10, flush_l2=True):
        ms = C.c_double()
        self.check(load_library().sb_kernel_bench(self.h, C.c_int(which), C.c_uint32(log_m), C.c_int(reps), C.c_int(1 if flush_l2 else 0), C.byref(ms)))
        return ms.value


_default_ctx = None


def device_count():
    return int(load_library().sb_device_count())


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("SB_USE_LOCAL_RANK") else 0)
    return _default_ctx


# ====================================================================== data structures
def eq_extension(t, ctx=None):
    """eq(t, x) for every boolean x as ONE product table (reference returns dim separate tables whose
    pointwise product this is; eq.rs:5-20)."""
    ctx = ctx or default_context()
    t = _fr(t)
    out = np.empty((1 << t.shape[0], 4), dtype=np.uint64)
    ctx.check(load_library().sb_eq_table(ctx.h, _p(t), C.c_uint32(t.shape[0]), _p(out)))
    return out


class IndexPK:
    """reference: IndexPK (src/ahp/indexer.rs:11-17); the device copy of the three matrices."""

    def __init__(self, ctx, handle, log_n, mats):
        self.ctx, self.h, self.log_n, self.n = ctx, handle, log_n, 1 << log_n
        self.mats = mats       # host CSR kept for vk()/serialization

    def close(self):
        if self.h:
            load_library().sb_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # MatrixExtension.sum_over_y x3
    def timing(self):
        """(plan_ms, hash_wait_ms) of the index call: device-side plans incl. upload, and the transcript hash's extra wait"""
        a, b = C.c_double(), C.c_double()
        load_library().sb_index_timing(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def sum_over_y(self, z):
        z = _fr(z, self.n)
        out = [np.empty((self.n, 4), dtype=np.uint64) for _ in range(3)]
        self.ctx.check(load_library().sb_sum_over_y(self.ctx.h, self.h, _p(z), _p(out[0]), _p(out[1]), _p(out[2])))
        return out

    # MatrixExtension.eval_on_x for matrix `which`, or the r_a/r_b/r_c combination
    def eval_on_x(self, r_x, which=None, r_abc=None):
        r_x = _fr(r_x)
        if (1 << r_x.shape[0]) != self.n:
            raise InvalidArgument("2^(r_x) should have size: num_constraints")
        out = np.empty((self.n, 4), dtype=np.uint64)
        rabc = _fr(r_abc, 3) if r_abc is not None else None
        self.ctx.check(load_library().sb_eval_on_x(self.ctx.h, self.h, _p(r_x), _p(rabc), C.c_int(-1 if which is None else which), _p(out)))
        return out


class PublicParameter:
    """reference: PublicParameter (src/commitment/data_structures.rs:10-17), resident in HBM."""

    def __init__(self, ctx, handle, nv, h):
        self.ctx, self.h, self.nv, self.h_point = ctx, handle, nv, h

    def close(self):
        if self.h:
            load_library().sb_pp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self, group, level):
        words = 12 if group == 1 else 24
        out = np.empty((1 << (self.nv - level), words), dtype=np.uint64)
        self.ctx.check(load_library().sb_pp_export(self.ctx.h, self.h, C.c_int(group), C.c_uint32(level), _p(out)))
        return out

    def g_mask_random(self):
        out = np.empty((self.nv, 12), dtype=np.uint64)
        self.ctx.check(load_library().sb_pp_export_g_mask(self.ctx.h, self.h, _p(out)))
        return out


class MLPolyCommit:
    """reference: MLPolyCommit (src/commitment/mod.rs:11-14)"""

    @staticmethod
    def keygen(nv, g, h, t, keep_all_levels=False, ctx=None):
        """setup.rs:27-105 with caller-supplied generators g (G1), h (G2) and trapdoor t (nv Fr)."""
        ctx = ctx or default_context()
        g = np.ascontiguousarray(g, dtype=np.uint64).reshape(12)
        h = np.ascontiguousarray(h, dtype=np.uint64).reshape(24)
        t = _fr(t, nv)
        hd = C.c_void_p()
        ctx.check(load_library().sb_pp_keygen(ctx.h, C.c_uint32(nv), _p(g), _p(h), _p(t), C.c_int(1 if keep_all_levels else 0), C.byref(hd)))
        return PublicParameter(ctx, hd, nv, h.copy())

    @staticmethod
    def load(nv, powers_of_g0, powers_of_h, h, ctx=None):
        """Upload a reference PublicParameter: powers_of_g[0], the nv levels of powers_of_h, h."""
        ctx = ctx or default_context()
        g0 = np.ascontiguousarray(powers_of_g0, dtype=np.uint64)
        hs = [np.ascontiguousarray(x, dtype=np.uint64) for x in powers_of_h]
        if g0.shape != (1 << nv, 12) or len(hs) != nv or any(x.shape != (1 << (nv - i), 24) for i, x in enumerate(hs)):
            raise InvalidArgument("public parameter arrays have the wrong shape")
        h = np.ascontiguousarray(h, dtype=np.uint64).reshape(24)
        ptrs = (C.c_void_p * nv)(*[x.ctypes.data for x in hs])
        hd = C.c_void_p()
        ctx.check(load_library().sb_pp_load(ctx.h, C.c_uint32(nv), _p(g0), ptrs, _p(h), C.byref(hd)))
        return PublicParameter(ctx, hd, nv, h.copy())

    @staticmethod
    def commit(pp, polynomial):
        """commit.rs:17-29 -> (nv, g_product) with g_product a G1 affine point."""
        z = _fr(polynomial, 1 << pp.nv)
        out = np.empty(12, dtype=np.uint64)
        pp.ctx.check(load_library().sb_commit(pp.ctx.h, pp.h, _p(z), _p(out)))
        return pp.nv, out

    @staticmethod
    def open(pp, polynomial, point):
        """open.rs:19-58 -> (eval, (h, proofs))"""
        z = _fr(polynomial, 1 << pp.nv)
        point = _fr(point, pp.nv)
        ev = np.empty(4, dtype=np.uint64)
        proofs = np.empty((pp.nv, 24), dtype=np.uint64)
        pp.ctx.check(load_library().sb_open(pp.ctx.h, pp.h, _p(z), _p(point), _p(ev), _p(proofs)))
        return ev, (pp.h_point, proofs)


def multi_scalar_mul(group, bases, scalars, ctx=None):
    """ark_ec::msm::VariableBaseMSM::multi_scalar_mul over caller bases."""
    ctx = ctx or default_context()
    words = 12 if group == 1 else 24
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    scalars = _fr(scalars)
    n = min(bases.shape[0], scalars.shape[0])      # zips, truncating to the shorter
    out = np.empty(words, dtype=np.uint64)
    ctx.check(load_library().sb_msm(ctx.h, C.c_int(group), _p(bases), _p(scalars), C.c_size_t(n), _p(out)))
    return out


# ====================================================================== AHP rounds
class ProverState:
    def __init__(self, ctx, handle, log_n, log_v):
        self.ctx, self.h, self.log_n, self.log_v = ctx, handle, log_n, log_v

    def close(self):
        if self.h:
            load_library().sb_prover_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export_abc(self):
        n = 1 << self.log_n
        out = [np.empty((n, 4), dtype=np.uint64) for _ in range(3)]
        self.ctx.check(load_library().sb_prover_export_abc(self.h, _p(out[0]), _p(out[1]), _p(out[2])))
        return out


class MLProofForR1CS:
    """reference: MLProofForR1CS (src/ahp/mod.rs:13), prover side."""

    @staticmethod
    def index(matrix_a, matrix_b, matrix_c, ctx=None):
        """indexer.rs:41-64.  Each matrix: (row_ptr u64[n+1], col u32[nnz], val (nnz,4) u64)."""
        ctx = ctx or default_context()
        n = len(matrix_a[0]) - 1
        if n <= 0 or n & (n - 1):
            raise InvalidArgument("Matrix width should be a power of 2.")
        log_n = n.bit_length() - 1
        keep, structs = [], []
        for (rp, col, val) in (matrix_a, matrix_b, matrix_c):
            rp = np.ascontiguousarray(rp, dtype=np.uint64); col = np.ascontiguousarray(col, dtype=np.uint32)
            val = np.ascontiguousarray(val, dtype=np.uint64).reshape(-1, 4)
            if rp.shape[0] != n + 1:
                raise InvalidArgument("matrix size is inconsistent with number of constraints")
            if int(rp[-1]) != col.shape[0] or col.shape[0] != val.shape[0]:
                raise InvalidArgument("CSR arrays are inconsistent")
            keep.append((rp, col, val))
            structs.append(CsrStruct(rp.ctypes.data, col.ctypes.data if col.size else None, val.ctypes.data if val.size else None))
        hd = C.c_void_p()
        ctx.check(load_library().sb_index_create(ctx.h, C.c_uint32(log_n), C.byref(structs[0]), C.byref(structs[1]), C.byref(structs[2]), C.byref(hd)))
        return IndexPK(ctx, hd, log_n, keep)

    @staticmethod
    def prover_init(pk, v, w):
        v = _fr(v); w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
        hd = C.c_void_p()
        pk.ctx.check(load_library().sb_prover_init(pk.ctx.h, pk.h, _p(v), C.c_size_t(v.shape[0]), _p(w), C.c_size_t(w.shape[0]), C.byref(hd)))
        return ProverState(pk.ctx, hd, pk.log_n, v.shape[0].bit_length() - 1)

    @staticmethod
    def prover_first_round(state, pp):
        out = np.empty(12, dtype=np.uint64)
        state.ctx.check(load_library().sb_prover_first_round(state.h, pp.h, _p(out)))
        return state, {"commitment": (state.log_n, out)}

    @staticmethod
    def prover_second_round(state, r_v, pp):
        r_v = _fr(r_v, state.log_v) if state.log_v else np.zeros((0, 4), dtype=np.uint64)
        ev = np.empty(4, dtype=np.uint64); proofs = np.empty((state.log_n, 24), dtype=np.uint64)
        state.ctx.check(load_library().sb_prover_second_round(state.h, pp.h, _p(r_v) if state.log_v else None, _p(ev), _p(proofs)))
        return state, {"z_rv_0": ev, "proof_for_z_rv_0": (pp.h_point, proofs)}

    @staticmethod
    def prover_third_round(state, tor):
        tor = _fr(tor, state.log_n)
        state.ctx.check(load_library().sb_prover_third_round(state.h, _p(tor)))
        return state, {"ml_index_info": (state.log_n + 2, state.log_n)}

    @staticmethod
    def prove_first_sumcheck_round(state, v_msg):
        out = np.empty((state.log_n + 3, 4), dtype=np.uint64)
        vm = _fr(v_msg, 1) if v_msg is not None else None
        state.ctx.check(load_library().sb_prover_first_sumcheck_round(state.h, _p(vm), _p(out)))
        return state, out

    @staticmethod
    def prove_fourth_round(state, last_random_point):
        out = np.empty((3, 4), dtype=np.uint64)
        r = _fr(last_random_point, 1)
        state.ctx.check(load_library().sb_prover_fourth_round(state.h, _p(r), _p(out)))
        return state, {"va": out[0], "vb": out[1], "vc": out[2]}

    @staticmethod
    def prove_fifth_round(state, r_a, r_b, r_c):
        r = np.stack([_fr(r_a, 1)[0], _fr(r_b, 1)[0], _fr(r_c, 1)[0]])
        state.ctx.check(load_library().sb_prover_fifth_round(state.h, _p(r)))
        return state, {"index_info": (2, state.log_n)}

    @staticmethod
    def prove_second_sumcheck_round(state, v_msg):
        out = np.empty((3, 4), dtype=np.uint64)
        vm = _fr(v_msg, 1) if v_msg is not None else None
        state.ctx.check(load_library().sb_prover_second_sumcheck_round(state.h, _p(vm), _p(out)))
        return state, out

    @staticmethod
    def prove_sixth_round(state, last_random_point, pp):
        r = _fr(last_random_point, 1)
        ev = np.empty(4, dtype=np.uint64); proofs = np.empty((state.log_n, 24), dtype=np.uint64)
        state.ctx.check(load_library().sb_prover_sixth_round(state.h, pp.h, _p(r), _p(ev), _p(proofs)))
        return {"z_ry": ev, "proof_for_z_ry": (pp.h_point, proofs)}


class Witness:
    """z = v || w uploaded once and kept in HBM (bench.py's device-resident arm)."""

    def __init__(self, pk, v, w):
        v = _fr(v); w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
        self.ctx, self.log_v = pk.ctx, max(v.shape[0].bit_length() - 1, 0)
        self.h = C.c_void_p()
        pk.ctx.check(load_library().sb_witness_upload(pk.ctx.h, pk.h, _p(v), C.c_size_t(v.shape[0]), _p(w), C.c_size_t(w.shape[0]), C.byref(self.h)))

    def close(self):
        if self.h:
            load_library().sb_witness_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ProveTrace:
    """Every intermediate the parity tests compare (see sb_trace in the C header)."""

    def __init__(self, log_n, log_v):
        n = 1 << log_n
        f = lambda k: np.zeros((k, 4), dtype=np.uint64)
        self.az, self.bz, self.cz = f(n), f(n), f(n)
        self.sc1_evals = np.zeros((log_n, log_n + 3, 4), dtype=np.uint64)
        self.sc2_evals = np.zeros((log_n, 3, 4), dtype=np.uint64)
        self.r_v, self.tor, self.r_x, self.r_abc, self.r_y = f(max(log_v, 1)), f(log_n), f(log_n), f(3), f(log_n)
        self.vabc = f(3)
        self.commitment = np.zeros(12, dtype=np.uint64)
        self.z_rv_0, self.z_ry = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        self.open1_proofs = np.zeros((log_n, 24), dtype=np.uint64)
        self.open2_proofs = np.zeros((log_n, 24), dtype=np.uint64)
        self.phase_ms = {}
        self.log_v = log_v

    def struct(self):
        s = TraceStruct()
        for k, _ in TraceStruct._fields_[:-1]:
            setattr(s, k, getattr(self, k).ctypes.data)
        return s


class MLArgumentForR1CS:
    """reference: MLArgumentForR1CS (src/lib.rs:41-212), prover side."""

    @staticmethod
    def index(matrix_a, matrix_b, matrix_c, ctx=None):
        return MLProofForR1CS.index(matrix_a, matrix_b, matrix_c, ctx)

    @staticmethod
    def prove(pk, v, w, pp, trace=False, witness=None):
        """lib.rs:58-146.  Returns the serialized Proof bytes (and a ProveTrace when trace=True).
        witness: a Witness already resident in HBM (then v, w are ignored)."""
        L = load_library()
        cap = L.sb_proof_size(pk.log_n)
        buf = np.empty(cap, dtype=np.uint8)
        ln = C.c_size_t(cap)
        if trace == "phases":          # per-phase wall times only, no intermediate is copied out
            ts = TraceStruct()
            if witness is not None:
                st = L.sb_prove_resident(pk.ctx.h, pk.h, pp.h, witness.h, _p(buf), C.byref(ln), C.byref(ts))
            else:
                v = _fr(v); w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
                st = L.sb_prove(pk.ctx.h, pk.h, pp.h, _p(v), C.c_size_t(v.shape[0]), _p(w), C.c_size_t(w.shape[0]), _p(buf), C.byref(ln), C.byref(ts))
            pk.ctx.check(st)
            phases, i = {}, 0
            while L.sb_phase_name(i):
                phases[L.sb_phase_name(i).decode()] = ts.phase_ms[i]
                i += 1
            return buf[:ln.value].tobytes(), phases
        if witness is not None:
            tr = ProveTrace(pk.log_n, witness.log_v) if trace else None
            ts = tr.struct() if tr else None
            st = L.sb_prove_resident(pk.ctx.h, pk.h, pp.h, witness.h, _p(buf), C.byref(ln), C.byref(ts) if tr else None)
        else:
            v = _fr(v); w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
            tr = ProveTrace(pk.log_n, max(v.shape[0].bit_length() - 1, 0)) if trace else None
            ts = tr.struct() if tr else None
            st = L.sb_prove(pk.ctx.h, pk.h, pp.h, _p(v), C.c_size_t(v.shape[0]), _p(w), C.c_size_t(w.shape[0]), _p(buf), C.byref(ln),
                            C.byref(ts) if tr else None)
        pk.ctx.check(st)
        proof = buf[:ln.value].tobytes()
        if tr:
            i = 0
            while True:
                name = L.sb_phase_name(i)
                if not name:
                    break
                tr.phase_ms[name.decode()] = ts.phase_ms[i]
                i += 1
            tr.r_v = tr.r_v[:tr.log_v]
            return proof, tr
        return proof
