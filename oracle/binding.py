"""ctypes binding for oracle/liboracle.so.  TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see README.md).

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference leg.
Field elements travel as numpy uint64 arrays of shape (..., 4) (Fr) / (..., 6) (Fq): the 32/48-byte
little-endian Montgomery images arkworks keeps in memory.  G1 affine = (..., 12) uint64, G2 = (..., 24).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FR_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FQ_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("spartan_oracle.cpp", "ff.hpp", "ec.hpp", "blake2s.hpp", "pairing.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        L = _LIB
        L.or_r1cs_synth.restype = C.c_void_p
        L.or_r1cs_synth.argtypes = [C.c_size_t, C.c_size_t, C.c_uint, C.c_uint64]
        L.or_r1cs_from_csr.restype = C.c_void_p
        L.or_keygen.restype = C.c_void_p
        L.or_keygen.argtypes = [C.c_size_t, C.c_uint64]
        L.or_keygen_with.restype = C.c_void_p
        L.or_pp_from_arrays.restype = C.c_void_p
        L.or_vp_from_arrays.restype = C.c_void_p
        L.or_prove.restype = C.c_void_p
        L.or_fs_new.restype = C.c_void_p
        for f in ("or_r1cs_n", "or_r1cs_num_public", "or_r1cs_nnz", "or_trace_get"):
            getattr(L, f).restype = C.c_size_t
        L.or_trace_time.restype = C.c_double
        L.or_init()
    return _LIB


def set_threads(k):
    """TESTS ONLY: worker threads of the oracle's heavy loops (identical results for any k; see ec.hpp).  A CPU-baseline
    timing must leave this at 1, the reference's configuration.  Returns the previous value."""
    return int(lib().or_set_threads(C.c_int(int(k))))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, words=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if words is not None:
        assert a.shape[-1] == words, a.shape
    return a


# ---------------------------------------------------------------- integer <-> limb helpers
def ints_to_limbs(vals, words):
    out = np.zeros((len(vals), words), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(words):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, a.shape[-1])
    return [sum(int(x) << (64 * j) for j, x in enumerate(row)) for row in a]


def fr_from_ints(vals):
    """canonical python ints -> Montgomery Fr array"""
    c = ints_to_limbs([v % FR_MOD for v in vals], 4)
    out = np.empty_like(c)
    lib().or_fr_from_canonical(_p(c), _p(out), C.c_size_t(len(vals)))
    return out


def fr_to_ints(a):
    a = _c(a, 4).reshape(-1, 4)
    out = np.empty_like(a)
    lib().or_fr_to_canonical(_p(a), _p(out), C.c_size_t(a.shape[0]))
    return limbs_to_ints(out)


def fq_from_ints(vals):
    c = ints_to_limbs([v % FQ_MOD for v in vals], 6)
    out = np.empty_like(c)
    lib().or_fq_from_canonical(_p(c), _p(out), C.c_size_t(len(vals)))
    return out


def fq_to_ints(a):
    a = _c(a, 6).reshape(-1, 6)
    out = np.empty_like(a)
    lib().or_fq_to_canonical(_p(a), _p(out), C.c_size_t(a.shape[0]))
    return limbs_to_ints(out)


def fr_binop(op, a, b):
    a = _c(a, 4); b = _c(b, 4)
    out = np.empty_like(a)
    lib().or_fr_binop(C.c_int({"add": 0, "sub": 1, "mul": 2}[op]), _p(a), _p(b), _p(out), C.c_size_t(a.size // 4))
    return out


def fq_binop(op, a, b):
    a = _c(a, 6); b = _c(b, 6)
    out = np.empty_like(a)
    lib().or_fq_binop(C.c_int({"add": 0, "sub": 1, "mul": 2}[op]), _p(a), _p(b), _p(out), C.c_size_t(a.size // 6))
    return out


def fr_rand(seed, n):
    out = np.empty((n, 4), dtype=np.uint64)
    lib().or_fr_rand(C.c_uint64(seed), _p(out), C.c_size_t(n))
    return out


def g1_to_py(a):
    """(12,) uint64 Montgomery affine -> python (x, y) or None"""
    a = _c(a).reshape(12)
    if not a.any():
        return None
    x, y = fq_to_ints(a.reshape(2, 6))
    return (x, y)


def g2_to_py(a):
    a = _c(a).reshape(24)
    if not a.any():
        return None
    c = fq_to_ints(a.reshape(4, 6))
    return ((c[0], c[1]), (c[2], c[3]))


def g1_from_py(p):
    if p is None:
        return np.zeros(12, dtype=np.uint64)
    return fq_from_ints([p[0], p[1]]).reshape(12)


def g2_from_py(p):
    if p is None:
        return np.zeros(24, dtype=np.uint64)
    return fq_from_ints([p[0][0], p[0][1], p[1][0], p[1][1]]).reshape(24)


def generators():
    g = np.empty(12, dtype=np.uint64); h = np.empty(24, dtype=np.uint64)
    lib().or_generators(_p(g), _p(h))
    return g, h


def g1_mul(p, k):
    out = np.empty(12, dtype=np.uint64); p = _c(p); k = _c(k)
    lib().or_g1_mul(_p(p), _p(k), _p(out)); return out


def g2_mul(p, k):
    out = np.empty(24, dtype=np.uint64); p = _c(p); k = _c(k)
    lib().or_g2_mul(_p(p), _p(k), _p(out)); return out


def g1_add(a, b):
    out = np.empty(12, dtype=np.uint64); a = _c(a); b = _c(b)
    lib().or_g1_add(_p(a), _p(b), _p(out)); return out


def g2_add(a, b):
    out = np.empty(24, dtype=np.uint64); a = _c(a); b = _c(b)
    lib().or_g2_add(_p(a), _p(b), _p(out)); return out


def ser_g1(p):
    out = np.empty(48, dtype=np.uint8); p = _c(p)
    lib().or_ser_g1(_p(p), _p(out)); return out.tobytes()


def ser_g2(p):
    out = np.empty(96, dtype=np.uint8); p = _c(p)
    lib().or_ser_g2(_p(p), _p(out)); return out.tobytes()


def msm_g1(bases, scalars):
    bases = _c(bases, 12); scalars = _c(scalars, 4); out = np.empty(12, dtype=np.uint64)
    lib().or_msm_g1(_p(bases), _p(scalars), C.c_size_t(scalars.shape[0]), _p(out)); return out


def msm_g2(bases, scalars):
    bases = _c(bases, 24); scalars = _c(scalars, 4); out = np.empty(24, dtype=np.uint64)
    lib().or_msm_g2(_p(bases), _p(scalars), C.c_size_t(scalars.shape[0]), _p(out)); return out


class FsRng:
    def __init__(self): self.h = C.c_void_p(lib().or_fs_new())
    def __del__(self):
        try: lib().or_fs_free(self.h)
        except Exception: pass
    def feed(self, data: bytes):
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        lib().or_fs_feed(self.h, buf, C.c_size_t(len(data)))
    def fill(self, n):
        buf = (C.c_uint8 * n)()
        lib().or_fs_fill(self.h, buf, C.c_size_t(n)); return bytes(buf)
    def fr_rand(self, n):
        out = np.empty((n, 4), dtype=np.uint64)
        lib().or_fs_fr_rand(self.h, _p(out), C.c_size_t(n)); return out


class R1CS:
    """Synthetic benchmark circuit (constraints.rs:39-110 + test_utils.rs:51-102) or caller CSR."""
    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        self.n = lib().or_r1cs_n(self.h)
        self.log_n = self.n.bit_length() - 1

    @classmethod
    def synth(cls, num_public, num_private, density=0, seed=0):
        return cls(lib().or_r1cs_synth(num_public, num_private, density, seed))

    @classmethod
    def from_csr(cls, log_n, mats):
        """mats: 3 x (rowptr u64[n+1], col u32[nnz], val (nnz,4) u64)"""
        keep = []
        rp = (C.c_void_p * 3)(); cl = (C.c_void_p * 3)(); vl = (C.c_void_p * 3)()
        for k, (r, c, v) in enumerate(mats):
            r = np.ascontiguousarray(r, dtype=np.uint64); c = np.ascontiguousarray(c, dtype=np.uint32)
            v = np.ascontiguousarray(v, dtype=np.uint64)
            keep += [r, c, v]
            rp[k] = r.ctypes.data; cl[k] = c.ctypes.data; vl[k] = v.ctypes.data
        return cls(lib().or_r1cs_from_csr(C.c_size_t(log_n), rp, cl, vl))

    def __del__(self):
        try: lib().or_r1cs_free(self.h)
        except Exception: pass

    def csr(self, which):
        nnz = lib().or_r1cs_nnz(self.h, C.c_int(which))
        rowptr = np.empty(self.n + 1, dtype=np.uint64); col = np.empty(nnz, dtype=np.uint32)
        val = np.empty((nnz, 4), dtype=np.uint64)
        lib().or_r1cs_export(self.h, C.c_int(which), _p(rowptr), _p(col), _p(val))
        return rowptr, col, val

    def vw(self):
        nv = lib().or_r1cs_num_public(self.h)
        v = np.empty((nv, 4), dtype=np.uint64); w = np.empty((self.n - nv, 4), dtype=np.uint64)
        lib().or_r1cs_vw(self.h, _p(v), _p(w)); return v, w

    def is_satisfied(self, z):
        z = _c(z, 4); return bool(lib().or_r1cs_is_satisfied(self.h, _p(z)))

    def sum_over_y(self, which, z):
        z = _c(z, 4); out = np.empty((self.n, 4), dtype=np.uint64)
        lib().or_sum_over_y(self.h, C.c_int(which), _p(z), _p(out)); return out

    def eval_on_x(self, which, r_x):
        r_x = _c(r_x, 4); out = np.empty((self.n, 4), dtype=np.uint64)
        lib().or_eval_on_x(self.h, C.c_int(which), _p(r_x), _p(out)); return out

    def rows_py(self, which):
        """python-int rows [(coeff, col), ...] for the big-int model"""
        rowptr, col, val = self.csr(which)
        vals = fr_to_ints(val) if len(col) else []
        return [[(vals[e], int(col[e])) for e in range(int(rowptr[r]), int(rowptr[r + 1]))] for r in range(self.n)]


def eq_extension(t):
    t = _c(t, 4); dim = t.shape[0]
    out = np.empty((dim, 1 << dim, 4), dtype=np.uint64)
    lib().or_eq_extension(_p(t), C.c_size_t(dim), _p(out)); return out


def mle_eval(table, point):
    table = _c(table, 4); point = _c(point, 4); out = np.empty(4, dtype=np.uint64)
    lib().or_mle_eval(_p(table), C.c_size_t(point.shape[0]), _p(point), _p(out)); return out


def sumcheck_prove(tables, products, challenges):
    """tables: (ntab, 2^nv, 4); products: list of lists of table indices; challenges (nv, 4).
    Returns (nv, max_mult + 1, 4) round evaluations of the literal upstream prover."""
    tables = _c(tables, 4); challenges = _c(challenges, 4)
    ntab, n = tables.shape[0], tables.shape[1]; nv = n.bit_length() - 1
    sizes = np.array([len(p) for p in products], dtype=np.uint32)
    idx = np.array([i for p in products for i in p], dtype=np.uint32)
    d = int(sizes.max()) + 1
    out = np.empty((nv, d, 4), dtype=np.uint64)
    lib().or_sumcheck_prove(_p(tables), C.c_size_t(ntab), C.c_size_t(nv), _p(sizes), C.c_size_t(len(products)),
                            _p(idx), _p(challenges), _p(out))
    return out


class PP:
    def __init__(self, handle, nv):
        self.h = C.c_void_p(handle); self.nv = nv

    @classmethod
    def keygen(cls, nv, seed):
        return cls(lib().or_keygen(nv, seed), nv)

    @classmethod
    def keygen_with(cls, nv, g, h, t):
        g = _c(g); h = _c(h); t = _c(t, 4)
        return cls(lib().or_keygen_with(C.c_size_t(nv), _p(g), _p(h), _p(t)), nv)

    @classmethod
    def verifier_only(cls, nv, g, h, g_mask):
        """VerifierParameter {nv, g, h, g_mask_random} (data_structures.rs:20-25)"""
        g = _c(g); h = _c(h); g_mask = _c(g_mask, 12)
        return cls(lib().or_vp_from_arrays(C.c_size_t(nv), _p(g), _p(h), _p(g_mask)), nv)

    @classmethod
    def from_arrays(cls, nv, g1_level0, g2_all, h):
        g1 = _c(g1_level0, 12); g2 = _c(g2_all, 24); h = _c(h)
        return cls(lib().or_pp_from_arrays(C.c_size_t(nv), _p(g1), _p(g2), _p(h)), nv)

    def __del__(self):
        try: lib().or_pp_free(self.h)
        except Exception: pass

    def g1(self, level):
        out = np.empty((1 << (self.nv - level), 12), dtype=np.uint64)
        lib().or_pp_export_g1(self.h, C.c_size_t(level), _p(out)); return out

    def g2(self, level):
        out = np.empty((1 << (self.nv - level), 24), dtype=np.uint64)
        lib().or_pp_export_g2(self.h, C.c_size_t(level), _p(out)); return out

    def g2_all(self):
        return np.concatenate([self.g2(i) for i in range(self.nv)], axis=0)

    def gh(self):
        g = np.empty(12, dtype=np.uint64); h = np.empty(24, dtype=np.uint64)
        lib().or_pp_gh(self.h, _p(g), _p(h)); return g, h

    def trapdoor(self):
        out = np.empty((self.nv, 4), dtype=np.uint64)
        lib().or_pp_trapdoor(self.h, _p(out)); return out

    def g_mask(self):
        out = np.empty((self.nv, 12), dtype=np.uint64)
        lib().or_pp_g_mask(self.h, _p(out)); return out

    def commit(self, z):
        z = _c(z, 4); out = np.empty(12, dtype=np.uint64)
        lib().or_commit(self.h, _p(z), C.c_size_t(z.shape[0]), _p(out)); return out

    def open(self, z, point, want_q=False):
        z = _c(z, 4); point = _c(point, 4); nv = point.shape[0]
        ev = np.empty(4, dtype=np.uint64); proofs = np.empty((nv, 24), dtype=np.uint64)
        q = np.empty(((1 << nv) - 1, 4), dtype=np.uint64) if want_q else None
        lib().or_open(self.h, _p(z), _p(point), C.c_size_t(nv), _p(ev), _p(proofs), _p(q) if want_q else None)
        return (ev, proofs, q) if want_q else (ev, proofs)


def verify(r1cs, vp, v, proof):
    """MLArgumentForR1CS::verify (lib.rs:147-212): 1 accept, 0 malformed, < 0 the failed check (see oracle source)."""
    v = _c(v, 4)
    buf = (C.c_uint8 * len(proof)).from_buffer_copy(proof)
    return int(lib().or_verify(r1cs.h, vp.h, _p(v), C.c_size_t(v.shape[0]), buf, C.c_size_t(len(proof))))


def pc_verify(vp, commitment, point, ev, proofs):
    """MLPolyCommit::verify (commitment/verify.rs:12-45)"""
    commitment = _c(commitment); point = _c(point, 4); ev = _c(ev); proofs = _c(proofs, 24)
    return bool(lib().or_pc_verify(vp.h, _p(commitment), _p(point), _p(ev), _p(proofs)))


def pairing_check(a, b):
    a = _c(a); b = _c(b)
    return bool(lib().or_pairing_check(_p(a), _p(b)))


def deser_g1(data):
    out = np.empty(12, dtype=np.uint64); buf = (C.c_uint8 * 48).from_buffer_copy(data)
    assert lib().or_deser_g1(buf, _p(out)); return out


def deser_g2(data):
    out = np.empty(24, dtype=np.uint64); buf = (C.c_uint8 * 96).from_buffer_copy(data)
    assert lib().or_deser_g2(buf, _p(out)); return out


class Trace:
    def __init__(self, handle): self.h = C.c_void_p(handle)
    def __del__(self):
        try: lib().or_trace_free(self.h)
        except Exception: pass
    def blob(self, name):
        ptr = C.POINTER(C.c_uint8)()
        n = lib().or_trace_get(self.h, name.encode(), C.byref(ptr))
        return bytes(C.string_at(ptr, n)) if n else b""
    def fr(self, name):
        b = self.blob(name)
        return np.frombuffer(b, dtype=np.uint64).reshape(-1, 4).copy()
    def time(self, name): return lib().or_trace_time(self.h, name.encode())


def prove(r1cs, pp, v, w):
    """The literal NI prover (lib.rs:58-146). Returns (proof_bytes, Trace)."""
    v = _c(v, 4); w = _c(w, 4); st = C.c_int(0)
    t = Trace(lib().or_prove(r1cs.h, pp.h, _p(v), C.c_size_t(v.shape[0]), _p(w), C.c_size_t(w.shape[0]), C.byref(st)))
    if st.value != 0:
        raise ValueError("InvalidArgument")
    return t.blob("proof"), t
