"""Wire formats of the reference's artefacts: the byte layouts that the
`#[derive(CanonicalSerialize, CanonicalDeserialize)]` types of the reference produce, read and written from the
array layouts this package's API uses, so that keys, indices and proofs can be exchanged with an arkworks build.

    Proof                src/data_structures/proof.rs:11-20  (messages: src/ahp/prover.rs:66-101)
    Commitment           src/commitment/commit.rs:11-14
    open::Proof          src/commitment/open.rs:12-15
    IndexPK / IndexVK    src/ahp/indexer.rs:10-27            (MatrixExtension: src/data_structures/r1cs_reader.rs:8-13)
    PublicParameter      src/commitment/data_structures.rs:9-17
    VerifierParameter    src/commitment/data_structures.rs:19-25
    key cache            src/commitment/mod.rs:48-62 (Vec<ParameterPair>, serialize_uncompressed)

The encodings themselves are arkworks' (`ark-serialize`, un-vendored and un-pinned in the reference, see
SURVEY.md Appendix A.5) and are restated here from their published description -- UNVERIFIED against a real
arkworks binary, like the rest of the parity chain:
  * usize / u64: 8 bytes little-endian;  Vec<T>: u64 length, then the items;  structs: fields in order
  * Fr: 32 bytes little-endian of the canonical (non-Montgomery) integer
  * G1Affine compressed: x (48 B LE), flags in the top bits of the last byte: bit 7 = y is the larger of
    {y, -y}, bit 6 = point at infinity;  uncompressed: x then y (48 B each), only the infinity flag, on y
  * G2Affine: the same over Fq2, an Fq2 element being c0 then c1; Fq2 elements are ordered by c1, then c0

Array conventions (the ones api.py uses): Fr = 4 x u64 little-endian limbs in Montgomery form (R = 2^256);
a G1 point = 12 x u64 (x, y in Montgomery form, R = 2^384), a G2 point = 24 x u64 (x.c0, x.c1, y.c0, y.c1);
the point at infinity is all zeros.  This module is host-side glue: pure Python/numpy, nothing on the prover path.
"""
import struct
from dataclasses import dataclass, field
from typing import List

import numpy as np

from .generators import FQ_MOD
from .workload import FR_MOD

_R_FR = 1 << 256
_RINV_FR = pow(_R_FR, -1, FR_MOD)
_R_FQ = 1 << 384
_RINV_FQ = pow(_R_FQ, -1, FQ_MOD)
_M64 = (1 << 64) - 1
FLAG_Y_LARGER = 1 << 7
FLAG_INFINITY = 1 << 6


class SerializationError(ValueError):
    """reference: Error::SerializationError (src/error.rs:5-14)"""


# ---------------------------------------------------------------------------------------------- field elements
def _limbs_to_int(limbs):
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def _int_to_limbs(v, words):
    return [(v >> (64 * i)) & _M64 for i in range(words)]


def fr_to_canonical(limbs4):
    return _limbs_to_int(limbs4) * _RINV_FR % FR_MOD


def fr_from_canonical(v):
    return _int_to_limbs(v % FR_MOD * _R_FR % FR_MOD, 4)


def fq_to_canonical(limbs6):
    return _limbs_to_int(limbs6) * _RINV_FQ % FQ_MOD


def fq_from_canonical(v):
    return _int_to_limbs(v % FQ_MOD * _R_FQ % FQ_MOD, 6)


def fr_array_to_bytes(a):
    """(n, 4) u64 Montgomery -> n * 32 bytes (no length prefix)"""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    # circuits have few distinct coefficients: convert each distinct residue once
    uniq, inv = np.unique(a, axis=0, return_inverse=True)
    table = np.frombuffer(b"".join(fr_to_canonical(r).to_bytes(32, "little") for r in uniq), dtype=np.uint8).reshape(-1, 32)
    return table[inv.reshape(-1)].tobytes()


def fr_array_from_bytes(data, n):
    if len(data) < 32 * n:
        raise SerializationError("truncated field-element array")
    raw = np.frombuffer(data, dtype="<u8", count=4 * n).reshape(n, 4)
    uniq, inv = np.unique(raw, axis=0, return_inverse=True)
    out = np.empty((len(uniq), 4), dtype=np.uint64)
    for i, r in enumerate(uniq):
        v = _limbs_to_int(r)
        if v >= FR_MOD:
            raise SerializationError("field element is not reduced")
        out[i] = fr_from_canonical(v)
    return out[inv.reshape(-1)]


# ---------------------------------------------------------------------------------------------- curve points
def _fq2_larger(y):
    """is y = (c0, c1) larger than -y in arkworks' Fq2 order (c1 first, then c0)?"""
    ny = ((-y[0]) % FQ_MOD, (-y[1]) % FQ_MOD)
    return (y[1], y[0]) > (ny[1], ny[0])


def _fq_sqrt(a):
    r = pow(a, (FQ_MOD + 1) // 4, FQ_MOD)          # p = 3 mod 4
    return r if r * r % FQ_MOD == a % FQ_MOD else None


def _fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % FQ_MOD, (a[0] * b[1] + a[1] * b[0]) % FQ_MOD)


def _fq2_sqrt(a):
    """square root in Fq[u]/(u^2 + 1) by the norm method; None for a non-residue"""
    a0, a1 = a[0] % FQ_MOD, a[1] % FQ_MOD
    if a1 == 0:
        r = _fq_sqrt(a0)
        if r is not None:
            return (r, 0)
        r = _fq_sqrt((-a0) % FQ_MOD)
        return None if r is None else (0, r)
    n = _fq_sqrt((a0 * a0 + a1 * a1) % FQ_MOD)
    if n is None:
        return None
    inv2 = pow(2, -1, FQ_MOD)
    for s in (n, (-n) % FQ_MOD):
        x0 = _fq_sqrt((a0 + s) * inv2 % FQ_MOD)
        if x0 is None or x0 == 0:
            continue
        x1 = a1 * pow(2 * x0, -1, FQ_MOD) % FQ_MOD
        if _fq2_mul((x0, x1), (x0, x1)) == (a0, a1):
            return (x0, x1)
    return None


def g1_to_bytes(pt12, compressed=True):
    pt = [int(v) for v in np.asarray(pt12, dtype=np.uint64).reshape(12)]
    if not any(pt):
        out = bytearray(48 if compressed else 96)
        out[-1] |= FLAG_INFINITY
        return bytes(out)
    x, y = fq_to_canonical(pt[:6]), fq_to_canonical(pt[6:])
    if compressed:
        out = bytearray(x.to_bytes(48, "little"))
        if y > (FQ_MOD - y) % FQ_MOD:
            out[47] |= FLAG_Y_LARGER
        return bytes(out)
    return x.to_bytes(48, "little") + y.to_bytes(48, "little")


def g2_to_bytes(pt24, compressed=True):
    pt = [int(v) for v in np.asarray(pt24, dtype=np.uint64).reshape(24)]
    if not any(pt):
        out = bytearray(96 if compressed else 192)
        out[-1] |= FLAG_INFINITY
        return bytes(out)
    x = (fq_to_canonical(pt[0:6]), fq_to_canonical(pt[6:12]))
    y = (fq_to_canonical(pt[12:18]), fq_to_canonical(pt[18:24]))
    xb = x[0].to_bytes(48, "little") + x[1].to_bytes(48, "little")
    if compressed:
        out = bytearray(xb)
        if _fq2_larger(y):
            out[95] |= FLAG_Y_LARGER
        return bytes(out)
    return xb + y[0].to_bytes(48, "little") + y[1].to_bytes(48, "little")


def _split_flags(chunk):
    last = chunk[-1]
    body = bytes(chunk[:-1]) + bytes([last & 0x3F])
    return body, last & 0xC0


def g1_from_bytes(data, compressed=True, check=True):
    size = 48 if compressed else 96
    if len(data) < size:
        raise SerializationError("truncated G1 point")
    body, flags = _split_flags(data[:size])
    if flags & FLAG_INFINITY:
        return np.zeros(12, dtype=np.uint64)
    x = int.from_bytes(body[:48], "little")
    if x >= FQ_MOD:
        raise SerializationError("G1 x coordinate is not reduced")
    if compressed:
        y = _fq_sqrt((x * x * x + 4) % FQ_MOD)
        if y is None:
            raise SerializationError("G1 x coordinate is not on the curve")
        if (y > (FQ_MOD - y) % FQ_MOD) != bool(flags & FLAG_Y_LARGER):
            y = (FQ_MOD - y) % FQ_MOD
    else:
        y = int.from_bytes(body[48:96], "little")
        if check and (y >= FQ_MOD or (y * y - x * x * x - 4) % FQ_MOD):
            raise SerializationError("G1 point is not on the curve")
    return np.array(fq_from_canonical(x) + fq_from_canonical(y), dtype=np.uint64)


def g2_from_bytes(data, compressed=True, check=True):
    size = 96 if compressed else 192
    if len(data) < size:
        raise SerializationError("truncated G2 point")
    body, flags = _split_flags(data[:size])
    if flags & FLAG_INFINITY:
        return np.zeros(24, dtype=np.uint64)
    x = (int.from_bytes(body[0:48], "little"), int.from_bytes(body[48:96], "little"))
    if x[0] >= FQ_MOD or x[1] >= FQ_MOD:
        raise SerializationError("G2 x coordinate is not reduced")
    rhs = _fq2_mul(_fq2_mul(x, x), x)
    rhs = ((rhs[0] + 4) % FQ_MOD, (rhs[1] + 4) % FQ_MOD)          # y^2 = x^3 + 4(1 + u)
    if compressed:
        y = _fq2_sqrt(rhs)
        if y is None:
            raise SerializationError("G2 x coordinate is not on the curve")
        if _fq2_larger(y) != bool(flags & FLAG_Y_LARGER):
            y = ((-y[0]) % FQ_MOD, (-y[1]) % FQ_MOD)
    else:
        y = (int.from_bytes(body[96:144], "little"), int.from_bytes(body[144:192], "little"))
        if check and (y[0] >= FQ_MOD or y[1] >= FQ_MOD or _fq2_mul(y, y) != rhs):
            raise SerializationError("G2 point is not on the curve")
    return np.array(fq_from_canonical(x[0]) + fq_from_canonical(x[1]) + fq_from_canonical(y[0]) + fq_from_canonical(y[1]),
                    dtype=np.uint64)


# ---------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, data):
        self.d = memoryview(bytes(data))
        self.o = 0

    def take(self, n):
        if self.o + n > len(self.d):
            raise SerializationError("unexpected end of data")
        out = self.d[self.o:self.o + n]
        self.o += n
        return out

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def fr(self):
        return fr_array_from_bytes(self.take(32), 1)[0]

    def fr_vec(self):
        n = self.u64()
        return fr_array_from_bytes(self.take(32 * n), n)

    def g1(self, compressed=True, check=True):
        return g1_from_bytes(self.take(48 if compressed else 96), compressed, check)

    def g2(self, compressed=True, check=True):
        return g2_from_bytes(self.take(96 if compressed else 192), compressed, check)

    def done(self):
        if self.o != len(self.d):
            raise SerializationError("%d trailing bytes" % (len(self.d) - self.o))


def _u64(v):
    return struct.pack("<Q", int(v))


# ---------------------------------------------------------------------------------------------- Proof
@dataclass
class OpenProof:
    """commitment::open::Proof { h, proofs } (src/commitment/open.rs:12-15)"""
    h: np.ndarray
    proofs: np.ndarray                       # (nv, 24)

    def to_bytes(self):
        return g2_to_bytes(self.h) + _u64(len(self.proofs)) + b"".join(g2_to_bytes(q) for q in self.proofs)

    @staticmethod
    def read(r):
        h = r.g2()
        n = r.u64()
        proofs = np.stack([r.g2() for _ in range(n)]) if n else np.zeros((0, 24), dtype=np.uint64)
        return OpenProof(h, proofs)


@dataclass
class Proof:
    """data_structures::proof::Proof (src/data_structures/proof.rs:11-20); Fr values are Montgomery limbs."""
    commitment_nv: int                       # ProverFirstMessage.commitment.nv          (commit.rs:11-14)
    commitment: np.ndarray                   # ProverFirstMessage.commitment.g_product   (12,)
    z_rv_0: np.ndarray                       # ProverSecondMessage                        (prover.rs:73-77)
    proof_for_z_rv_0: OpenProof
    third_index_info: tuple                  # ProverThirdMessage.ml_index_info: the two usize fields, as serialized
    first_sumcheck_messages: List[np.ndarray] = field(default_factory=list)   # each (evals, 4)
    va: np.ndarray = None                    # ProverFourthMessage                        (prover.rs:86-91)
    vb: np.ndarray = None
    vc: np.ndarray = None
    fifth_index_info: tuple = (0, 0)
    second_sumcheck_messages: List[np.ndarray] = field(default_factory=list)
    z_ry: np.ndarray = None                  # ProverSixthMessage                         (prover.rs:99-103)
    proof_for_z_ry: OpenProof = None

    @staticmethod
    def from_bytes(data):
        r = _Reader(data)
        nv = r.u64(); com = r.g1()
        z_rv_0 = r.fr(); open1 = OpenProof.read(r)
        info3 = (r.u64(), r.u64())
        sc1 = [r.fr_vec() for _ in range(r.u64())]
        va, vb, vc = r.fr(), r.fr(), r.fr()
        info5 = (r.u64(), r.u64())
        sc2 = [r.fr_vec() for _ in range(r.u64())]
        z_ry = r.fr(); open2 = OpenProof.read(r)
        r.done()
        return Proof(nv, com, z_rv_0, open1, info3, sc1, va, vb, vc, info5, sc2, z_ry, open2)

    def to_bytes(self):
        def msgs(ms):
            return _u64(len(ms)) + b"".join(_u64(len(m)) + fr_array_to_bytes(m) for m in ms)
        return b"".join([
            _u64(self.commitment_nv), g1_to_bytes(self.commitment),
            fr_array_to_bytes(self.z_rv_0), self.proof_for_z_rv_0.to_bytes(),
            _u64(self.third_index_info[0]), _u64(self.third_index_info[1]),
            msgs(self.first_sumcheck_messages),
            fr_array_to_bytes(self.va), fr_array_to_bytes(self.vb), fr_array_to_bytes(self.vc),
            _u64(self.fifth_index_info[0]), _u64(self.fifth_index_info[1]),
            msgs(self.second_sumcheck_messages),
            fr_array_to_bytes(self.z_ry), self.proof_for_z_ry.to_bytes()])


# ---------------------------------------------------------------------------------------------- IndexPK / IndexVK
def matrix_to_bytes(csr, n):
    """MatrixExtension { constraint: Vec<Vec<(F, usize)>>, num_constraints } from CSR arrays
    (row_ptr u64[n+1], col u32[nnz], val (nnz, 4) u64 Montgomery)."""
    row_ptr, col, val = csr
    row_ptr = np.asarray(row_ptr, dtype=np.uint64)
    col = np.asarray(col, dtype=np.uint64)
    nnz = int(row_ptr[-1])
    rows = len(row_ptr) - 1
    vb = np.frombuffer(fr_array_to_bytes(np.asarray(val)[:nnz]), dtype=np.uint8).reshape(nnz, 32) if nnz else np.zeros((0, 32), np.uint8)
    out = np.zeros(8 + 8 * rows + 40 * nnz + 8, dtype=np.uint8)
    # entry e of row i sits at 8 (outer length) + 8 (i + 1) (row lengths so far) + 40 e
    ent_row = np.repeat(np.arange(rows, dtype=np.int64), np.diff(row_ptr).astype(np.int64))
    ent_off = 8 + 8 * (ent_row + 1) + 40 * np.arange(nnz, dtype=np.int64)
    idx = ent_off[:, None] + np.arange(32)[None, :]
    out[idx] = vb
    colb = col[:nnz].astype("<u8").view(np.uint8).reshape(nnz, 8)
    out[ent_off[:, None] + 32 + np.arange(8)[None, :]] = colb
    row_off = 8 + 8 * np.arange(rows, dtype=np.int64) + 40 * row_ptr[:-1].astype(np.int64)
    lens = np.diff(row_ptr).astype("<u8").view(np.uint8).reshape(rows, 8)
    out[row_off[:, None] + np.arange(8)[None, :]] = lens
    out[0:8] = np.frombuffer(_u64(rows), dtype=np.uint8)
    out[-8:] = np.frombuffer(_u64(n), dtype=np.uint8)
    return out.tobytes()


def _matrix_read(r):
    rows = r.u64()
    row_ptr = np.zeros(rows + 1, dtype=np.uint64)
    cols, vals = [], []
    for i in range(rows):
        k = r.u64()
        body = np.frombuffer(r.take(40 * k), dtype=np.uint8).reshape(k, 40)
        vals.append(body[:, :32].copy())
        cols.append(body[:, 32:].copy().view("<u8").reshape(k))
        row_ptr[i + 1] = row_ptr[i] + np.uint64(k)
    n = r.u64()
    nnz = int(row_ptr[-1])
    col = np.concatenate(cols).astype(np.uint32) if nnz else np.zeros(0, np.uint32)
    val = fr_array_from_bytes(np.concatenate(vals).tobytes(), nnz) if nnz else np.zeros((0, 4), np.uint64)
    return (row_ptr, col, val), n


def index_to_bytes(mats, log_n):
    """IndexPK and IndexVK share one layout (src/ahp/indexer.rs:10-27): matrix_a, matrix_b, matrix_c, log_n."""
    n = 1 << log_n
    return b"".join(matrix_to_bytes(m, n) for m in mats) + _u64(log_n)


def index_from_bytes(data):
    r = _Reader(data)
    mats = []
    for _ in range(3):
        m, n = _matrix_read(r)
        mats.append(m)
    log_n = r.u64()
    r.done()
    if any(len(m[0]) - 1 != (1 << log_n) for m in mats) or n != (1 << log_n):
        raise SerializationError("matrix sizes do not match log_n")
    return mats, log_n


# ---------------------------------------------------------------------------------------------- commitment keys
def _points_to_bytes(arr, group, compressed):
    f = g1_to_bytes if group == 1 else g2_to_bytes
    return _u64(len(arr)) + b"".join(f(p, compressed) for p in arr)


def _points_read(r, group, compressed, check):
    n = r.u64()
    words = 12 if group == 1 else 24
    out = np.zeros((n, words), dtype=np.uint64)
    for i in range(n):
        out[i] = r.g1(compressed, check) if group == 1 else r.g2(compressed, check)
    return out


def public_parameter_to_bytes(nv, powers_of_g, powers_of_h, g, h, compressed=True):
    """PublicParameter { nv, powers_of_g, powers_of_h, g, h } (src/commitment/data_structures.rs:9-17);
    powers_of_x[i] has 2^(nv - i) points (setup.rs:75-84)."""
    out = [_u64(nv), _u64(len(powers_of_g))]
    out += [_points_to_bytes(lvl, 1, compressed) for lvl in powers_of_g]
    out.append(_u64(len(powers_of_h)))
    out += [_points_to_bytes(lvl, 2, compressed) for lvl in powers_of_h]
    out += [g1_to_bytes(g, compressed), g2_to_bytes(h, compressed)]
    return b"".join(out)


def _public_parameter_read(r, compressed, check):
    nv = r.u64()
    pg = [_points_read(r, 1, compressed, check) for _ in range(r.u64())]
    ph = [_points_read(r, 2, compressed, check) for _ in range(r.u64())]
    g = r.g1(compressed, check); h = r.g2(compressed, check)
    if len(pg) != nv or len(ph) != nv or any(len(x) != 1 << (nv - i) for i, x in enumerate(pg)) \
            or any(len(x) != 1 << (nv - i) for i, x in enumerate(ph)):
        raise SerializationError("public parameter levels do not match nv")
    return dict(nv=nv, powers_of_g=pg, powers_of_h=ph, g=g, h=h)


def public_parameter_from_bytes(data, compressed=True, check=True):
    r = _Reader(data)
    out = _public_parameter_read(r, compressed, check)
    r.done()
    return out


def verifier_parameter_to_bytes(nv, g, h, g_mask_random, compressed=True):
    """VerifierParameter { nv, g, h, g_mask_random } (src/commitment/data_structures.rs:19-25)"""
    return _u64(nv) + g1_to_bytes(g, compressed) + g2_to_bytes(h, compressed) + _points_to_bytes(g_mask_random, 1, compressed)


def _verifier_parameter_read(r, compressed, check):
    nv = r.u64()
    g = r.g1(compressed, check); h = r.g2(compressed, check)
    mask = _points_read(r, 1, compressed, check)
    if len(mask) != nv:
        raise SerializationError("g_mask_random does not match nv")
    return dict(nv=nv, g=g, h=h, g_mask_random=mask)


def verifier_parameter_from_bytes(data, compressed=True, check=True):
    r = _Reader(data)
    out = _verifier_parameter_read(r, compressed, check)
    r.done()
    return out


def key_cache_to_bytes(pairs):
    """`benchmark_cached_keys` (src/commitment/mod.rs:48-56): Vec<ParameterPair { pp, vp }>, uncompressed."""
    out = [_u64(len(pairs))]
    for pp, vp in pairs:
        out.append(public_parameter_to_bytes(pp["nv"], pp["powers_of_g"], pp["powers_of_h"], pp["g"], pp["h"], compressed=False))
        out.append(verifier_parameter_to_bytes(vp["nv"], vp["g"], vp["h"], vp["g_mask_random"], compressed=False))
    return b"".join(out)


def key_cache_from_bytes(data, check=False):
    """mod.rs:58-63 reads the cache with `deserialize_unchecked`, hence check=False by default."""
    r = _Reader(data)
    pairs = []
    for _ in range(r.u64()):
        pp = _public_parameter_read(r, False, check)
        vp = _verifier_parameter_read(r, False, check)
        pairs.append((pp, vp))
    r.done()
    return pairs
