"""Synthetic R1CS workload of the reference's own benchmark, generated host-side for bench.py and tests.

Restates /root/reference/src/data_structures/constraints.rs:39-110 (`TestSynthesizer`) as driven by
/root/reference/src/test_utils.rs:51-102 (`generate_circuit_with_random_input(.., pad_to_square=true, ..)`)
and /root/reference/src/benchmark.rs:63-65 (32 public inputs, density 0):

  * instance variables: column 0 = constant one, columns 1.. = public inputs; witness j -> column
    num_public + j (arkworks `to_matrices` column convention);
  * constraint i < num_sparse, i even:  (a + b + off) * 1 = c      A: 3 non-zeros, B: column 0, C: 1
                               i odd:   a * (b + off) = c          A: 1, B: 2, C: 1
    with `off` a uniformly chosen public input and the sliding window a <- b, b <- c;
  * one (density 0) final dense constraint (sum of every assigned variable)^2 = c, where the first
    public input carries coefficient 2 (constraints.rs:43,47 push (a_val, a_var) twice);
  * empty padding rows up to n = num_public + num_private (make_matrices_square).

Randomness: the reference draws from ark_ff::test_rng(), which is not reproducible outside arkworks;
this generator uses SplitMix64(seed) and arkworks' Fr sampling rule (4 x u64 limbs, top bit cleared,
rejection, limbs taken as the Montgomery residue).  The CPU oracle restates the same generator
independently (oracle/spartan_oracle.cpp synth_r1cs) and tests/ checks they agree bit for bit.
"""
import numpy as np

FR_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
_R = 1 << 256
_RINV = pow(_R, -1, FR_MOD)
_ONE_M = _R % FR_MOD
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _M64

    def next_u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)


def fr_rand_mont(rng):
    """Montgomery residue of a fresh field element (arkworks `Fr::rand`)."""
    while True:
        v = 0
        for i in range(4):
            limb = rng.next_u64()
            if i == 3:
                limb &= (1 << 63) - 1
            v |= limb << (64 * i)
        if v < FR_MOD:
            return v


def mont_to_limbs(vals):
    """list of Montgomery residues (python ints) -> (len, 4) uint64"""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for j in range(4):
        out[:, j] = np.fromiter(((v >> (64 * j)) & _M64 for v in vals), dtype=np.uint64, count=len(vals))
    return out


def limbs_to_mont(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def fr_from_int(v):
    """canonical integer -> (4,) uint64 Montgomery limbs"""
    return mont_to_limbs([(v % FR_MOD) * _R % FR_MOD])[0]


def fr_to_int(limbs):
    return limbs_to_mont(np.asarray(limbs).reshape(1, 4))[0] * _RINV % FR_MOD


class SyntheticR1CS:
    """CSR matrices a, b, c (each: row_ptr u64[n+1], col u32[nnz], val (nnz,4) u64 Montgomery) + v, w."""

    def __init__(self, num_public, num_private, density=0, seed=0):
        if num_public <= 3:
            raise ValueError("number of public variables should be greater to 3")
        rng = SplitMix64(seed)
        p = FR_MOD
        inst = [_ONE_M]                   # Instance(0) = One
        wit = []
        a_val = fr_rand_mont(rng); inst.append(a_val); a_var = 1
        b_val = fr_rand_mont(rng); inst.append(b_val); b_var = 2
        assign_val = [a_val, a_val]       # sic: (a_val, a_var) twice
        assign_var = [1, 1]
        for _ in range(num_public - 3):
            val = fr_rand_mont(rng); inst.append(val)
            assign_val.append(val); assign_var.append(len(inst) - 1)
        n_inst = len(inst)
        assert n_inst == num_public
        num_sparse = (num_private - 1) * (510 - density) // 510
        rows = [[], [], []]               # per matrix: list of rows, each [(col, coeff_mont)]
        one = _ONE_M
        for i in range(num_sparse):
            oi = 2 + rng.next_u64() % (num_public - 3)          # gen_range(2, num_public - 1)
            off_val, off_var = assign_val[oi], assign_var[oi]
            c_var = n_inst + len(wit)
            if i % 2 != 0:
                c_val = a_val * ((b_val + off_val) % p) % p * _RINV % p
                rows[0].append(_lc([a_var]))
                rows[1].append(_lc([b_var, off_var]))
            else:
                c_val = (a_val + b_val + off_val) % p
                rows[0].append(_lc([a_var, b_var, off_var]))
                rows[1].append(_lc([0]))
            rows[2].append(_lc([c_var]))
            wit.append(c_val)
            assign_val.append(c_val); assign_var.append(c_var)
            a_val, a_var, b_val, b_var = b_val, b_var, c_val, c_var
        for _ in range(num_sparse, num_private):
            lc = _lc(assign_var)
            s = sum(assign_val) % p
            c_val = s * s % p * _RINV % p
            c_var = n_inst + len(wit)
            rows[0].append(lc); rows[1].append(list(lc)); rows[2].append(_lc([c_var]))
            wit.append(c_val)
        num_formatted = num_public + num_private
        num_constraints = len(rows[0])
        if num_formatted > num_constraints:
            for m in rows:
                m.extend([[] for _ in range(num_formatted - num_constraints)])
        else:
            wit.extend([one] * (num_constraints - num_formatted))
        self.n = len(rows[0])
        self.log_n = self.n.bit_length() - 1
        self.num_public = num_public
        self.mats = [_to_csr(m) for m in rows]
        self.v = mont_to_limbs(inst)
        self.w = mont_to_limbs(wit)

    @property
    def nnz(self):
        return [int(m[0][-1]) for m in self.mats]


def _lc(variables):
    """`lc!() + var + ...`: sorted by variable, duplicates merged by adding coefficients (Montgomery)."""
    acc = {}
    for v in variables:
        acc[v] = (acc.get(v, 0) + _ONE_M) % FR_MOD
    return sorted((c, k) for c, k in acc.items() if k != 0)


def _to_csr(rows):
    n = len(rows)
    row_ptr = np.zeros(n + 1, dtype=np.uint64)
    lens = np.fromiter((len(r) for r in rows), dtype=np.uint64, count=n)
    np.cumsum(lens, out=row_ptr[1:])
    nnz = int(row_ptr[-1])
    col = np.fromiter((c for r in rows for (c, _) in r), dtype=np.uint32, count=nnz)
    val = mont_to_limbs([k for r in rows for (_, k) in r]) if nnz else np.zeros((0, 4), dtype=np.uint64)
    return row_ptr, col, val
