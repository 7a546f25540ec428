"""Python big-integer model of the r1cs-spartan prover path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (tsunrise/r1cs-spartan, Rust + un-pinned arkworks git
dependencies) cannot be compiled in this environment and ships no golden vectors, so this
model and the C restatement in this directory are pinned against EACH OTHER and against the
algebraic identities the reference's own tests check -- not against arkworks output.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline/reference leg may import
this module.  The product path (r1cs-spartan_b200/) never does.

Everything here follows the reference literally (no algebraic shortcuts):
  * eq_extension            -> /root/reference/src/data_structures/eq.rs:5-20
  * sum_over_y / eval_on_x  -> /root/reference/src/data_structures/r1cs_reader.rs:75-117
  * commit / open           -> /root/reference/src/commitment/commit.rs:17-29, open.rs:19-58
  * keygen                  -> /root/reference/src/commitment/setup.rs:27-105
  * prover rounds           -> /root/reference/src/ahp/prover.rs:109-281
  * Fiat-Shamir driver      -> /root/reference/src/lib.rs:58-146
Upstream (arkworks, late 2020) behaviour is restated from its published algorithm; each such
function is tagged UPSTREAM and is a single swappable definition.
"""
import hashlib

X_BLS = -0xd201000000010000
P = (X_BLS - 1) ** 2 * (X_BLS ** 4 - X_BLS ** 2 + 1) // 3 + X_BLS   # base field modulus (381 bit)
R = X_BLS ** 4 - X_BLS ** 2 + 1                                      # scalar field modulus (255 bit)
R_MONT = (1 << 256) % R        # Montgomery radix for Fr (4 x u64 limbs)
R_MONT_INV = pow(R_MONT, -1, R)
H1 = (X_BLS - 1) ** 2 // 3     # G1 cofactor

# ---------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)
F2_ZERO = (0, 0)
F2_ONE = (1, 0)


class Fld:
    """Tiny field vtable so one Jacobian implementation serves G1 (Fq) and G2 (Fq2)."""
    def __init__(self, add, sub, mul, neg, inv, zero, one):
        self.add, self.sub, self.mul, self.neg, self.inv = add, sub, mul, neg, inv
        self.zero, self.one = zero, one

FQ = Fld(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P,
         lambda a: (-a) % P, lambda a: pow(a, -1, P), 0, 1)
FQ2 = Fld(f2_add, f2_sub, f2_mul, f2_neg, f2_inv, F2_ZERO, F2_ONE)

# ---------------------------------------------------------------- short Weierstrass, a = 0
# points are None (infinity) or affine (x, y); Jacobian (X, Y, Z) internally.
def jac_double(F, pt):
    X, Y, Z = pt
    if Z == F.zero:
        return pt
    A = F.mul(X, X); B = F.mul(Y, Y); C = F.mul(B, B)
    t = F.add(X, B); D = F.sub(F.sub(F.mul(t, t), A), C); D = F.add(D, D)
    E = F.add(F.add(A, A), A); Fv = F.mul(E, E)
    X3 = F.sub(Fv, F.add(D, D))
    C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    Z3 = F.mul(F.add(Y, Y), Z)
    return (X3, Y3, Z3)

def jac_add(F, p1, p2):
    if p1[2] == F.zero: return p2
    if p2[2] == F.zero: return p1
    X1, Y1, Z1 = p1; X2, Y2, Z2 = p2
    Z1Z1 = F.mul(Z1, Z1); Z2Z2 = F.mul(Z2, Z2)
    U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2); S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 == S2:
            return jac_double(F, p1)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1); Rr = F.sub(S2, S1)
    HH = F.mul(H, H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.mul(Rr, Rr), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)

def to_jac(F, a):
    return (F.one, F.one, F.zero) if a is None else (a[0], a[1], F.one)

def to_affine(F, j):
    if j[2] == F.zero: return None
    zi = F.inv(j[2]); zi2 = F.mul(zi, zi)
    return (F.mul(j[0], zi2), F.mul(j[1], F.mul(zi2, zi)))

def pt_mul(F, a, k):
    """k * a for affine a (None = infinity); k any non-negative integer."""
    acc = (F.one, F.one, F.zero)
    if a is None or k == 0: return None
    base = to_jac(F, a)
    for bit in bin(k)[2:]:
        acc = jac_double(F, acc)
        if bit == '1':
            acc = jac_add(F, acc, base)
    return to_affine(F, acc)

def pt_add(F, a, b):
    return to_affine(F, jac_add(F, to_jac(F, a), to_jac(F, b)))

def pt_neg(F, a):
    return None if a is None else (a[0], F.neg(a[1]))

def msm(F, bases, scalars):
    """UPSTREAM ark_ec::msm::VariableBaseMSM::multi_scalar_mul: zips (truncates to shorter);
    the result is method independent, so the model uses plain double-and-add."""
    acc = (F.one, F.one, F.zero)
    for b, s in zip(bases, scalars):
        if s % R and b is not None:
            acc = jac_add(F, acc, to_jac(F, pt_mul(F, b, s % R)))
    return to_affine(F, acc)

def on_curve_g1(a): return a is None or (a[1] * a[1] - a[0] ** 3 - 4) % P == 0
def on_curve_g2(a):
    if a is None: return True
    x3 = f2_mul(f2_mul(a[0], a[0]), a[0])
    return f2_sub(f2_mul(a[1], a[1]), f2_add(x3, (4, 4))) == F2_ZERO

def fq_sqrt(a):
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None

def f2_sqrt(a):
    """sqrt in Fq2 (p = 3 mod 4), complex method."""
    if a == F2_ZERO: return F2_ZERO
    n = (a[0] * a[0] + a[1] * a[1]) % P
    s = fq_sqrt(n)
    if s is None: return None
    inv2 = pow(2, -1, P)
    for sg in (s, (-s) % P):
        d = (a[0] + sg) * inv2 % P
        x0 = fq_sqrt(d)
        if x0 is None or x0 == 0: continue
        x1 = a[1] * pow(2 * x0, -1, P) % P
        c = (x0, x1)
        if f2_mul(c, c) == (a[0] % P, a[1] % P): return c
    return None

def derive_generators():
    """Deterministic generators of the r-torsion subgroups (the reference draws g, h with
    G::rand(test_rng()), setup.rs:28-29 -- any subgroup generator is an equally valid pp)."""
    x = 1
    while True:
        y = fq_sqrt((x ** 3 + 4) % P)
        if y is not None:
            g = pt_mul(FQ, (x, min(y, P - y)), H1)
            if g is not None: break
        x += 1
    assert pt_mul(FQ, g, R) is None and on_curve_g1(g)
    # order of the sextic twist E'(Fq2): y^2 = x^3 + 4(1+u)
    t = X_BLS + 1
    t2 = t * t - 2 * P
    f2sq = (4 * P * P - t2 * t2) // 3
    f = int(f2sq ** 0.5) if f2sq < 1 << 52 else _isqrt(f2sq)
    assert f * f == f2sq
    cands = [P * P + 1 - (t2 + 3 * f) // 2, P * P + 1 - (t2 - 3 * f) // 2]
    n2 = [c for c in cands if c % R == 0][0]
    h2 = n2 // R
    xx = 1
    while True:
        X = (xx, 1)
        rhs = f2_add(f2_mul(f2_mul(X, X), X), (4, 4))
        Y = f2_sqrt(rhs)
        if Y is not None:
            h = pt_mul(FQ2, (X, Y), h2)
            if h is not None and pt_mul(FQ2, h, R) is None: break
        xx += 1
    assert on_curve_g2(h)
    return g, h

def _isqrt(n):
    import math
    return math.isqrt(n)

# ---------------------------------------------------------------- deterministic inputs
class SplitMix64:
    """Workload PRNG (ours; the reference uses ark_ff::test_rng(), not reproducible here)."""
    M = (1 << 64) - 1
    def __init__(self, seed): self.s = seed & self.M
    def next_u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & self.M
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M
        return z ^ (z >> 31)

def fr_rand(rng):
    """UPSTREAM ark_ff Fp256 `Standard` sampling: four next_u64 -> limbs[0..4], clear the top
    REPR_SHAVE_BITS=1 bit, accept iff < r; the accepted limbs ARE the Montgomery residue."""
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= (1 << 63) - 1
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < R:
            return v * R_MONT_INV % R

# ---------------------------------------------------------------- transcript
class Blake2sRng:
    """UPSTREAM linear_sumcheck::data_structures::random::Blake2s512Rng (FeedableRNG)."""
    def __init__(self): self.h = hashlib.blake2s()
    def feed(self, data: bytes): self.h.update(data)
    def fill_bytes(self, n):
        out = self.h.copy().digest()
        res = bytearray(); dp = 0
        while len(res) < n:
            res.append(out[dp]); dp += 1
            if dp == 32:
                self.h.update(out); out = self.h.copy().digest(); dp = 0
        self.h.update(out)
        return bytes(res)
    def next_u64(self): return int.from_bytes(self.fill_bytes(8), 'little')

# ---------------------------------------------------------------- CanonicalSerialize (UPSTREAM)
def ser_u64(v): return int(v).to_bytes(8, 'little')
def ser_fr(v): return int(v % R).to_bytes(32, 'little')
def ser_fr_vec(vs): return ser_u64(len(vs)) + b''.join(ser_fr(v) for v in vs)
def ser_g1(a):
    if a is None:
        b = bytearray(48); b[47] |= 1 << 6; return bytes(b)
    b = bytearray(a[0].to_bytes(48, 'little'))
    if a[1] > (P - a[1]) % P: b[47] |= 1 << 7
    return bytes(b)
def ser_g2(a):
    if a is None:
        b = bytearray(96); b[95] |= 1 << 6; return bytes(b)
    b = bytearray(a[0][0].to_bytes(48, 'little') + a[0][1].to_bytes(48, 'little'))
    y = a[1]; ny = f2_neg(y)
    if (y[1], y[0]) > (ny[1], ny[0]): b[95] |= 1 << 7    # Fq2 order: c1 then c0
    return bytes(b)
def ser_matrix(rows, n):
    """MatrixExtension{constraint: Vec<Vec<(F,usize)>>, num_constraints} r1cs_reader.rs:9-13"""
    out = [ser_u64(len(rows))]
    for row in rows:
        out.append(ser_u64(len(row)))
        for (val, col) in row:
            out.append(ser_fr(val)); out.append(ser_u64(col))
    out.append(ser_u64(n))
    return b''.join(out)
def ser_index_info(max_mult, nv): return ser_u64(max_mult) + ser_u64(nv)
def ser_commitment(nv, pt): return ser_u64(nv) + ser_g1(pt)
def ser_open_proof(h, proofs): return ser_g2(h) + ser_u64(len(proofs)) + b''.join(ser_g2(q) for q in proofs)

# ---------------------------------------------------------------- MLE helpers (UPSTREAM MLExtensionArray)
def mle_fold(tab, r):
    return [(tab[2 * b] * (1 - r) + tab[2 * b + 1] * r) % R for b in range(len(tab) // 2)]
def mle_eval(tab, point):
    for r in point: tab = mle_fold(tab, r)
    return tab[0]

def eq_extension(t):
    """eq.rs:5-20 -- dim tables of size 2^dim."""
    dim = len(t); res = []
    for i in range(dim):
        poly = []
        for x in range(1 << dim):
            xi = (x >> i) & 1
            ti_xi = t[i] * xi
            poly.append((ti_xi + ti_xi - xi - t[i] + 1) % R)
        res.append(poly)
    return res

def sum_over_y(rows, z):
    """r1cs_reader.rs:75-85"""
    return [sum(a * z[y] for (a, y) in row) % R for row in rows]

def eval_on_x(rows, r_x):
    """r1cs_reader.rs:91-117: M(r_x, y) via partial evaluation of the low (x) variables."""
    n = len(rows)
    eq = [1]
    for r in r_x:            # eq(r_x, x), variable i <-> bit i of x
        eq = [e * (1 - r) % R for e in eq] + [e * r % R for e in eq]
    out = [0] * n
    for x, row in enumerate(rows):
        for (val, y) in row:
            out[y] = (out[y] + val * eq[x]) % R
    return out

def sumcheck_round(products, r):
    """UPSTREAM AHPForMLSumcheck::prove_round: optional fold of every table with r, then
    evaluations[t] = sum_b sum_products prod_tables (T[2b](1-t)+T[2b+1]t), t=0..max_mult."""
    if r is not None:
        products = [[mle_fold(tab, r) for tab in prod] for prod in products]
    deg = max(len(p) for p in products)
    half = len(products[0][0]) // 2
    evals = [0] * (deg + 1)
    for b in range(half):
        for t in range(deg + 1):
            for prod in products:
                acc = 1
                for tab in prod:
                    acc = acc * (tab[2 * b] * (1 - t) + tab[2 * b + 1] * t) % R
                evals[t] = (evals[t] + acc) % R
    return products, evals

# ---------------------------------------------------------------- commitment (Libra / PST)
def keygen(nv, g, h, t):
    """setup.rs:27-105 with caller-supplied (g, h, t): powers_of_x[i][b] = x * eq(t[i..], b)."""
    pg, ph = [], []
    for i in range(nv):
        eq = [1]
        for tj in t[i:]:
            eq = [e * (1 - tj) % R for e in eq] + [e * tj % R for e in eq]
        pg.append([pt_mul(FQ, g, e) for e in eq])
        ph.append([pt_mul(FQ2, h, e) for e in eq])
    vp_mask = [pt_mul(FQ, g, tj) for tj in t]
    return dict(nv=nv, g=g, h=h, powers_of_g=pg, powers_of_h=ph), dict(nv=nv, g=g, h=h, g_mask=vp_mask)

def commit(pp, z):
    """commit.rs:17-29"""
    return msm(FQ, pp['powers_of_g'][0], z)

def pc_open(pp, z, point):
    """open.rs:19-58 (returns eval, proofs, q)"""
    nv = len(point)
    ev = mle_eval(list(z), point)
    r = list(z); proofs = []; qs = {}
    for i in range(nv):
        k = nv - i
        q = [(r[2 * b + 1] - r[2 * b]) % R for b in range(1 << (k - 1))]
        r = [(r[2 * b] * (1 - point[i]) + r[2 * b + 1] * point[i]) % R for b in range(1 << (k - 1))]
        scalars = [q[x >> 1] for x in range(1 << k)]
        proofs.append(msm(FQ2, pp['powers_of_h'][i], scalars))
        qs[k] = q
    return ev, proofs, qs

# ---------------------------------------------------------------- full NI prover (lib.rs:58-146)
def prove(rows_a, rows_b, rows_c, v, w, pp, trace=None):
    n = len(rows_a); log_n = n.bit_length() - 1
    assert 1 << log_n == n and len(v) + len(w) == n and len(v) & (len(v) - 1) == 0
    log_v = len(v).bit_length() - 1
    fs = Blake2sRng()
    for m in (rows_a, rows_b, rows_c): fs.feed(ser_matrix(m, n))
    fs.feed(ser_fr_vec(v))
    z = list(v) + list(w)
    # round 1: commitment
    com = commit(pp, z)
    pm1 = ser_commitment(log_n, com); fs.feed(pm1)
    r_v = [fr_rand(fs) for _ in range(log_v)]
    # round 2: open at (r_v, 0..0)
    ev, proofs, _ = pc_open(pp, z, r_v + [0] * (log_n - log_v))
    pm2 = ser_fr(ev) + ser_open_proof(pp['h'], proofs); fs.feed(pm2)
    tor = [fr_rand(fs) for _ in range(log_n)]
    # round 3
    eq = eq_extension(tor)
    az, bz, cz = sum_over_y(rows_a, z), sum_over_y(rows_b, z), sum_over_y(rows_c, z)
    products = [[az, bz] + eq, [[(-c) % R for c in cz]] + eq]
    pm3 = ser_index_info(log_n + 2, log_n); fs.feed(pm3)
    sc1 = []; r = None; r_x = []
    for _ in range(log_n):
        products, evals = sumcheck_round(products, r)
        msg = ser_fr_vec(evals); fs.feed(msg); sc1.append(msg)
        r = fr_rand(fs); r_x.append(r)
    va, vb, vc = mle_eval(az, r_x), mle_eval(bz, r_x), mle_eval(cz, r_x)
    pm4 = ser_fr(va) + ser_fr(vb) + ser_fr(vc); fs.feed(pm4)
    r_a, r_b, r_c = fr_rand(fs), fr_rand(fs), fr_rand(fs)
    # round 5
    ma = [x * r_a % R for x in eval_on_x(rows_a, r_x)]
    mb = [x * r_b % R for x in eval_on_x(rows_b, r_x)]
    mc = [x * r_c % R for x in eval_on_x(rows_c, r_x)]
    products = [[ma, list(z)], [mb, list(z)], [mc, list(z)]]
    pm5 = ser_index_info(2, log_n); fs.feed(pm5)
    sc2 = []; r = None; r_y = []
    for _ in range(log_n):
        products, evals = sumcheck_round(products, r)
        msg = ser_fr_vec(evals); fs.feed(msg); sc2.append(msg)
        r = fr_rand(fs); r_y.append(r)
    ev2, proofs2, _ = pc_open(pp, z, r_y)
    pm6 = ser_fr(ev2) + ser_open_proof(pp['h'], proofs2)
    if trace is not None:
        trace.update(az=az, bz=bz, cz=cz, r_v=r_v, tor=tor, r_x=r_x, r_y=r_y, va=va, vb=vb, vc=vc,
                     r_abc=(r_a, r_b, r_c), com=com, z_rv_0=ev, z_ry=ev2)
    return (pm1 + pm2 + pm3 + ser_u64(len(sc1)) + b''.join(sc1) + pm4 + pm5 +
            ser_u64(len(sc2)) + b''.join(sc2) + pm6)
