#!/usr/bin/env python3
"""Generator for csrc/fp_gen.cuh: fully unrolled Montgomery multiplication / add / sub in PTX carry
chains (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32 with carry predicates)
for the two BLS12-381 prime fields in 32-bit limbs (Fr: 8 limbs, Fq: 12 limbs).

The instruction stream is produced ONCE as an abstract list and rendered twice: as PTX text for the
header, and through a bit-exact 32-bit-register + carry-flag emulator (`emulate`) that
tests/test_cpu_oracle.py (test_generated_field_streams_under_emulation, test_lazy_reduction_sequences_under_emulation) runs against Python big integers -- so the sequence nvcc compiles is the
sequence that was verified on the CPU.

Scheme (per row i of b):   T = E + 2^32 * O   (E: even-offset accumulator, O: odd-offset accumulator)
    E += sum_{j even} a[j] b_i 2^(32 j)      one carry chain over N limbs (64-bit products abut)
    O += sum_{j odd}  a[j] b_i 2^(32 (j-1))  one carry chain
    m  = E[0] * (-p^-1 mod 2^32);  E += (p_even) m;  O += (p_odd) m        => E[0] == 0
    shift right by one limb: the arrays swap roles, the old E is consumed two limbs further up.
Bounds: p < 2^(32N-1) so T < 2p after every row and no carry ever leaves the odd accumulator.
"""
import sys

FIELDS = {
    "fr": dict(N=8, p=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001),
    "fq": dict(N=12, p=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab),
}
M32 = 0xFFFFFFFF


def limbs(v, n):
    return [(v >> (32 * i)) & M32 for i in range(n)]


def gen_mul(N, p):
    """abstract instruction list computing r = a*b*2^(-32N) mod p (inputs < p)"""
    P = limbs(p, N)
    inv = (-pow(p, -1, 1 << 32)) & M32
    ins = []
    X = ["x%d" % i for i in range(N)]   # offset-0 accumulator at the start
    Y = ["y%d" % i for i in range(N)]   # offset-1 accumulator
    a = ["a%d" % i for i in range(N)]
    b = ["b%d" % i for i in range(N)]

    def chain_mad(dst, src_mul, k, first_has_carry_in=False, addend=None, last_cc=True):
        """dst[2t], dst[2t+1] (+)= src_mul[t] * k along one carry chain.  addend: list of N operands
        (defaults to dst itself)."""
        add = addend if addend is not None else dst
        for t in range(N // 2):
            lo = ("madc.lo.cc" if (t > 0 or first_has_carry_in) else "mad.lo.cc")
            ins.append((lo, dst[2 * t], src_mul[t], k, add[2 * t]))
            hi = "madc.hi.cc" if (last_cc or t < N // 2 - 1) else "madc.hi"
            ins.append((hi, dst[2 * t + 1], src_mul[t], k, add[2 * t + 1]))

    a_even = [a[j] for j in range(0, N, 2)]
    a_odd = [a[j] for j in range(1, N, 2)]
    # modulus limbs enter the mad chains as REGISTER operands fed from __constant__ memory: with
    # immediates ptxas strength-reduces the special limbs (Fr: 0x00000001, 0xffffffff) into IADD3
    # forms, which breaks the lo/hi pairing of the whole chain (179 instead of 128 IMAD per Fr mul).
    p_even = ["p%d" % j for j in range(0, N, 2)]
    p_odd = ["p%d" % j for j in range(1, N, 2)]

    for i in range(N):
        if i == 0:
            for t in range(N // 2):
                ins.append(("mul.lo", X[2 * t], a_even[t], b[0]))
                ins.append(("mul.hi", X[2 * t + 1], a_even[t], b[0]))
                ins.append(("mul.lo", Y[2 * t], a_odd[t], b[0]))
                ins.append(("mul.hi", Y[2 * t + 1], a_odd[t], b[0]))
            E, O = X, Y
        else:
            # previous row left: X = offset-0 array with X[0] == 0, Y = offset-1 array.
            # new offset-0 array E = Y (+ X[1] into limb 0), new offset-1 array O = X >> 64 (+ odd products)
            ins.append(("add.cc", Y[0], Y[0], X[1]))
            shifted = [X[j + 2] for j in range(N - 2)] + ["0", "0"]
            chain_mad(X, a_odd, b[i], first_has_carry_in=True, addend=shifted, last_cc=False)
            chain_mad(Y, a_even, b[i])
            ins.append(("addc", X[N - 1], X[N - 1], "0"))
            E, O = Y, X
        ins.append(("mul.lo", "m", E[0], "pinv"))
        chain_mad(O, p_odd, "m", last_cc=False)
        chain_mad(E, p_even, "m")
        ins.append(("addc", O[N - 1], O[N - 1], "0"))
        X, Y = E, O          # X: offset 0 with X[0] == 0; Y: offset 1
    # merge: T = (X >> 32) + Y
    r = ["r%d" % i for i in range(N)]
    for k in range(N):
        src = X[k + 1] if k + 1 < N else "0"
        op = "add.cc" if k == 0 else ("addc.cc" if k < N - 1 else "addc")
        ins.append((op, r[k], Y[k], src))
    gen_final_sub(ins, N, P, r)
    return ins


def gen_final_sub(ins, N, P, r):
    """r = r >= p ? r - p : r  (r < 2p)"""
    s = ["s%d" % i for i in range(N)]
    for k in range(N):
        op = "sub.cc" if k == 0 else "subc.cc"
        ins.append((op, s[k], r[k], hex(P[k])))
    ins.append(("subc", "bw", "0", "0"))          # bw = 0xffffffff iff borrow (r < p)
    for k in range(N):
        ins.append(("selnz", r[k], r[k], s[k], "bw"))   # r = bw != 0 ? r : s


def gen_add(N, p):
    P = limbs(p, N); ins = []
    r = ["r%d" % i for i in range(N)]
    for k in range(N):
        op = "add.cc" if k == 0 else ("addc.cc" if k < N - 1 else "addc")
        ins.append((op, r[k], "a%d" % k, "b%d" % k))
    gen_final_sub(ins, N, P, r)      # a + b < 2p < 2^(32N): no carry out
    return ins


def gen_sub(N, p):
    P = limbs(p, N); ins = []
    r = ["r%d" % i for i in range(N)]
    for k in range(N):
        op = "sub.cc" if k == 0 else "subc.cc"
        ins.append((op, r[k], "a%d" % k, "b%d" % k))
    ins.append(("subc", "bw", "0", "0"))
    # add back p & mask
    for k in range(N):
        ins.append(("and", "s%d" % k, "bw", hex(P[k])))
    for k in range(N):
        op = "add.cc" if k == 0 else "addc.cc"      # the top carry is the intended wrap-around
        ins.append((op, r[k], r[k], "s%d" % k))
    return ins



# ------------------------------------------------------------------ lazy-reduction building blocks (Fq2 products)
def gen_mul_wide(N):
    """t[0..2N) = a * b, no reduction.  Even/odd accumulators X (pairs at even limbs) and Y (value Y << 32):
    row i adds a_even*b_i and a_odd*b_i as two carry chains of N/2 fused pairs each; N^2 IMAD.WIDE in all."""
    ins = []
    X = ["x%d" % i for i in range(2 * N + 2)]
    Y = ["y%d" % i for i in range(2 * N + 2)]
    wx, wy = set(), set()

    def chain(acc, written, base, mults, bi):
        """(acc[base+2t], acc[base+2t+1]) += mults[t] * bi, carry into acc[base+N]"""
        for t, m in enumerate(mults):
            lo, hi = acc[base + 2 * t], acc[base + 2 * t + 1]
            al = lo if lo in written else "0"
            ah = hi if hi in written else "0"
            ins.append(("mad.lo.cc" if t == 0 else "madc.lo.cc", lo, m, bi, al))
            ins.append(("madc.hi.cc", hi, m, bi, ah))
            written.add(lo); written.add(hi)
        top = acc[base + N]
        ins.append(("addc", top, top if top in written else "0", "0"))
        written.add(top)

    a_even = ["a%d" % j for j in range(0, N, 2)]
    a_odd = ["a%d" % j for j in range(1, N, 2)]
    for i in range(N):
        bi = "b%d" % i
        if i % 2 == 0:
            chain(X, wx, i, a_even, bi)          # limbs i + 2t (even)
            chain(Y, wy, i, a_odd, bi)           # limbs i + 2t + 1 -> Y index i + 2t
        else:
            chain(Y, wy, i - 1, a_even, bi)      # limbs i + 2t (odd) -> Y index i - 1 + 2t
            chain(X, wx, i + 1, a_odd, bi)       # limbs i + 1 + 2t (even)
    # merge: t[k] = X[k] + Y[k-1]
    for k in range(2 * N):
        xs = X[k] if X[k] in wx else "0"
        ys = (Y[k - 1] if (k >= 1 and Y[k - 1] in wy) else "0")
        op = "add.cc" if k == 0 else ("addc.cc" if k < 2 * N - 1 else "addc")
        ins.append((op, "t%d" % k, xs, ys))
    return ins


def gen_redc(N, p):
    """r = (t_lo + m p) / 2^(32N) + t_hi, conditionally reduced: Montgomery reduction of a 2N-limb t < p * 2^(32N).
    Same even/odd row structure as gen_mul with the a*b_i products removed (N^2 IMAD.WIDE)."""
    P = limbs(p, N)
    ins = []
    X = ["x%d" % i for i in range(N)]
    Y = ["y%d" % i for i in range(N)]
    p_even = ["p%d" % j for j in range(0, N, 2)]
    p_odd = ["p%d" % j for j in range(1, N, 2)]
    for i in range(N):
        if i == 0:
            for k in range(N):
                ins.append(("mov", X[k], "t%d" % k))
            E, O = X, Y
            ins.append(("mul.lo", "m", E[0], "pinv"))
            for t in range(N // 2):
                ins.append(("mul.lo", O[2 * t], p_odd[t], "m"))
                ins.append(("mul.hi", O[2 * t + 1], p_odd[t], "m"))
        else:
            # X: offset-0 array with X[0] == 0, Y: offset-1 array.  New offset-0 array E = Y (+ X[1] into limb 0),
            # new offset-1 array O = X >> 64.  m only needs E[0], so the shift of X (and the carry out of E[0]) is
            # folded into the p_odd * m chain instead of being propagated by a separate chain of additions.
            ins.append(("add.cc", Y[0], Y[0], X[1]))
            E, O = Y, X
            ins.append(("mul.lo", "m", E[0], "pinv"))            # does not touch the carry flag
            for t in range(N // 2):
                a_lo = X[2 * t + 2] if 2 * t + 2 < N else "0"
                a_hi = X[2 * t + 3] if 2 * t + 3 < N else "0"
                ins.append(("madc.lo.cc", O[2 * t], p_odd[t], "m", a_lo))
                hi = "madc.hi.cc" if t < N // 2 - 1 else "madc.hi"
                ins.append((hi, O[2 * t + 1], p_odd[t], "m", a_hi))
        for t in range(N // 2):
            lo = "mad.lo.cc" if t == 0 else "madc.lo.cc"
            ins.append((lo, E[2 * t], p_even[t], "m", E[2 * t]))
            ins.append(("madc.hi.cc", E[2 * t + 1], p_even[t], "m", E[2 * t + 1]))
        ins.append(("addc", O[N - 1], O[N - 1], "0"))
        X, Y = E, O
    r = ["r%d" % i for i in range(N)]
    for k in range(N):
        src = X[k + 1] if k + 1 < N else "0"
        op = "add.cc" if k == 0 else ("addc.cc" if k < N - 1 else "addc")
        ins.append((op, r[k], Y[k], src))
    for k in range(N):                                   # + t_hi
        op = "add.cc" if k == 0 else ("addc.cc" if k < N - 1 else "addc")
        ins.append((op, r[k], r[k], "t%d" % (N + k)))
    gen_final_sub(ins, N, P, r)
    return ins


def gen_wide_op(N, kind, p):
    """2N-limb r = a - b + p^2 ("subp2"), a - b + 2 p^2 ("sub2p2"), a - b ("sub"), a + b ("add2"), or N-limb
    unreduced r = a + b ("addn")."""
    ins = []
    if kind in ("addn", "add2"):
        M = N if kind == "addn" else 2 * N
        for k in range(M):
            op = "add.cc" if k == 0 else ("addc.cc" if k < M - 1 else "addc")
            ins.append((op, "r%d" % k, "a%d" % k, "b%d" % k))
        return ins
    M = 2 * N
    for k in range(M):
        op = "sub.cc" if k == 0 else "subc.cc"
        ins.append((op, "r%d" % k, "a%d" % k, "b%d" % k))
    if kind in ("subp2", "sub2p2"):
        P2 = limbs((1 if kind == "subp2" else 2) * p * p, M)
        for k in range(M):
            op = "add.cc" if k == 0 else "addc.cc"
            ins.append((op, "r%d" % k, "r%d" % k, hex(P2[k])))
    return ins

# ------------------------------------------------------------------ emulator
def emulate(ins, env, strict=True):
    """Execute the abstract stream on 32-bit registers with one carry flag (PTX CC.CF semantics)."""
    cf = 0

    def val(x):
        if x.startswith("0x"):
            return int(x, 16)
        if x == "0":
            return 0
        return env[x]
    for t in ins:
        op = t[0]
        if op in ("mul.lo", "mul.hi"):
            pr = val(t[2]) * val(t[3])
            env[t[1]] = (pr & M32) if op == "mul.lo" else (pr >> 32)
        elif op in ("mad.lo.cc", "madc.lo.cc", "madc.hi.cc", "madc.hi", "mad.hi.cc"):
            pr = val(t[2]) * val(t[3])
            part = (pr & M32) if ".lo" in op else (pr >> 32)
            s = part + val(t[4]) + (cf if op.startswith("madc") else 0)
            env[t[1]] = s & M32
            if op.endswith(".cc"):
                cf = s >> 32
            elif strict:
                assert s >> 32 == 0, ("dropped carry", t)
        elif op in ("add.cc", "addc.cc", "addc"):
            s = val(t[2]) + val(t[3]) + (cf if op.startswith("addc") else 0)
            env[t[1]] = s & M32
            if op.endswith(".cc"):
                cf = s >> 32
            elif strict:
                assert s >> 32 == 0, ("dropped carry", t)
        elif op in ("sub.cc", "subc.cc", "subc"):
            s = val(t[2]) - val(t[3]) - (cf if op.startswith("subc") else 0)
            env[t[1]] = s & M32
            if op.endswith(".cc"):
                cf = 1 if s < 0 else 0
        elif op == "mov":
            env[t[1]] = val(t[2])
        elif op == "selnz":
            env[t[1]] = val(t[2]) if val(t[4]) != 0 else val(t[3])
        elif op == "and":
            env[t[1]] = val(t[2]) & val(t[3])
        else:
            raise ValueError(op)
    return env


def run_emulated(kind, field, a, b):
    f = FIELDS[field]; N = f["N"]
    ins = {"mul": gen_mul, "add": gen_add, "sub": gen_sub}[kind](N, f["p"])
    env = {}
    for i, (x, y) in enumerate(zip(limbs(a, N), limbs(b, N))):
        env["a%d" % i] = x; env["b%d" % i] = y
    for i, x in enumerate(limbs(f["p"], N)):
        env["p%d" % i] = x
    env["pinv"] = (-pow(f["p"], -1, 1 << 32)) & M32
    emulate(ins, env)
    return sum(env["r%d" % i] << (32 * i) for i in range(N))


def run_emulated_wide(field, a, b):
    """(a * b) as an integer through gen_mul_wide"""
    f = FIELDS[field]; N = f["N"]
    env = {}
    for i, (x, y) in enumerate(zip(limbs(a, N), limbs(b, N))):
        env["a%d" % i] = x; env["b%d" % i] = y
    emulate(gen_mul_wide(N), env)
    return sum(env["t%d" % i] << (32 * i) for i in range(2 * N))


def run_emulated_redc(field, t):
    f = FIELDS[field]; N = f["N"]
    env = {}
    for i, x in enumerate(limbs(t, 2 * N)):
        env["t%d" % i] = x
    for i, x in enumerate(limbs(f["p"], N)):
        env["p%d" % i] = x
    env["pinv"] = (-pow(f["p"], -1, 1 << 32)) & M32
    emulate(gen_redc(N, f["p"]), env)
    return sum(env["r%d" % i] << (32 * i) for i in range(N))


def run_emulated_wide_op(field, kind, a, b):
    f = FIELDS[field]; N = f["N"]
    M = N if kind == "addn" else 2 * N
    env = {}
    for i, (x, y) in enumerate(zip(limbs(a, M), limbs(b, M))):
        env["a%d" % i] = x; env["b%d" % i] = y
    emulate(gen_wide_op(N, kind, f["p"]), env, strict=False)
    return sum(env["r%d" % i] << (32 * i) for i in range(M))


# ------------------------------------------------------------------ PTX rendering
def render(name, ins, N, modc=None):
    regs = set()
    for t in ins:
        for x in t[1:]:
            if not (x.startswith("0x") or x == "0"):
                regs.add(x)
    inputs = ["a%d" % i for i in range(N)] + ["b%d" % i for i in range(N)]
    if modc:
        inputs += ["p%d" % i for i in range(N)] + ["pinv"]
    outputs = ["r%d" % i for i in range(N)]
    temps = sorted(regs - set(inputs) - set(outputs))
    opmap = {}
    for i, r in enumerate(outputs):
        opmap[r] = "%%%d" % i
    for i, r in enumerate(inputs):
        opmap[r] = "%%%d" % (N + i)

    def o(x):
        if x.startswith("0x") or x == "0":
            return x if x != "0" else "0"
        return opmap.get(x, x)
    lines = []
    lines.append("    \"{\\n\\t\"")
    lines.append("    \".reg .u32 %s;\\n\\t\"" % ", ".join(temps))
    lines.append("    \".reg .pred pz;\\n\\t\"")
    for t in ins:
        op = t[0]
        if op in ("mul.lo", "mul.hi"):
            s = "%s.u32 %s, %s, %s;" % (op, o(t[1]), o(t[2]), o(t[3]))
        elif op.startswith("mad"):
            s = "%s.u32 %s, %s, %s, %s;" % (op, o(t[1]), o(t[2]), o(t[3]), o(t[4]))
        elif op.startswith("add") or op.startswith("sub"):
            s = "%s.u32 %s, %s, %s;" % (op, o(t[1]), o(t[2]), o(t[3]))
        elif op == "and":
            s = "and.b32 %s, %s, %s;" % (o(t[1]), o(t[2]), o(t[3]))
        elif op == "mov":
            env[t[1]] = val(t[2])
        elif op == "selnz":
            s = "setp.ne.u32 pz, %s, 0; selp.u32 %s, %s, %s, pz;" % (o(t[4]), o(t[1]), o(t[2]), o(t[3]))
        else:
            raise ValueError(op)
        lines.append("    \"%s\\n\\t\"" % s)
    lines.append("    \"}\"")
    outs = ", ".join("\"=&r\"(r[%d])" % i for i in range(N))
    ins_ = ", ".join(["\"r\"(a[%d])" % i for i in range(N)] + ["\"r\"(b[%d])" % i for i in range(N)] +
                     (["\"r\"(%s[%d])" % (modc, i) for i in range(N + 1)] if modc else []))
    body = "\n".join(lines)
    return ("__device__ __forceinline__ void %s(uint32_t* __restrict__ r, const uint32_t* a, const uint32_t* b) {\n"
            "  asm(\n%s\n    : %s\n    : %s);\n}\n" % (name, body, outs, ins_))


def render_general(name, ins, outputs, inputs, params):
    """outputs / inputs: lists of (register name, C expression); params: C parameter list"""
    regs = set()
    for t in ins:
        for x in t[1:]:
            if not (x.startswith("0x") or x == "0"):
                regs.add(x)
    onames = [o for o, _ in outputs]; inames = [i for i, _ in inputs]
    temps = sorted(regs - set(onames) - set(inames))
    opmap = {}
    for k, r in enumerate(onames):
        opmap[r] = "%%%d" % k
    for k, r in enumerate(inames):
        opmap[r] = "%%%d" % (len(onames) + k)

    def o(x):
        if x.startswith("0x") or x == "0":
            return x
        return opmap.get(x, x)
    lines = ["    \"{\\n\\t\""]
    if temps:
        lines.append("    \".reg .u32 %s;\\n\\t\"" % ", ".join(temps))
    lines.append("    \".reg .pred pz;\\n\\t\"")
    for t in ins:
        op = t[0]
        if op in ("mul.lo", "mul.hi"):
            q = "%s.u32 %s, %s, %s;" % (op, o(t[1]), o(t[2]), o(t[3]))
        elif op.startswith("mad"):
            q = "%s.u32 %s, %s, %s, %s;" % (op, o(t[1]), o(t[2]), o(t[3]), o(t[4]))
        elif op.startswith("add") or op.startswith("sub"):
            q = "%s.u32 %s, %s, %s;" % (op, o(t[1]), o(t[2]), o(t[3]))
        elif op == "and":
            q = "and.b32 %s, %s, %s;" % (o(t[1]), o(t[2]), o(t[3]))
        elif op == "mov":
            q = "mov.u32 %s, %s;" % (o(t[1]), o(t[2]))
        elif op == "selnz":
            q = "setp.ne.u32 pz, %s, 0; selp.u32 %s, %s, %s, pz;" % (o(t[4]), o(t[1]), o(t[2]), o(t[3]))
        else:
            raise ValueError(op)
        lines.append("    \"%s\\n\\t\"" % q)
    lines.append("    \"}\"")
    outs = ", ".join("\"=&r\"(%s)" % e for _, e in outputs)
    inps = ", ".join("\"r\"(%s)" % e for _, e in inputs)
    return ("__device__ __forceinline__ void %s(%s) {\n  asm(\n%s\n    : %s\n    : %s);\n}\n"
            % (name, params, "\n".join(lines), outs, inps))


def render_lazy(fname, f):
    """mul_wide / redc / wide add-sub helpers used by the lazy-reduction Fq2 product"""
    N, p, U = f["N"], f["p"], fname.upper()
    parts = []
    parts.append(render_general("%s_mul_wide_ptx" % fname, gen_mul_wide(N),
                                [("t%d" % k, "t[%d]" % k) for k in range(2 * N)],
                                [("a%d" % k, "a[%d]" % k) for k in range(N)] + [("b%d" % k, "b[%d]" % k) for k in range(N)],
                                "uint32_t* __restrict__ t, const uint32_t* a, const uint32_t* b"))
    parts.append(render_general("%s_redc_ptx" % fname, gen_redc(N, p),
                                [("r%d" % k, "r[%d]" % k) for k in range(N)],
                                [("t%d" % k, "t[%d]" % k) for k in range(2 * N)] + [("p%d" % k, "%s_MOD_C[%d]" % (U, k)) for k in range(N)] +
                                [("pinv", "%s_MOD_C[%d]" % (U, N))],
                                "uint32_t* __restrict__ r, const uint32_t* t"))
    for kind, M in (("subp2", 2 * N), ("sub", 2 * N), ("addn", N)):
        parts.append(render_general("%s_wide_%s_ptx" % (fname, kind), gen_wide_op(N, kind, p),
                                    [("r%d" % k, "r[%d]" % k) for k in range(M)],
                                    [("a%d" % k, "a[%d]" % k) for k in range(M)] + [("b%d" % k, "b[%d]" % k) for k in range(M)],
                                    "uint32_t* __restrict__ r, const uint32_t* a, const uint32_t* b"))
    return "\n".join(parts)


def main(out_path):
    parts = ["// GENERATED by tools/gen_field.py -- do not edit.  Unrolled PTX carry-chain field arithmetic\n"
             "// (32-bit limbs, Montgomery radix 2^(32 N)) for BLS12-381 Fr (N=8) and Fq (N=12).\n"
             "// The same instruction streams are verified against big integers by tests/test_cpu_oracle.py (test_generated_field_streams_under_emulation, test_lazy_reduction_sequences_under_emulation).\n"
             "#pragma once\n#include <cstdint>\n"]
    for fname, f in FIELDS.items():
        N = f["N"]; p = f["p"]; R = 1 << (32 * N)
        def arr(v):
            return "{" + ", ".join("0x%08xu" % x for x in limbs(v, N)) + "}"
        U = fname.upper()
        parts.append("#define %s_LIMBS %d" % (U, N))
        parts.append("#define %s_MOD_INIT %s" % (U, arr(p)))
        parts.append("#define %s_R1_INIT %s   /* R mod p: Montgomery one */" % (U, arr(R % p)))
        parts.append("#define %s_R2_INIT %s   /* R^2 mod p */" % (U, arr(R * R % p)))
        parts.append("#define %s_INV32 0x%08xu   /* -p^-1 mod 2^32 */\n" % (U, (-pow(p, -1, 1 << 32)) & M32))
    parts.append("#if defined(__CUDACC__)\n")
    for fname, f in FIELDS.items():
        N = f["N"]
        # outputs may alias inputs at the call site: the wrappers in fp.cuh copy through temporaries
        parts.append("// modulus limbs followed by -p^-1 mod 2^32; deliberately not const (see gen_field.py)\n"
                     "static __device__ __constant__ uint32_t %s_MOD_C[%d] = {%s, 0x%08xu};\n"
                     % (fname.upper(), N + 1, ", ".join("0x%08xu" % x for x in limbs(f["p"], N)), (-pow(f["p"], -1, 1 << 32)) & M32))
        parts.append(render("%s_mul_ptx" % fname, gen_mul(N, f["p"]), N, modc="%s_MOD_C" % fname.upper()))
        parts.append(render("%s_add_ptx" % fname, gen_add(N, f["p"]), N))
        parts.append(render("%s_sub_ptx" % fname, gen_sub(N, f["p"]), N))
        if fname == "fq":
            parts.append(render_lazy(fname, f))
    parts.append("#endif  // __CUDACC__\n")
    open(out_path, "w").write("\n".join(parts))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "fp_gen.cuh")
