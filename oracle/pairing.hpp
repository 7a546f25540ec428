// oracle/pairing.hpp -- a bilinear pairing on BLS12-381 for the CPU verifier.  TEST INFRASTRUCTURE ONLY.
//
// Used by the oracle's restatement of MLPolyCommit::verify (/root/reference/src/commitment/verify.rs:12-45),
// which checks  e(C - g^eval, h) == prod_i e(g^{t_i} - g^{p_i}, pi_i).  Both sides use the same map, so ANY
// bilinear non-degenerate pairing decides the equation exactly as arkworks' optimal-ate pairing does; this one is
// the plain ate pairing f_{|x|, Q}(P)^((p^12 - 1) / r) computed the slow, convention-free way: the G2 point is
// untwisted into E(Fq12) and the Miller loop runs in affine coordinates over Fq12 with textbook line functions.
// (arkworks' PairingEngine is upstream code that is not on this machine; PARITY UNPINNED, see README.md.)
// Bilinearity and non-degeneracy are checked by tests/test_cpu_oracle.py.
#pragma once
#include "ff.hpp"
#include "ec.hpp"

// Fq6 = Fq2[v] / (v^3 - xi), xi = 1 + u
struct Fq6 {
    Fq2 c0, c1, c2;
    static Fq2 mul_xi(const Fq2& a) {          // a * (1 + u) = (a0 - a1) + (a0 + a1) u
        Fq2 o; o.c0 = Fq::sub(a.c0, a.c1); o.c1 = Fq::add(a.c0, a.c1); return o;
    }
    static Fq6 zero() { Fq6 z; z.c0 = Fq2::zero(); z.c1 = Fq2::zero(); z.c2 = Fq2::zero(); return z; }
    static Fq6 one() { Fq6 z = zero(); z.c0 = Fq2::one(); return z; }
    static Fq6 add(const Fq6& a, const Fq6& b) { Fq6 o; o.c0 = Fq2::add(a.c0, b.c0); o.c1 = Fq2::add(a.c1, b.c1); o.c2 = Fq2::add(a.c2, b.c2); return o; }
    static Fq6 sub(const Fq6& a, const Fq6& b) { Fq6 o; o.c0 = Fq2::sub(a.c0, b.c0); o.c1 = Fq2::sub(a.c1, b.c1); o.c2 = Fq2::sub(a.c2, b.c2); return o; }
    static Fq6 neg(const Fq6& a) { return sub(zero(), a); }
    static Fq6 mul(const Fq6& a, const Fq6& b) {   // schoolbook, v^3 = xi
        Fq2 t00 = Fq2::mul(a.c0, b.c0), t01 = Fq2::mul(a.c0, b.c1), t02 = Fq2::mul(a.c0, b.c2);
        Fq2 t10 = Fq2::mul(a.c1, b.c0), t11 = Fq2::mul(a.c1, b.c1), t12 = Fq2::mul(a.c1, b.c2);
        Fq2 t20 = Fq2::mul(a.c2, b.c0), t21 = Fq2::mul(a.c2, b.c1), t22 = Fq2::mul(a.c2, b.c2);
        Fq6 o;
        o.c0 = Fq2::add(t00, mul_xi(Fq2::add(t12, t21)));
        o.c1 = Fq2::add(Fq2::add(t01, t10), mul_xi(t22));
        o.c2 = Fq2::add(Fq2::add(t02, t11), t20);
        return o;
    }
    static Fq6 mul_v(const Fq6& a) { Fq6 o; o.c0 = mul_xi(a.c2); o.c1 = a.c0; o.c2 = a.c1; return o; }
    static Fq6 inv(const Fq6& a) {
        // standard: with A = c0^2 - xi c1 c2, B = xi c2^2 - c0 c1, C = c1^2 - c0 c2, F = c0 A + xi (c2 B + c1 C)
        Fq2 A = Fq2::sub(Fq2::sqr(a.c0), mul_xi(Fq2::mul(a.c1, a.c2)));
        Fq2 B = Fq2::sub(mul_xi(Fq2::sqr(a.c2)), Fq2::mul(a.c0, a.c1));
        Fq2 C = Fq2::sub(Fq2::sqr(a.c1), Fq2::mul(a.c0, a.c2));
        Fq2 F = Fq2::add(Fq2::mul(a.c0, A), mul_xi(Fq2::add(Fq2::mul(a.c2, B), Fq2::mul(a.c1, C))));
        Fq2 Fi = Fq2::inv(F);
        Fq6 o; o.c0 = Fq2::mul(A, Fi); o.c1 = Fq2::mul(B, Fi); o.c2 = Fq2::mul(C, Fi); return o;
    }
    bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
};

// Fq12 = Fq6[w] / (w^2 - v)
struct Fq12 {
    Fq6 c0, c1;
    static Fq12 zero() { Fq12 z; z.c0 = Fq6::zero(); z.c1 = Fq6::zero(); return z; }
    static Fq12 one() { Fq12 z; z.c0 = Fq6::one(); z.c1 = Fq6::zero(); return z; }
    static Fq12 add(const Fq12& a, const Fq12& b) { Fq12 o; o.c0 = Fq6::add(a.c0, b.c0); o.c1 = Fq6::add(a.c1, b.c1); return o; }
    static Fq12 sub(const Fq12& a, const Fq12& b) { Fq12 o; o.c0 = Fq6::sub(a.c0, b.c0); o.c1 = Fq6::sub(a.c1, b.c1); return o; }
    static Fq12 dbl(const Fq12& a) { return add(a, a); }
    static Fq12 neg(const Fq12& a) { return sub(zero(), a); }
    static Fq12 mul(const Fq12& a, const Fq12& b) {
        Fq6 t0 = Fq6::mul(a.c0, b.c0), t1 = Fq6::mul(a.c1, b.c1);
        Fq12 o;
        o.c0 = Fq6::add(t0, Fq6::mul_v(t1));
        o.c1 = Fq6::add(Fq6::mul(a.c0, b.c1), Fq6::mul(a.c1, b.c0));
        return o;
    }
    static Fq12 sqr(const Fq12& a) { return mul(a, a); }
    static Fq12 conj(const Fq12& a) { Fq12 o; o.c0 = a.c0; o.c1 = Fq6::neg(a.c1); return o; }   // = a^(p^6)
    static Fq12 inv(const Fq12& a) {
        Fq6 d = Fq6::sub(Fq6::mul(a.c0, a.c0), Fq6::mul_v(Fq6::mul(a.c1, a.c1)));
        Fq6 di = Fq6::inv(d);
        Fq12 o; o.c0 = Fq6::mul(a.c0, di); o.c1 = Fq6::neg(Fq6::mul(a.c1, di)); return o;
    }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fq12& o) const { return !(*this == o); }
    static Fq12 from_fq(const Fq& a) { Fq12 z = zero(); z.c0.c0.c0 = a; return z; }
    static Fq12 from_fq2(const Fq2& a) { Fq12 z = zero(); z.c0.c0 = a; return z; }
    static Fq12 w() { Fq12 z = zero(); z.c1.c0 = Fq2::one(); return z; }
    // a^e, e a big-endian hex string
    static Fq12 pow_hex(const Fq12& a, const char* hex) {
        Fq12 acc = one();
        for (const char* ch = hex; *ch; ch++) {
            int d = (*ch >= '0' && *ch <= '9') ? *ch - '0' : *ch - 'a' + 10;
            for (int b = 3; b >= 0; b--) {
                acc = sqr(acc);
                if ((d >> b) & 1) acc = mul(acc, a);
            }
        }
        return acc;
    }
};

// (p^4 - p^2 + 1) / r and p^2 + 1 (computed with Python from the BLS parameter x = -0xd201000000010000)
static const char* PAIRING_EXP_HARD =
    "f686b3d807d01c0bd38c3195c899ed3cde88eeb996ca394506632528d6a9a2f230063cf081517f68f7764c28b6f8ae5a72bce8d63cb9f827eca0ba621315b2076995003fc77a17988f8761bdc51dc2378b9039096d1b767f17fcbde783765915c97f36c6f18212ed0b283ed237db421d160aeb6a1e79983774940996754c8c71a2629b0dea236905ce937335d5b68fa9912aae208ccf1e516c3f438e3ba79";
static const char* PAIRING_EXP_P2_PLUS_1 =
    "2a437a4b8c35fc74bd278eaa22f25e9e2dc90e50e7046b466e59e49349e8bd050a62cfd16ddca6ef53149330978ef011d68619c86185c7b292e85a87091a04966bf91ed3e71b743162c338362113cfd7ced6b1d76382eab26aa00001c718e3a";

// Miller loop f_{|x|, psi(Q)}(P) with psi: E'(Fq2) -> E(Fq12), (x', y') -> (x' / w^2, y' / w^3)  (w^6 = xi)
static Fq12 miller_loop(const G1Affine& P, const G2Affine& Q) {
    if (P.is_inf() || Q.is_inf()) return Fq12::one();
    const Fq12 w = Fq12::w();
    const Fq12 w2 = Fq12::mul(w, w), w3 = Fq12::mul(w2, w);
    const Fq12 xq = Fq12::mul(Fq12::from_fq2(Q.x), Fq12::inv(w2));
    const Fq12 yq = Fq12::mul(Fq12::from_fq2(Q.y), Fq12::inv(w3));
    const Fq12 xp = Fq12::from_fq(P.x), yp = Fq12::from_fq(P.y);
    Fq12 xt = xq, yt = yq, f = Fq12::one();
    bool t_inf = false;
    const uint64_t X = 0xd201000000010000ULL;
    const Fq12 three = Fq12::from_fq(Fq::from_u64(3));
    for (int i = 62; i >= 0; i--) {
        f = Fq12::sqr(f);
        if (!t_inf) {
            // tangent at T
            Fq12 lam = Fq12::mul(Fq12::mul(three, Fq12::sqr(xt)), Fq12::inv(Fq12::dbl(yt)));
            f = Fq12::mul(f, Fq12::sub(Fq12::sub(yp, yt), Fq12::mul(lam, Fq12::sub(xp, xt))));
            Fq12 x3 = Fq12::sub(Fq12::sqr(lam), Fq12::dbl(xt));
            Fq12 y3 = Fq12::sub(Fq12::mul(lam, Fq12::sub(xt, x3)), yt);
            xt = x3; yt = y3;
        }
        if ((X >> i) & 1) {
            if (t_inf) { xt = xq; yt = yq; t_inf = false; continue; }
            if (xt == xq) {                      // T = -Q (T = Q cannot happen mid-loop for a point of order r)
                f = Fq12::mul(f, Fq12::sub(xp, xt));
                t_inf = true; continue;
            }
            Fq12 lam = Fq12::mul(Fq12::sub(yq, yt), Fq12::inv(Fq12::sub(xq, xt)));
            f = Fq12::mul(f, Fq12::sub(Fq12::sub(yp, yt), Fq12::mul(lam, Fq12::sub(xp, xt))));
            Fq12 x3 = Fq12::sub(Fq12::sub(Fq12::sqr(lam), xt), xq);
            Fq12 y3 = Fq12::sub(Fq12::mul(lam, Fq12::sub(xt, x3)), yt);
            xt = x3; yt = y3;
        }
    }
    return f;
}
static Fq12 final_exponentiation(const Fq12& f) {
    Fq12 t = Fq12::mul(Fq12::conj(f), Fq12::inv(f));        // f^(p^6 - 1)
    t = Fq12::pow_hex(t, PAIRING_EXP_P2_PLUS_1);            // ^(p^2 + 1)
    return Fq12::pow_hex(t, PAIRING_EXP_HARD);              // ^((p^4 - p^2 + 1) / r)
}
static Fq12 pairing(const G1Affine& P, const G2Affine& Q) { return final_exponentiation(miller_loop(P, Q)); }
