// ec.cuh -- y^2 = x^3 + b (a = 0) group arithmetic in extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), templated on the coordinate field (Fq for G1, Fq2 for G2).
//
// The reference does its group arithmetic inside arkworks (`VariableBaseMSM::multi_scalar_mul`,
// src/commitment/commit.rs:25, open.rs:49; Jacobian coordinates).  Group elements are exact, so the
// affine result -- the only thing that is serialized -- does not depend on the coordinate system;
// XYZZ is used here because the bucket-accumulation inner loop is a mixed addition (8M + 2S).
#pragma once
#include "field.cuh"

// The point operations are deliberately NOT inlined on the device: one G2 addition is ~13k SASS
// instructions, and kernels that use several of them would otherwise take ptxas tens of minutes.
// Operands travel through local memory (a few hundred LD/ST against thousands of IMADs per call).
#if defined(__CUDACC__)
#define SB_EC_FN __host__ __device__ __noinline__
#else
#define SB_EC_FN
#endif

template <class F>
struct alignas(16) AffinePt {
    F x, y;    // (0, 0) = point at infinity (never on the curve since b != 0)
    SB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    SB_HD static AffinePt inf() { AffinePt p; p.x = F::zero(); p.y = F::zero(); return p; }
    SB_HD bool operator==(const AffinePt& o) const { return x == o.x && y == o.y; }
};

template <class F>
struct alignas(16) XyzzPt {
    F X, Y, ZZ, ZZZ;
    SB_HD static XyzzPt inf() { XyzzPt p; p.X = F::zero(); p.Y = F::zero(); p.ZZ = F::zero(); p.ZZZ = F::zero(); return p; }
    SB_HD bool is_inf() const { return ZZ.is_zero(); }
    SB_HD static XyzzPt from_affine(const AffinePt<F>& a) {
        if (a.is_inf()) return inf();
        XyzzPt p; p.X = a.x; p.Y = a.y; p.ZZ = F::one(); p.ZZZ = F::one(); return p;
    }
    SB_HD static XyzzPt neg(const XyzzPt& p) { XyzzPt o = p; o.Y = F::neg(p.Y); return o; }

    // dbl-2008-s-1
    SB_EC_FN static XyzzPt dbl(const XyzzPt& p) {
        if (p.is_inf()) return p;
        F U = F::dbl(p.Y), V = F::sqr(U), W = F::mul(U, V), S = F::mul(p.X, V);
        F XX = F::sqr(p.X), M = F::add(F::dbl(XX), XX);
        XyzzPt o;
        o.X = F::sub(F::sqr(M), F::dbl(S));
        o.Y = F::sub(F::mul(M, F::sub(S, o.X)), F::mul(W, p.Y));
        o.ZZ = F::mul(V, p.ZZ);
        o.ZZZ = F::mul(W, p.ZZZ);
        return o;
    }
    // mdbl-2008-s-1: 2 * affine
    SB_EC_FN static XyzzPt dbl_affine(const AffinePt<F>& a) {
        if (a.is_inf()) return inf();
        F U = F::dbl(a.y), V = F::sqr(U), W = F::mul(U, V), S = F::mul(a.x, V);
        F XX = F::sqr(a.x), M = F::add(F::dbl(XX), XX);
        XyzzPt o;
        o.X = F::sub(F::sqr(M), F::dbl(S));
        o.Y = F::sub(F::mul(M, F::sub(S, o.X)), F::mul(W, a.y));
        o.ZZ = V; o.ZZZ = W;
        return o;
    }
    // madd-2008-s with the exceptional cases (acc = inf, q = inf, q = +-acc) handled exactly
    SB_EC_FN static XyzzPt add_mixed(const XyzzPt& p, const AffinePt<F>& q) {
        if (q.is_inf()) return p;
        if (p.is_inf()) return from_affine(q);
        F U2 = F::mul(q.x, p.ZZ), S2 = F::mul(q.y, p.ZZZ);
        F Pp = F::sub(U2, p.X), R = F::sub(S2, p.Y);
        if (Pp.is_zero()) {
            if (R.is_zero()) return dbl_affine(q);
            return inf();
        }
        F PP = F::sqr(Pp), PPP = F::mul(Pp, PP), Q = F::mul(p.X, PP);
        XyzzPt o;
        o.X = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
        o.Y = F::sub(F::mul(R, F::sub(Q, o.X)), F::mul(p.Y, PPP));
        o.ZZ = F::mul(p.ZZ, PP);
        o.ZZZ = F::mul(p.ZZZ, PPP);
        return o;
    }
    // add-2008-s
    SB_EC_FN static XyzzPt add(const XyzzPt& p, const XyzzPt& q) {
        if (p.is_inf()) return q;
        if (q.is_inf()) return p;
        F U1 = F::mul(p.X, q.ZZ), U2 = F::mul(q.X, p.ZZ);
        F S1 = F::mul(p.Y, q.ZZZ), S2 = F::mul(q.Y, p.ZZZ);
        F Pp = F::sub(U2, U1), R = F::sub(S2, S1);
        if (Pp.is_zero()) {
            if (R.is_zero()) return dbl(p);
            return inf();
        }
        F PP = F::sqr(Pp), PPP = F::mul(Pp, PP), Q = F::mul(U1, PP);
        XyzzPt o;
        o.X = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
        o.Y = F::sub(F::mul(R, F::sub(Q, o.X)), F::mul(S1, PPP));
        o.ZZ = F::mul(F::mul(p.ZZ, q.ZZ), PP);
        o.ZZZ = F::mul(F::mul(p.ZZZ, q.ZZZ), PPP);
        return o;
    }
};

typedef AffinePt<Fq> G1Aff;
typedef AffinePt<Fq2> G2Aff;
typedef XyzzPt<Fq> G1Xyzz;
typedef XyzzPt<Fq2> G2Xyzz;

// host-side finishing: XYZZ -> affine (one field inversion)
template <class F>
static inline AffinePt<F> xyzz_to_affine_host(const XyzzPt<F>& p) {
    if (p.is_inf()) return AffinePt<F>::inf();
    F zi3 = F::inv_host(p.ZZZ);             // 1/ZZZ
    F zi2 = F::mul(F::sqr(zi3), F::sqr(p.ZZ)); // ZZ^2/ZZZ^2 = 1/ZZ   (ZZ^3 = ZZZ^2)
    AffinePt<F> a; a.x = F::mul(p.X, zi2); a.y = F::mul(p.Y, zi3);
    return a;
}
