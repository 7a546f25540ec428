// oracle/ec.hpp -- short-Weierstrass (a = 0) group arithmetic for the parity oracle.
// TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see oracle/README.md).
//
// Restates the published algorithms arkworks' `ark_ec::models::short_weierstrass_jacobian` uses:
// Jacobian doubling (dbl-2009-l), mixed addition (madd-2007-bl) and full addition (add-2007-bl),
// and `ark_ec::msm::VariableBaseMSM::multi_scalar_mul` (serial bucket method,
// window c = 3 if n < 32 else ln(n) + 2).  Every result here is a group element, so the affine
// output is independent of the formulas chosen.
#pragma once
#include "ff.hpp"
#include <vector>
#include <cmath>

template <typename F>
struct Affine {
    F x, y;          // (0, 0) encodes the point at infinity (never on y^2 = x^3 + b, b != 0)
    inline bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static inline Affine inf() { Affine a; a.x = F::zero(); a.y = F::zero(); return a; }
    inline bool operator==(const Affine& o) const { return x == o.x && y == o.y; }
};

template <typename F>
struct Jac {
    F X, Y, Z;
    static inline Jac inf() { Jac j; j.X = F::one(); j.Y = F::one(); j.Z = F::zero(); return j; }
    inline bool is_inf() const { return Z.is_zero(); }
    static inline Jac from_affine(const Affine<F>& a) {
        if (a.is_inf()) return inf();
        Jac j; j.X = a.x; j.Y = a.y; j.Z = F::one(); return j;
    }
    static Jac dbl(const Jac& p) {
        if (p.is_inf()) return p;
        F A = F::sqr(p.X), B = F::sqr(p.Y), C = F::sqr(B);
        F t = F::add(p.X, B);
        F D = F::dbl(F::sub(F::sub(F::sqr(t), A), C));
        F E = F::add(F::dbl(A), A);
        F Fv = F::sqr(E);
        Jac o;
        o.X = F::sub(Fv, F::dbl(D));
        F C8 = F::dbl(F::dbl(F::dbl(C)));
        o.Y = F::sub(F::mul(E, F::sub(D, o.X)), C8);
        o.Z = F::mul(F::dbl(p.Y), p.Z);
        return o;
    }
    static Jac add(const Jac& p, const Jac& q) {
        if (p.is_inf()) return q;
        if (q.is_inf()) return p;
        F Z1Z1 = F::sqr(p.Z), Z2Z2 = F::sqr(q.Z);
        F U1 = F::mul(p.X, Z2Z2), U2 = F::mul(q.X, Z1Z1);
        F S1 = F::mul(F::mul(p.Y, q.Z), Z2Z2), S2 = F::mul(F::mul(q.Y, p.Z), Z1Z1);
        if (U1 == U2) {
            if (S1 == S2) return dbl(p);
            return inf();
        }
        F H = F::sub(U2, U1), Rr = F::sub(S2, S1);
        F HH = F::sqr(H), HHH = F::mul(H, HH), V = F::mul(U1, HH);
        Jac o;
        o.X = F::sub(F::sub(F::sqr(Rr), HHH), F::dbl(V));
        o.Y = F::sub(F::mul(Rr, F::sub(V, o.X)), F::mul(S1, HHH));
        o.Z = F::mul(F::mul(p.Z, q.Z), H);
        return o;
    }
    static Jac add_mixed(const Jac& p, const Affine<F>& q) {
        if (q.is_inf()) return p;
        if (p.is_inf()) return from_affine(q);
        F Z1Z1 = F::sqr(p.Z);
        F U2 = F::mul(q.x, Z1Z1);
        F S2 = F::mul(F::mul(q.y, p.Z), Z1Z1);
        if (p.X == U2) {
            if (p.Y == S2) return dbl(p);
            return inf();
        }
        F H = F::sub(U2, p.X), Rr = F::sub(S2, p.Y);
        F HH = F::sqr(H), HHH = F::mul(H, HH), V = F::mul(p.X, HH);
        Jac o;
        o.X = F::sub(F::sub(F::sqr(Rr), HHH), F::dbl(V));
        o.Y = F::sub(F::mul(Rr, F::sub(V, o.X)), F::mul(p.Y, HHH));
        o.Z = F::mul(p.Z, H);
        return o;
    }
    static inline Jac neg(const Jac& p) { Jac o = p; o.Y = F::neg(p.Y); return o; }
    Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        F zi = F::inv(Z), zi2 = F::sqr(zi);
        Affine<F> a; a.x = F::mul(X, zi2); a.y = F::mul(Y, F::mul(zi2, zi)); return a;
    }
    // k * p, k given as canonical little-endian limbs
    static Jac mul(const Jac& p, const uint64_t* k, int nlimbs) {
        Jac acc = inf();
        for (int i = nlimbs * 64 - 1; i >= 0; i--) {
            acc = dbl(acc);
            if ((k[i / 64] >> (i % 64)) & 1) acc = add(acc, p);
        }
        return acc;
    }
};

// ark_ec batch_normalization_into_affine: Montgomery's simultaneous inversion
// Worker threads of the heavy loops.  1 (the default) is the reference's configuration (Cargo.toml:26 never enables
// `parallel`) and the only setting a CPU-baseline timing may use; or_set_threads(k > 1) is for the parity TESTS at the
// full sizes (2^16, 2^20), where only the results matter -- every loop below is exact field / group arithmetic
// whose result does not depend on the order of summation, so the bytes are identical for any thread count.
static int g_oracle_threads = 1;

template <typename F>
static void batch_to_affine_range(const std::vector<Jac<F>>& in, std::vector<Affine<F>>& out, size_t lo, size_t hi) {
    std::vector<F> prefix(hi - lo);
    F acc = F::one();
    for (size_t i = lo; i < hi; i++) {
        prefix[i - lo] = acc;
        if (!in[i].is_inf()) acc = F::mul(acc, in[i].Z);
    }
    F inv = F::inv(acc);
    for (size_t i = hi; i-- > lo;) {
        if (in[i].is_inf()) { out[i] = Affine<F>::inf(); continue; }
        F zi = F::mul(inv, prefix[i - lo]);
        inv = F::mul(inv, in[i].Z);
        F zi2 = F::sqr(zi);
        out[i].x = F::mul(in[i].X, zi2);
        out[i].y = F::mul(in[i].Y, F::mul(zi2, zi));
    }
}
template <typename F>
static void batch_to_affine(const std::vector<Jac<F>>& in, std::vector<Affine<F>>& out) {
    size_t n = in.size();
    out.resize(n);
    const size_t parts = (g_oracle_threads > 1 && n >= 4096) ? (size_t)g_oracle_threads : 1;
#pragma omp parallel for num_threads(g_oracle_threads) if (parts > 1) schedule(static)
    for (size_t k = 0; k < parts; k++) batch_to_affine_range(in, out, n * k / parts, n * (k + 1) / parts);
}

// UPSTREAM ark_ec::msm::VariableBaseMSM::multi_scalar_mul (late-2020 shape): scalars are canonical
// 4-limb integers ("into_repr", commit.rs:20-21 / open.rs:46); zips and truncates to the shorter.
static inline int ark_ln_without_floats(size_t a) {
    int lg = 0; while ((size_t(1) << lg) < a) lg++;    // ark_std::log2 = ceil(log2)
    return lg * 69 / 100;
}
template <typename F>
static Jac<F> msm_pippenger(const Affine<F>* bases, const uint64_t* scalars /* n x 4 */, size_t n) {
    int c = n < 32 ? 3 : ark_ln_without_floats(n) + 2;
    const int num_bits = 255;
    Jac<F> total = Jac<F>::inf();
    const int n_windows = (num_bits + c - 1) / c;
    std::vector<Jac<F>> window_sums(n_windows);
    // the windows are independent of one another (upstream runs them in a plain loop without `parallel`)
#pragma omp parallel for num_threads(g_oracle_threads) if (g_oracle_threads > 1 && n >= 256) schedule(dynamic, 1)
    for (int wi = 0; wi < n_windows; wi++) {
        const int w_start = wi * c;
        Jac<F> res = Jac<F>::inf();
        std::vector<Jac<F>> buckets((size_t(1) << c) - 1, Jac<F>::inf());
        for (size_t i = 0; i < n; i++) {
            const uint64_t* s = scalars + 4 * i;
            if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;
            bool is_one = (s[0] == 1 && (s[1] | s[2] | s[3]) == 0);
            if (is_one) {
                if (w_start == 0) res = Jac<F>::add_mixed(res, bases[i]);
                continue;
            }
            // bits [w_start, w_start + c)
            int limb = w_start / 64, off = w_start % 64;
            uint64_t v = s[limb] >> off;
            if (off + c > 64 && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
            v &= (uint64_t(1) << c) - 1;
            if (v != 0) buckets[v - 1] = Jac<F>::add_mixed(buckets[v - 1], bases[i]);
        }
        Jac<F> running = Jac<F>::inf();
        for (size_t b = buckets.size(); b-- > 0;) {
            running = Jac<F>::add(running, buckets[b]);
            res = Jac<F>::add(res, running);
        }
        window_sums[wi] = res;
    }
    // lowest window + sum_{w>0} 2^{cw} * window_w, Horner from the top
    Jac<F> acc = Jac<F>::inf();
    for (size_t w = window_sums.size(); w-- > 1;) {
        acc = Jac<F>::add(acc, window_sums[w]);
        for (int i = 0; i < c; i++) acc = Jac<F>::dbl(acc);
    }
    total = Jac<F>::add(acc, window_sums[0]);
    return total;
}

// UPSTREAM ark_ec::msm::FixedBaseMSM (windowed table): out[i] = scalars[i] * g.  Any method gives
// the same group elements; a 2^w-entry table per window keeps keygen O(255/w) additions per scalar.
template <typename F>
static void fixed_base_mul(const Affine<F>& g, const uint64_t* scalars /* n x 4 canonical */, size_t n,
                           std::vector<Affine<F>>& out) {
    const int w = n < 64 ? 4 : 8;
    const int nwin = (255 + w - 1) / w;
    std::vector<Jac<F>> tabj((size_t)nwin << w);
    Jac<F> base = Jac<F>::from_affine(g);
    for (int win = 0; win < nwin; win++) {
        Jac<F> cur = Jac<F>::inf();
        for (int k = 0; k < (1 << w); k++) {
            tabj[((size_t)win << w) + k] = cur;
            cur = Jac<F>::add(cur, base);
        }
        base = cur;   // 2^w * previous base
    }
    std::vector<Affine<F>> tab;
    batch_to_affine(tabj, tab);
    std::vector<Jac<F>> res(n);
#pragma omp parallel for num_threads(g_oracle_threads) if (g_oracle_threads > 1 && n >= 256) schedule(static)
    for (size_t i = 0; i < n; i++) {
        const uint64_t* s = scalars + 4 * i;
        Jac<F> acc = Jac<F>::inf();
        for (int win = 0; win < nwin; win++) {
            int bit = win * w, limb = bit / 64, off = bit % 64;
            uint64_t v = s[limb] >> off;
            if (off + w > 64 && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
            v &= (uint64_t(1) << w) - 1;
            if (v) acc = Jac<F>::add_mixed(acc, tab[((size_t)win << w) + v]);
        }
        res[i] = acc;
    }
    batch_to_affine(res, out);
}

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef Jac<Fq> G1Jac;
typedef Jac<Fq2> G2Jac;
