// oracle/ff.hpp -- CPU prime-field arithmetic for the parity oracle.  TEST INFRASTRUCTURE ONLY.
//
// PARITY UNPINNED (see oracle/README.md): restates the published arkworks `Fp256`/`Fp384`
// Montgomery representation (4 / 6 little-endian u64 limbs, R = 2^(64 N)); arkworks itself is an
// un-vendored, un-pinned git dependency of the reference (/root/reference/Cargo.toml:10-16) and is
// not on this machine.  Nothing under r1cs-spartan_b200/ includes this file.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdio>

typedef unsigned __int128 u128;

template <int N, int TAG>
struct Fp {
    uint64_t l[N];

    static uint64_t MOD[N];
    static uint64_t INV;      // -MOD^{-1} mod 2^64
    static Fp R1;             // R mod p   (Montgomery one)
    static Fp R2;             // R^2 mod p
    static Fp ZERO;

    static void init(const uint64_t* mod) {
        for (int i = 0; i < N; i++) MOD[i] = mod[i];
        uint64_t inv = 1;
        for (int i = 0; i < 63; i++) { inv *= inv; inv *= mod[0]; }   // mod[0]^(2^63-1) = mod[0]^-1
        INV = (uint64_t)0 - inv;
        memset(&ZERO, 0, sizeof(Fp));
        // R mod p by doubling 1 (64 N) times; R^2 by doubling (64 N) more times
        Fp x; memset(&x, 0, sizeof x); x.l[0] = 1;
        for (int i = 0; i < 64 * N; i++) x = dbl(x);
        R1 = x;
        for (int i = 0; i < 64 * N; i++) x = dbl(x);
        R2 = x;
    }
    static inline Fp zero() { return ZERO; }
    static inline Fp one() { return R1; }
    static inline bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > MOD[i]) return true;
            if (a[i] < MOD[i]) return false;
        }
        return true;
    }
    static inline void sub_mod(uint64_t* a) {
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)a[i] - MOD[i] - borrow;
            a[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    static inline Fp add(const Fp& a, const Fp& b) {
        Fp o; uint64_t carry = 0;
        for (int i = 0; i < N; i++) {
            u128 s = (u128)a.l[i] + b.l[i] + carry;
            o.l[i] = (uint64_t)s; carry = (uint64_t)(s >> 64);
        }
        if (carry || geq_mod(o.l)) sub_mod(o.l);
        return o;
    }
    static inline Fp dbl(const Fp& a) { return add(a, a); }
    static inline Fp sub(const Fp& a, const Fp& b) {
        Fp o; uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)a.l[i] - b.l[i] - borrow;
            o.l[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            uint64_t carry = 0;
            for (int i = 0; i < N; i++) {
                u128 s = (u128)o.l[i] + MOD[i] + carry;
                o.l[i] = (uint64_t)s; carry = (uint64_t)(s >> 64);
            }
        }
        return o;
    }
    static inline Fp neg(const Fp& a) { return a.is_zero() ? a : sub(ZERO, a); }
    // CIOS Montgomery product a*b*R^-1 mod p
    static inline Fp mul(const Fp& a, const Fp& b) {
        uint64_t t[N + 2];
        for (int i = 0; i < N + 2; i++) t[i] = 0;
        for (int i = 0; i < N; i++) {
            uint64_t carry = 0;
            for (int j = 0; j < N; j++) {
                u128 s = (u128)a.l[j] * b.l[i] + t[j] + carry;
                t[j] = (uint64_t)s; carry = (uint64_t)(s >> 64);
            }
            u128 s = (u128)t[N] + carry;
            t[N] = (uint64_t)s; t[N + 1] = (uint64_t)(s >> 64);
            uint64_t m = t[0] * INV;
            s = (u128)m * MOD[0] + t[0];
            carry = (uint64_t)(s >> 64);
            for (int j = 1; j < N; j++) {
                s = (u128)m * MOD[j] + t[j] + carry;
                t[j - 1] = (uint64_t)s; carry = (uint64_t)(s >> 64);
            }
            s = (u128)t[N] + carry;
            t[N - 1] = (uint64_t)s;
            t[N] = t[N + 1] + (uint64_t)(s >> 64);
        }
        Fp o;
        for (int i = 0; i < N; i++) o.l[i] = t[i];
        if (t[N] || geq_mod(o.l)) sub_mod(o.l);
        return o;
    }
    static inline Fp sqr(const Fp& a) { return mul(a, a); }
    inline bool is_zero() const {
        uint64_t x = 0;
        for (int i = 0; i < N; i++) x |= l[i];
        return x == 0;
    }
    inline bool operator==(const Fp& o) const { return memcmp(l, o.l, sizeof l) == 0; }
    inline bool operator!=(const Fp& o) const { return !(*this == o); }
    // canonical integer -> Montgomery, and back ("into_repr")
    static inline Fp from_canonical(const uint64_t* c) {
        Fp x; memcpy(x.l, c, sizeof x.l); return mul(x, R2);
    }
    inline void to_canonical(uint64_t* out) const {
        Fp one; memset(&one, 0, sizeof one); one.l[0] = 1;
        Fp c = mul(*this, one); memcpy(out, c.l, sizeof c.l);
    }
    static inline Fp from_u64(uint64_t v) {
        uint64_t c[N]; for (int i = 0; i < N; i++) c[i] = 0;
        c[0] = v; return from_canonical(c);
    }
    // a^e for a little-endian exponent of `n` limbs
    static Fp pow(const Fp& a, const uint64_t* e, int n) {
        Fp acc = R1;
        for (int i = n * 64 - 1; i >= 0; i--) {
            acc = sqr(acc);
            if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    static Fp inv(const Fp& a) {   // Fermat; a != 0
        uint64_t e[N]; memcpy(e, MOD, sizeof e);
        uint64_t borrow = 2;
        for (int i = 0; i < N && borrow; i++) {
            uint64_t old = e[i]; e[i] -= borrow; borrow = (old < borrow) ? 1 : 0;
        }
        return pow(a, e, N);
    }
    // compare canonical values: returns -1, 0, 1
    static int cmp_canonical(const Fp& a, const Fp& b) {
        uint64_t ca[N], cb[N]; a.to_canonical(ca); b.to_canonical(cb);
        for (int i = N - 1; i >= 0; i--) {
            if (ca[i] > cb[i]) return 1;
            if (ca[i] < cb[i]) return -1;
        }
        return 0;
    }
};
template <int N, int TAG> uint64_t Fp<N, TAG>::MOD[N];
template <int N, int TAG> uint64_t Fp<N, TAG>::INV;
template <int N, int TAG> Fp<N, TAG> Fp<N, TAG>::R1;
template <int N, int TAG> Fp<N, TAG> Fp<N, TAG>::R2;
template <int N, int TAG> Fp<N, TAG> Fp<N, TAG>::ZERO;

typedef Fp<4, 0> Fr;   // BLS12-381 scalar field
typedef Fp<6, 1> Fq;   // BLS12-381 base field

// Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
    Fq c0, c1;
    static inline Fq2 zero() { Fq2 z; z.c0 = Fq::ZERO; z.c1 = Fq::ZERO; return z; }
    static inline Fq2 one() { Fq2 z; z.c0 = Fq::R1; z.c1 = Fq::ZERO; return z; }
    static inline Fq2 add(const Fq2& a, const Fq2& b) { Fq2 o; o.c0 = Fq::add(a.c0, b.c0); o.c1 = Fq::add(a.c1, b.c1); return o; }
    static inline Fq2 sub(const Fq2& a, const Fq2& b) { Fq2 o; o.c0 = Fq::sub(a.c0, b.c0); o.c1 = Fq::sub(a.c1, b.c1); return o; }
    static inline Fq2 dbl(const Fq2& a) { return add(a, a); }
    static inline Fq2 neg(const Fq2& a) { Fq2 o; o.c0 = Fq::neg(a.c0); o.c1 = Fq::neg(a.c1); return o; }
    static inline Fq2 mul(const Fq2& a, const Fq2& b) {   // Karatsuba, u^2 = -1
        Fq v0 = Fq::mul(a.c0, b.c0), v1 = Fq::mul(a.c1, b.c1);
        Fq s = Fq::mul(Fq::add(a.c0, a.c1), Fq::add(b.c0, b.c1));
        Fq2 o; o.c0 = Fq::sub(v0, v1); o.c1 = Fq::sub(Fq::sub(s, v0), v1); return o;
    }
    static inline Fq2 sqr(const Fq2& a) {
        Fq s = Fq::add(a.c0, a.c1), d = Fq::sub(a.c0, a.c1), m = Fq::mul(a.c0, a.c1);
        Fq2 o; o.c0 = Fq::mul(s, d); o.c1 = Fq::dbl(m); return o;
    }
    static inline Fq2 inv(const Fq2& a) {
        Fq n = Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1));
        Fq ni = Fq::inv(n);
        Fq2 o; o.c0 = Fq::mul(a.c0, ni); o.c1 = Fq::neg(Fq::mul(a.c1, ni)); return o;
    }
    inline bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    inline bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    inline bool operator!=(const Fq2& o) const { return !(*this == o); }
    // UPSTREAM ark_ff QuadExtField Ord: c1 first, then c0
    static int cmp_canonical(const Fq2& a, const Fq2& b) {
        int c = Fq::cmp_canonical(a.c1, b.c1);
        return c ? c : Fq::cmp_canonical(a.c0, b.c0);
    }
};

static const uint64_t FR_MODULUS[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t FQ_MODULUS[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                        0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};

static inline void ff_init_all() {
    static bool done = false;
    if (done) return;
    Fr::init(FR_MODULUS); Fq::init(FQ_MODULUS);
    done = true;
}
