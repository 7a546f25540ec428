// field.cuh -- BLS12-381 Fr / Fq / Fq2 in 32-bit limbs, Montgomery form (R = 2^(32 N)).
//
// Memory image == arkworks `Fp256` / `Fp384` (4 / 6 little-endian u64 Montgomery limbs), so Rust
// slices of `E::Fr` cross the C ABI zero-copy (reference: src/commitment/commit.rs:20-21 `into_repr`
// is the only place the reference leaves Montgomery form).
//
// Device code uses the generated PTX carry chains (fp_gen.cuh: mad.lo.cc/madc.hi.cc pairs that ptxas
// fuses into IMAD.WIDE.U32 with carry predicates); host code (transcript-side scalar work,
// serialization, final affine conversion) uses the portable 64-bit-accumulator path below.
#pragma once
#include <cstdint>
#include <cstring>
#include "fp_gen.cuh"

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#define SB_D __device__ __forceinline__
#else
#define SB_HD inline
#define SB_D inline
#endif

struct FrParams {
    static constexpr int N = FR_LIMBS;
    static constexpr uint32_t INV = FR_INV32;
    SB_HD static constexpr uint32_t mod(int i) { constexpr uint32_t m[N] = FR_MOD_INIT; return m[i]; }
    SB_HD static constexpr uint32_t r1(int i) { constexpr uint32_t m[N] = FR_R1_INIT; return m[i]; }
    SB_HD static constexpr uint32_t r2(int i) { constexpr uint32_t m[N] = FR_R2_INIT; return m[i]; }
#if defined(__CUDACC__)
    SB_D static void mul_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fr_mul_ptx(r, a, b); }
    SB_D static void add_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fr_add_ptx(r, a, b); }
    SB_D static void sub_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fr_sub_ptx(r, a, b); }
#endif
};
struct FqParams {
    static constexpr int N = FQ_LIMBS;
    static constexpr uint32_t INV = FQ_INV32;
    SB_HD static constexpr uint32_t mod(int i) { constexpr uint32_t m[N] = FQ_MOD_INIT; return m[i]; }
    SB_HD static constexpr uint32_t r1(int i) { constexpr uint32_t m[N] = FQ_R1_INIT; return m[i]; }
    SB_HD static constexpr uint32_t r2(int i) { constexpr uint32_t m[N] = FQ_R2_INIT; return m[i]; }
#if defined(__CUDACC__)
    SB_D static void mul_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fq_mul_ptx(r, a, b); }
    SB_D static void add_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fq_add_ptx(r, a, b); }
    SB_D static void sub_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fq_sub_ptx(r, a, b); }
#endif
};

template <class P>
struct alignas(16) Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    SB_HD static Fp zero() { Fp z; for (int i = 0; i < N; i++) z.l[i] = 0; return z; }
    SB_HD static Fp one() { Fp z; for (int i = 0; i < N; i++) z.l[i] = P::r1(i); return z; }
    SB_HD static Fp rr() { Fp z; for (int i = 0; i < N; i++) z.l[i] = P::r2(i); return z; }
    SB_HD bool is_zero() const { uint32_t x = 0; for (int i = 0; i < N; i++) x |= l[i]; return x == 0; }
    SB_HD bool operator==(const Fp& o) const { uint32_t x = 0; for (int i = 0; i < N; i++) x |= l[i] ^ o.l[i]; return x == 0; }
    SB_HD bool operator!=(const Fp& o) const { return !(*this == o); }

    // ---- portable paths (host; also usable on device for cross-checking the PTX path)
    SB_HD static bool geq_mod(const uint32_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > P::mod(i)) return true;
            if (a[i] < P::mod(i)) return false;
        }
        return true;
    }
    SB_HD static void sub_mod(uint32_t* a) {
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            uint64_t d = (uint64_t)a[i] - P::mod(i) - borrow;
            a[i] = (uint32_t)d; borrow = (d >> 32) & 1;
        }
    }
    SB_HD static Fp add_portable(const Fp& a, const Fp& b) {
        Fp o; uint64_t c = 0;
        for (int i = 0; i < N; i++) { c += (uint64_t)a.l[i] + b.l[i]; o.l[i] = (uint32_t)c; c >>= 32; }
        if (c || geq_mod(o.l)) sub_mod(o.l);
        return o;
    }
    SB_HD static Fp sub_portable(const Fp& a, const Fp& b) {
        Fp o; uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            uint64_t d = (uint64_t)a.l[i] - b.l[i] - borrow;
            o.l[i] = (uint32_t)d; borrow = (d >> 32) & 1;
        }
        if (borrow) {
            uint64_t c = 0;
            for (int i = 0; i < N; i++) { c += (uint64_t)o.l[i] + P::mod(i); o.l[i] = (uint32_t)c; c >>= 32; }
        }
        return o;
    }
    SB_HD static Fp mul_portable(const Fp& a, const Fp& b) {
        uint32_t t[N + 2];
        for (int i = 0; i < N + 2; i++) t[i] = 0;
        for (int i = 0; i < N; i++) {
            uint64_t c = 0;
            for (int j = 0; j < N; j++) { c += (uint64_t)a.l[j] * b.l[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
            c += t[N]; t[N] = (uint32_t)c; t[N + 1] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * P::INV;
            c = (uint64_t)m * P::mod(0) + t[0]; c >>= 32;
            for (int j = 1; j < N; j++) { c += (uint64_t)m * P::mod(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
            c += t[N]; t[N - 1] = (uint32_t)c; t[N] = t[N + 1] + (uint32_t)(c >> 32);
        }
        Fp o;
        for (int i = 0; i < N; i++) o.l[i] = t[i];
        if (t[N] || geq_mod(o.l)) sub_mod(o.l);
        return o;
    }

    // ---- host fast path: the same Montgomery product (R = 2^(32 N) = 2^(64 N/2)) on 64-bit limbs with 128-bit products,
    // ~8x faster than the 32-bit loop above.  The host sums the shards' partial group elements and converts every proof
    // element to affine coordinates with these (round 2: at 8 GPUs 119 G2 additions per opening on the 32-bit path were
    // 2 ms of a 5 ms phase).  Bit-identical by construction (the result is the unique residue below p); checked against
    // mul_portable by sb_selftest_host_field.
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__)
    static inline Fp mul_host64(const Fp& a, const Fp& b) {
        constexpr int M = N / 2;
        static_assert(N % 2 == 0, "limb count must be even");
        typedef unsigned __int128 u128;
        uint64_t A[M], B[M], Pm[M], t[M + 2];
        for (int i = 0; i < M; i++) {
            A[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
            B[i] = (uint64_t)b.l[2 * i] | ((uint64_t)b.l[2 * i + 1] << 32);
            Pm[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
        }
        // -p^-1 mod 2^64 from the 32-bit constant: y = p^-1 mod 2^32, one Newton step doubles the precision
        const uint64_t y32 = (uint64_t)(uint32_t)(0u - P::INV);
        const uint64_t inv64 = 0 - (y32 * (2 - Pm[0] * y32));
        for (int i = 0; i < M + 2; i++) t[i] = 0;
        for (int i = 0; i < M; i++) {
            u128 c = 0;
            for (int j = 0; j < M; j++) { c += (u128)A[j] * B[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[M]; t[M] = (uint64_t)c; t[M + 1] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * inv64;
            c = (u128)m * Pm[0] + t[0]; c >>= 64;
            for (int j = 1; j < M; j++) { c += (u128)m * Pm[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[M]; t[M - 1] = (uint64_t)c; t[M] = t[M + 1] + (uint64_t)(c >> 64);
        }
        Fp o;
        for (int i = 0; i < M; i++) { o.l[2 * i] = (uint32_t)t[i]; o.l[2 * i + 1] = (uint32_t)(t[i] >> 32); }
        if (t[M] || geq_mod(o.l)) sub_mod(o.l);
        return o;
    }
#endif

    // ---- dispatch
    SB_HD static Fp add(const Fp& a, const Fp& b) {
#if defined(__CUDA_ARCH__)
        Fp o; P::add_ptx(o.l, a.l, b.l); return o;
#else
        return add_portable(a, b);
#endif
    }
    SB_HD static Fp sub(const Fp& a, const Fp& b) {
#if defined(__CUDA_ARCH__)
        Fp o; P::sub_ptx(o.l, a.l, b.l); return o;
#else
        return sub_portable(a, b);
#endif
    }
    SB_HD static Fp mul(const Fp& a, const Fp& b) {
#if defined(__CUDA_ARCH__)
        Fp o; P::mul_ptx(o.l, a.l, b.l); return o;
#elif defined(__SIZEOF_INT128__)
        return mul_host64(a, b);
#else
        return mul_portable(a, b);
#endif
    }
    SB_HD static Fp sqr(const Fp& a) { return mul(a, a); }
    SB_HD static Fp dbl(const Fp& a) { return add(a, a); }
    SB_HD static Fp neg(const Fp& a) { return sub(zero(), a); }

    // Montgomery <-> canonical ("into_repr")
    SB_HD static Fp from_canonical(const Fp& c) { return mul(c, rr()); }
    SB_HD Fp to_canonical() const { Fp o = zero(); o.l[0] = 1; return mul(*this, o); }
    SB_HD static Fp from_u32(uint32_t v) { Fp o = zero(); o.l[0] = v; return from_canonical(o); }

    // host helpers
    static Fp pow_host(const Fp& a, const uint32_t* e, int nlimbs) {
        Fp acc = one();
        for (int i = nlimbs * 32 - 1; i >= 0; i--) {
            acc = sqr(acc);
            if ((e[i / 32] >> (i % 32)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    static Fp inv_host(const Fp& a) { return inv(a); }
    // Fermat inversion a^(p-2), a != 0 (also used on the device by the batch-affine kernels)
#if defined(__CUDACC__)
    __host__ __device__ __noinline__
#endif
    static Fp inv(const Fp& a) {
        uint32_t e[N], borrow = 2;                       // e = p - 2
        for (int i = 0; i < N; i++) { uint32_t m = P::mod(i); e[i] = m - borrow; borrow = (m < borrow) ? 1u : 0u; }
        Fp acc = one();
#pragma unroll 1
        for (int i = N * 32 - 1; i >= 0; i--) {
            acc = sqr(acc);
            if ((e[i / 32] >> (i % 32)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    // ---- fast inversion: binary extended GCD with the updates delayed and applied 31 steps at a time
    // (T. Pornin, "Optimized Binary GCD for Modular Inversion", 2020, Algorithm 2, k = 32, 32-bit limbs).
    // 25 outer rounds for a 381-bit modulus, each: 31 cheap steps on 64-bit approximations of (a, b) that build a
    // 2 x 2 transition matrix (f0 g0; f1 g1), then ONE pass of 12-limb multiply-accumulates that applies it to (a, b)
    // exactly and to (u, v) modulo p -- about 30 k instructions against the 170 k of the Fermat ladder above, and no
    // Montgomery products at all.  Not constant time (the values inverted here are public).  The result is the exact
    // modular inverse, so it is bit-identical to inv() for every input (checked on the host by tests/test_cpu_oracle.py
    // through sb_selftest_host_field, and on the device by the GPU parity tests of the batched-affine MSM rounds).
    // a != 0 (mod p); returns the Montgomery form of the inverse of the Montgomery-form input.
#if defined(__CUDACC__)
    __host__ __device__ __noinline__
#endif
    static Fp inv_fast(const Fp& x) {
        constexpr int K = 31;                                   // inner steps per round = k - 1
        const uint32_t minv = P::INV & 0x7fffffffu;            // -p^-1 mod 2^31
        uint32_t a[N], b[N], u[N], v[N];
        for (int i = 0; i < N; i++) { a[i] = x.l[i]; b[i] = P::mod(i); u[i] = 0; v[i] = 0; }
        u[0] = 1;
        // invariants: a = u * x, b = v * x (mod p), as plain integers (x is whatever residue the caller holds)
        for (int round = 0; round < (2 * 32 * N - 1 + K - 1) / K; round++) {
            // 64-bit approximations: the low 31 bits and the top 33 bits of the n-bit window, n = max(len a, len b, 64).
            // The three limbs at the top of the longer value are picked with a fixed, fully unrolled scan (no dynamic
            // indexing: the limb arrays stay in registers) and every step below is branch-free (the lanes of a warp
            // invert different values: data-dependent branches would serialise them).
            uint32_t ah = a[N - 1], am = a[N - 2], al = a[N - 3], bh = b[N - 1], bm = b[N - 2], bl = b[N - 3];
#pragma unroll
            for (int i = N - 2; i >= 2; i--) {
                const bool empty = (ah | bh) == 0;                       // nothing found above limb i yet: slide the window down
                ah = empty ? a[i] : ah; am = empty ? a[i - 1] : am; al = empty ? a[i - 2] : al;
                bh = empty ? b[i] : bh; bm = empty ? b[i - 1] : bm; bl = empty ? b[i - 2] : bl;
            }
            uint64_t abar, bbar;
            {
                const uint32_t hi = ah | bh;
                const bool exact = hi == 0;                             // both values below 2^64: (am, al) = limbs (1, 0) are the values themselves
                int lz = 0;
#if defined(__CUDA_ARCH__)
                lz = exact ? 0 : __clz((int)hi);
#else
                if (!exact) lz = __builtin_clz(hi);
#endif
                // top 64 bits of the window (ah, am, al) << lz, then keep its upper 33 bits
                uint64_t wa = ((uint64_t)ah << 32) | am, wb = ((uint64_t)bh << 32) | bm;
                wa = lz ? (wa << lz) | ((uint64_t)al >> (32 - lz)) : wa;
                wb = lz ? (wb << lz) | ((uint64_t)bl >> (32 - lz)) : wb;
                const uint64_t xa = ((wa >> 31) << 31) | (a[0] & 0x7fffffffu), xb = ((wb >> 31) << 31) | (b[0] & 0x7fffffffu);
                const uint64_t ea = ((uint64_t)am << 32) | al, eb = ((uint64_t)bm << 32) | bl;
                abar = exact ? ea : xa; bbar = exact ? eb : xb;
            }
            int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
            for (int j = 0; j < K; j++) {
                const uint64_t odd = (uint64_t)0 - (abar & 1);                          // all ones when a is odd
                const uint64_t sw = odd & ((uint64_t)0 - (uint64_t)(abar < bbar));       // ... and smaller than b: swap first
                uint64_t x = (abar ^ bbar) & sw; abar ^= x; bbar ^= x;
                x = (uint64_t)(f0 ^ f1) & sw; f0 ^= (int64_t)x; f1 ^= (int64_t)x;
                x = (uint64_t)(g0 ^ g1) & sw; g0 ^= (int64_t)x; g1 ^= (int64_t)x;
                abar -= bbar & odd; f0 -= f1 & (int64_t)odd; g0 -= g1 & (int64_t)odd;
                abar >>= 1; f1 <<= 1; g1 <<= 1;
            }
            // (a, b) <- ((a f0 + b g0) / 2^31, (a f1 + b g1) / 2^31), exact; a negative result is negated together with its row
            uint32_t na[N], nb[N];
            if (lincomb_shift(na, a, f0, b, g0, false, minv)) { f0 = -f0; g0 = -g0; }
            if (lincomb_shift(nb, a, f1, b, g1, false, minv)) { f1 = -f1; g1 = -g1; }
            // (u, v) <- the same combinations modulo p, with the division by 2^31 done Montgomery-style
            uint32_t nu[N], nv[N];
            lincomb_shift(nu, u, f0, v, g0, true, minv);
            lincomb_shift(nv, u, f1, v, g1, true, minv);
            for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
        }
        // b = gcd = 1 and v = x^-1 as a plain integer residue: x = X R  ->  v = X^-1 R^-1; two products by R^2 give X^-1 R
        Fp r; for (int i = 0; i < N; i++) r.l[i] = v[i];
        return mul(mul(r, rr()), rr());
    }
    // out <- (x f + y g) / 2^31 for |f| + |g| <= 2^31.  modular = false: the division is exact, the result may be negative:
    // its absolute value is stored and `true` returned in that case.  modular = true: x, y in [0, p); a multiple of p is
    // added first so that the low 31 bits vanish, and the result is normalised into [0, p); returns false.
    SB_HD static bool lincomb_shift(uint32_t* out, const uint32_t* x, int64_t f, const uint32_t* y, int64_t g, bool modular, uint32_t minv) {
        const bool fneg = f < 0, gneg = g < 0;
        const uint64_t fa = (uint64_t)(fneg ? -f : f), ga = (uint64_t)(gneg ? -g : g);      // <= 2^31
        // t = x fa (sign fneg) + y ga (sign gneg) in two's complement over N + 2 limbs
        uint32_t t[N + 2];
        {
            uint64_t c1 = 0, c2 = 0;
            uint32_t p1[N + 1], p2[N + 1];
            for (int i = 0; i < N; i++) {
                c1 += (uint64_t)x[i] * fa; p1[i] = (uint32_t)c1; c1 >>= 32;
                c2 += (uint64_t)y[i] * ga; p2[i] = (uint32_t)c2; c2 >>= 32;
            }
            p1[N] = (uint32_t)c1; p2[N] = (uint32_t)c2;
            // add / subtract
            uint64_t carry = (fneg ? 1 : 0) + (uint64_t)(gneg ? 1 : 0);     // +1 of each two's complement negation
            for (int i = 0; i <= N; i++) {
                const uint32_t w1 = fneg ? ~p1[i] : p1[i], w2 = gneg ? ~p2[i] : p2[i];
                carry += (uint64_t)w1 + w2; t[i] = (uint32_t)carry; carry >>= 32;
            }
            const uint32_t e1 = fneg ? 0xffffffffu : 0, e2 = gneg ? 0xffffffffu : 0;   // sign extensions
            carry += (uint64_t)e1 + e2; t[N + 1] = (uint32_t)carry;
        }
        if (modular) {
            // t += q p with q = (t mod 2^31) * (-p^-1) mod 2^31
            const uint32_t q = (t[0] * minv) & 0x7fffffffu;
            uint64_t c = 0;
            for (int i = 0; i < N; i++) { c += (uint64_t)q * P::mod(i) + t[i]; t[i] = (uint32_t)c; c >>= 32; }
            c += t[N]; t[N] = (uint32_t)c; c >>= 32;
            t[N + 1] += (uint32_t)c;
        }
        // arithmetic shift right by 31
        const bool neg = (t[N + 1] & 0x80000000u) != 0;
        uint32_t r[N + 1];
        for (int i = 0; i <= N; i++) r[i] = (t[i] >> 31) | (t[i + 1] << 1);
        if (!modular) {
            if (neg) {      // store the absolute value
                uint64_t c = 1;
                for (int i = 0; i < N; i++) { c += (uint64_t)(~r[i]); out[i] = (uint32_t)c; c >>= 32; }
            } else {
                for (int i = 0; i < N; i++) out[i] = r[i];
            }
            return neg;
        }
        // modular: r in (-p, 2p); bring into [0, p)
        if (neg) {
            uint64_t c = 0;
            for (int i = 0; i < N; i++) { c += (uint64_t)r[i] + P::mod(i); out[i] = (uint32_t)c; c >>= 32; }
        } else {
            for (int i = 0; i < N; i++) out[i] = r[i];
            if (r[N] || geq_mod(out)) sub_mod(out);
        }
        return false;
    }
    static int cmp_canonical_host(const Fp& a, const Fp& b) {
        Fp ca = a.to_canonical(), cb = b.to_canonical();
        for (int i = N - 1; i >= 0; i--) {
            if (ca.l[i] > cb.l[i]) return 1;
            if (ca.l[i] < cb.l[i]) return -1;
        }
        return 0;
    }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// Fq2 = Fq[u] / (u^2 + 1).  mul/sqr are real (non-inlined) device functions: one G2 point addition is
// ~14 of them, and with everything inlined its ~190 KB of straight-line code streams through the
// instruction cache once per addition (ncu: `no_instruction` was the second largest stall reason).
#if defined(__CUDACC__)
#define SB_FQ2_FN __host__ __device__ __noinline__
#else
#define SB_FQ2_FN
#endif
struct alignas(16) Fq2 {
    Fq c0, c1;
    SB_HD static Fq2 zero() { Fq2 z; z.c0 = Fq::zero(); z.c1 = Fq::zero(); return z; }
    SB_HD static Fq2 one() { Fq2 z; z.c0 = Fq::one(); z.c1 = Fq::zero(); return z; }
    SB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    SB_HD bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    SB_HD bool operator!=(const Fq2& o) const { return !(*this == o); }
    SB_HD static Fq2 add(const Fq2& a, const Fq2& b) { Fq2 o; o.c0 = Fq::add(a.c0, b.c0); o.c1 = Fq::add(a.c1, b.c1); return o; }
    SB_HD static Fq2 sub(const Fq2& a, const Fq2& b) { Fq2 o; o.c0 = Fq::sub(a.c0, b.c0); o.c1 = Fq::sub(a.c1, b.c1); return o; }
    SB_HD static Fq2 dbl(const Fq2& a) { return add(a, a); }
    SB_HD static Fq2 neg(const Fq2& a) { Fq2 o; o.c0 = Fq::neg(a.c0); o.c1 = Fq::neg(a.c1); return o; }
    SB_FQ2_FN static Fq2 mul(const Fq2& a, const Fq2& b) {
#if defined(__CUDA_ARCH__)
        // Karatsuba with lazy reduction: three 768-bit products, combined before reducing, two Montgomery
        // reductions instead of three (720 instead of 864 IMAD.WIDE).  Bounds: every product of operands < 2p is
        // < 4p^2 < 2^764; c1 = a0 b1 + a1 b0 < 2p^2 and c0 = a0 b0 - a1 b1 + p^2 in (0, 2p^2), both < p 2^384.
        // (the operands arrive by reference -- this function is outlined -- and usually sit in local or shared memory:
        // fetch each with six 128-bit loads instead of 24 scalar ones)
        uint32_t al[24], bl[24];
        {
            const uint4* pa = reinterpret_cast<const uint4*>(&a); const uint4* pb = reinterpret_cast<const uint4*>(&b);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const uint4 x = pa[i], y = pb[i];
                al[4 * i] = x.x; al[4 * i + 1] = x.y; al[4 * i + 2] = x.z; al[4 * i + 3] = x.w;
                bl[4 * i] = y.x; bl[4 * i + 1] = y.y; bl[4 * i + 2] = y.z; bl[4 * i + 3] = y.w;
            }
        }
        uint32_t t0[24], t1[24], t2[24], u[24], sa[12], sb[12];
        fq_mul_wide_ptx(t0, al, bl);
        fq_mul_wide_ptx(t1, al + 12, bl + 12);
        fq_wide_addn_ptx(sa, al, al + 12);
        fq_wide_addn_ptx(sb, bl, bl + 12);
        fq_mul_wide_ptx(t2, sa, sb);
        fq_wide_sub_ptx(u, t2, t0);
        fq_wide_sub_ptx(t2, u, t1);
        fq_wide_subp2_ptx(u, t0, t1);
        Fq2 o;
        fq_redc_ptx(o.c0.l, u);
        fq_redc_ptx(o.c1.l, t2);
        return o;
#else
        Fq v0 = Fq::mul(a.c0, b.c0), v1 = Fq::mul(a.c1, b.c1);
        Fq s = Fq::mul(Fq::add(a.c0, a.c1), Fq::add(b.c0, b.c1));
        Fq2 o; o.c0 = Fq::sub(v0, v1); o.c1 = Fq::sub(Fq::sub(s, v0), v1); return o;
#endif
    }
    SB_FQ2_FN static Fq2 sqr(const Fq2& a_) {
#if defined(__CUDA_ARCH__)
        Fq2 a;                       // six 128-bit loads of the by-reference operand (see mul)
        {
            const uint4* pa = reinterpret_cast<const uint4*>(&a_);
            uint4* da = reinterpret_cast<uint4*>(&a);
#pragma unroll
            for (int i = 0; i < 6; i++) da[i] = pa[i];
        }
#else
        const Fq2& a = a_;
#endif
        Fq s = Fq::add(a.c0, a.c1), d = Fq::sub(a.c0, a.c1), m = Fq::mul(a.c0, a.c1);
        Fq2 o; o.c0 = Fq::mul(s, d); o.c1 = Fq::dbl(m); return o;
    }
    static Fq2 inv_host(const Fq2& a) { return inv(a); }
#if defined(__CUDACC__)
    __host__ __device__
#endif
    static Fq2 inv(const Fq2& a) {
        Fq n = Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1));
        Fq ni = Fq::inv(n);
        Fq2 o; o.c0 = Fq::mul(a.c0, ni); o.c1 = Fq::neg(Fq::mul(a.c1, ni)); return o;
    }
#if defined(__CUDACC__)
    __host__ __device__
#endif
    static Fq2 inv_fast(const Fq2& a) {
        Fq n = Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1));
        Fq ni = Fq::inv_fast(n);
        Fq2 o; o.c0 = Fq::mul(a.c0, ni); o.c1 = Fq::neg(Fq::mul(a.c1, ni)); return o;
    }
    // arkworks QuadExtField ordering: c1 first, then c0
    static int cmp_canonical_host(const Fq2& a, const Fq2& b) {
        int c = Fq::cmp_canonical_host(a.c1, b.c1);
        return c ? c : Fq::cmp_canonical_host(a.c0, b.c0);
    }
};
