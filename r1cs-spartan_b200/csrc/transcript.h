// transcript.h -- host side of the Fiat-Shamir transform: BLAKE2s-256 sponge-style RNG and the
// arkworks CanonicalSerialize byte encodings of everything that is fed to it or put in a proof.
//
// Mirrors how /root/reference/src/lib.rs:58-146 drives upstream `Blake2s512Rng`
// (`feed_randomness` = absorb canonical bytes; challenges via `Fr::rand(&mut fs_rng)`), restated
// from the published upstream algorithm (linear_sumcheck::data_structures::random, ark_ff
// `UniformRand for Fp256`, ark_serialize flags); upstream sources are not available here, see
// DESIGN.md "Parity status".
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "field.cuh"
#include "ec.cuh"

namespace sbhost {

// ---------------------------------------------------------------- BLAKE2s (RFC 7693), unkeyed, 32-byte digest
class Blake2s256 {
  public:
    Blake2s256() { reset(); }
    void reset() {
        for (int i = 0; i < 8; i++) h_[i] = kIV[i];
        h_[0] ^= 0x01010000u ^ 32u;
        count_ = 0; fill_ = 0;
    }
    void absorb(const void* data, size_t len) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        while (len) {
            if (fill_ == 64) { count_ += 64; round_block(block_, false); fill_ = 0; }
            // whole blocks straight from the input (the last block always stays buffered: it may be the final one)
            while (fill_ == 0 && len > 64) { count_ += 64; round_block(p, false); p += 64; len -= 64; }
            size_t take = 64 - fill_; if (take > len) take = len;
            memcpy(block_ + fill_, p, take);
            fill_ += take; p += take; len -= take;
        }
    }
    // digest of everything absorbed so far; the object itself is left untouched
    void peek_digest(uint8_t out[32]) const {
        Blake2s256 t = *this;
        t.count_ += t.fill_;
        memset(t.block_ + t.fill_, 0, 64 - t.fill_);
        t.round_block(t.block_, true);
        for (int i = 0; i < 32; i++) out[i] = (uint8_t)(t.h_[i >> 2] >> (8 * (i & 3)));
    }

  private:
    static constexpr uint32_t kIV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                        0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    static inline uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    static inline __attribute__((always_inline)) void mix(uint32_t* v, int a, int b, int c, int d, uint32_t x, uint32_t y) {
        v[a] += v[b] + x; v[d] = ror(v[d] ^ v[a], 16);
        v[c] += v[d];     v[b] = ror(v[b] ^ v[c], 12);
        v[a] += v[b] + y; v[d] = ror(v[d] ^ v[a], 8);
        v[c] += v[d];     v[b] = ror(v[b] ^ v[c], 7);
    }
    // one round with its message schedule as compile-time constants (the indices fold into register names)
#define SB_B2_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15)       \
    mix(v, 0, 4, 8, 12, m[s0], m[s1]);   mix(v, 1, 5, 9, 13, m[s2], m[s3]);                       \
    mix(v, 2, 6, 10, 14, m[s4], m[s5]);  mix(v, 3, 7, 11, 15, m[s6], m[s7]);                      \
    mix(v, 0, 5, 10, 15, m[s8], m[s9]);  mix(v, 1, 6, 11, 12, m[s10], m[s11]);                    \
    mix(v, 2, 7, 8, 13, m[s12], m[s13]); mix(v, 3, 4, 9, 14, m[s14], m[s15]);
    void round_block(const uint8_t* blk, bool final_block) {
        uint32_t m[16], v[16];
        memcpy(m, blk, 64);                                               // little-endian host
        for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[8 + i] = kIV[i]; }
        v[12] ^= (uint32_t)count_; v[13] ^= (uint32_t)(count_ >> 32);
        if (final_block) v[14] = ~v[14];
        SB_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
        SB_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
        SB_B2_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
        SB_B2_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
        SB_B2_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
        SB_B2_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
        SB_B2_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
        SB_B2_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
        SB_B2_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
        SB_B2_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[8 + i];
    }
#undef SB_B2_ROUND
    uint32_t h_[8];
    uint64_t count_;
    uint8_t block_[64];
    size_t fill_;
};

// ---------------------------------------------------------------- the feedable RNG
class Transcript {
  public:
    void feed(const std::vector<uint8_t>& bytes) { hash_.absorb(bytes.data(), bytes.size()); }
    void feed(const void* p, size_t n) { hash_.absorb(p, n); }
    void fill(uint8_t* dest, size_t n) {
        uint8_t out[32];
        hash_.peek_digest(out);
        size_t used = 0;
        for (size_t i = 0; i < n; i++) {
            dest[i] = out[used++];
            if (used == 32) { hash_.absorb(out, 32); hash_.peek_digest(out); used = 0; }
        }
        hash_.absorb(out, 32);
    }
    uint64_t next_u64() { uint64_t v; fill(reinterpret_cast<uint8_t*>(&v), 8); return v; }
    // Fr::rand: four u64 draws fill the limbs, the top bit is shaved, retry until below the modulus;
    // the accepted limbs are the Montgomery residue as is.
    Fr challenge() {
        for (;;) {
            Fr x;
            for (int i = 0; i < 4; i++) { uint64_t u = next_u64(); x.l[2 * i] = (uint32_t)u; x.l[2 * i + 1] = (uint32_t)(u >> 32); }
            x.l[7] &= 0x7fffffffu;
            if (!Fr::geq_mod(x.l)) return x;
        }
    }
  private:
    Blake2s256 hash_;
};

// ---------------------------------------------------------------- CanonicalSerialize encodings
typedef std::vector<uint8_t> Bytes;
inline void put_u64(Bytes& o, uint64_t v) { for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i))); }
template <class P>
inline void put_fp_canonical(Bytes& o, const Fp<P>& x) {
    Fp<P> c = x.to_canonical();
    for (int i = 0; i < P::N; i++) for (int b = 0; b < 4; b++) o.push_back((uint8_t)(c.l[i] >> (8 * b)));
}
inline void put_fr(Bytes& o, const Fr& x) { put_fp_canonical(o, x); }
inline void put_fr_vec(Bytes& o, const Fr* v, size_t n) { put_u64(o, n); for (size_t i = 0; i < n; i++) put_fr(o, v[i]); }
// compressed points: x (little-endian canonical), flags in the top two bits of the last byte:
// bit 7 = y is the larger of {y, -y}, bit 6 = infinity
inline void put_g1(Bytes& o, const G1Aff& p) {
    if (p.is_inf()) { o.insert(o.end(), 48, 0); o.back() |= 0x40; return; }
    put_fp_canonical(o, p.x);
    if (Fq::cmp_canonical_host(p.y, Fq::neg(p.y)) > 0) o.back() |= 0x80;
}
inline void put_g2(Bytes& o, const G2Aff& p) {
    if (p.is_inf()) { o.insert(o.end(), 96, 0); o.back() |= 0x40; return; }
    put_fp_canonical(o, p.x.c0); put_fp_canonical(o, p.x.c1);
    if (Fq2::cmp_canonical_host(p.y, Fq2::neg(p.y)) > 0) o.back() |= 0x80;
}

}  // namespace sbhost
