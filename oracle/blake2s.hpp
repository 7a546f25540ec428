// oracle/blake2s.hpp -- BLAKE2s-256 (RFC 7693, unkeyed) + the Fiat-Shamir RNG built on it.
// TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see oracle/README.md).
//
// UPSTREAM linear_sumcheck::data_structures::random::Blake2s512Rng (despite the name it wraps
// Blake2s with a 32-byte digest) as used at /root/reference/src/lib.rs:61-65,74-135:
//   feed_randomness(m): hasher.update(canonical_serialize(m))
//   fill_bytes(dest):   out = hasher.clone().finalize(); copy bytes; every 32 bytes consumed:
//                       hasher.update(out), recompute out; after filling: hasher.update(out).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

struct Blake2s {
    uint32_t h[8];
    uint64_t t;
    uint8_t buf[64];
    size_t buflen;

    static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void init() {
        static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        for (int i = 0; i < 8; i++) h[i] = IV[i];
        h[0] ^= 0x01010020u;
        t = 0; buflen = 0;
    }
    void compress(const uint8_t* block, bool last) {
        static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                       0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        static const uint8_t S[10][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
            {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4},
            {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13},
            {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11},
            {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5},
            {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; i++) {
            m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) |
                   ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
        }
        for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
        v[12] ^= (uint32_t)t; v[13] ^= (uint32_t)(t >> 32);
        if (last) v[14] = ~v[14];
#define B2S_G(a, b, c, d, x, y)                                  \
        v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 16);  \
        v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 12);  \
        v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 8);   \
        v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 7);
        for (int r = 0; r < 10; r++) {
            const uint8_t* s = S[r];
            B2S_G(0, 4, 8, 12, m[s[0]], m[s[1]]);
            B2S_G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            B2S_G(2, 6, 10, 14, m[s[4]], m[s[5]]);
            B2S_G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            B2S_G(0, 5, 10, 15, m[s[8]], m[s[9]]);
            B2S_G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            B2S_G(2, 7, 8, 13, m[s[12]], m[s[13]]);
            B2S_G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
#undef B2S_G
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const uint8_t* in, size_t len) {
        while (len > 0) {
            if (buflen == 64) {            // buffer full and more input follows: not the last block
                t += 64; compress(buf, false); buflen = 0;
            }
            size_t take = 64 - buflen; if (take > len) take = len;
            memcpy(buf + buflen, in, take);
            buflen += take; in += take; len -= take;
        }
    }
    void finalize(uint8_t out[32]) const {   // const: works on a copy ("hasher.clone().finalize()")
        Blake2s c = *this;
        c.t += c.buflen;
        memset(c.buf + c.buflen, 0, 64 - c.buflen);
        c.compress(c.buf, true);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)c.h[i]; out[4 * i + 1] = (uint8_t)(c.h[i] >> 8);
            out[4 * i + 2] = (uint8_t)(c.h[i] >> 16); out[4 * i + 3] = (uint8_t)(c.h[i] >> 24);
        }
    }
};

struct FsRng {
    Blake2s hasher;
    void setup() { hasher.init(); }
    void feed(const uint8_t* data, size_t len) { hasher.update(data, len); }
    void fill_bytes(uint8_t* dest, size_t n) {
        uint8_t out[32];
        hasher.finalize(out);
        size_t dp = 0;
        for (size_t p = 0; p < n; p++) {
            dest[p] = out[dp++];
            if (dp == 32) { hasher.update(out, 32); hasher.finalize(out); dp = 0; }
        }
        hasher.update(out, 32);
    }
    uint64_t next_u64() {
        uint8_t b[8]; fill_bytes(b, 8);
        uint64_t v = 0;
        for (int i = 7; i >= 0; i--) v = (v << 8) | b[i];
        return v;
    }
};
