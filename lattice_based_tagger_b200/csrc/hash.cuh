// hash.cuh — string and key hashing shared by the host table builder and the device kernels.
//
// Every string the decode path compares (dictionary entries, lemma stems/eomis, feature words)
// is identified by a pair of 64-bit polynomial hashes over its UTF-16 code units plus, for the
// dictionary, its length.  Polynomial hashes compose: a sentence substring comes from two prefix
// hashes, a lemma `prefix + stem` or `eomi + suffix` from the parts without touching characters.
// A table key is the pair (k1, k2) = two independently mixed combinations of its components;
// k1 picks the slot, k2 is stored and compared, so a false match needs ~64 + log2(slots) equal
// bits.  The table builder rejects any two distinct keys with equal (k1, k2).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LT_HD __host__ __device__ __forceinline__
#else
#define LT_HD inline
#endif

namespace lt {

constexpr uint64_t kBaseA = 0x9E3779B97F4A7C15ull;   // odd
constexpr uint64_t kBaseB = 0xD6E8FEB86659FD93ull;   // odd

struct H2 {
    uint64_t a, b;
};

LT_HD uint64_t fmix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// h(s + c)
LT_HD H2 h2_push(H2 h, uint32_t c) {
    return H2{h.a * kBaseA + (uint64_t)(c + 1u), h.b * kBaseB + (uint64_t)(c + 1u)};
}

// h(x ++ y) from h(x), h(y) and base^|y|
LT_HD H2 h2_concat(H2 x, H2 y, H2 pow_len_y) {
    return H2{x.a * pow_len_y.a + y.a, x.b * pow_len_y.b + y.b};
}

// h(s[b:e]) from prefix hashes P[b], P[e] and base^(e-b)
LT_HD H2 h2_sub(H2 pre_b, H2 pre_e, H2 pow_len) {
    return H2{pre_e.a - pre_b.a * pow_len.a, pre_e.b - pre_b.b * pow_len.b};
}

// ---- dictionary key: (string, length) -------------------------------------------------------
LT_HD uint64_t dict_slot_hash(H2 h, uint32_t len) {
    return fmix64(h.a + (uint64_t)len * 0xA24BAED4963EE407ull);
}
LT_HD uint64_t dict_fp(H2 h, uint32_t len) {
    uint64_t f = fmix64(h.b ^ ((uint64_t)len * 0x9FB21C651E98DF25ull));
    return f ? f : 1;
}

// ---- rule key: up to three code units, exact -------------------------------------------------
LT_HD uint64_t rule_key(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t len) {
    return (uint64_t)c0 | ((uint64_t)c1 << 16) | ((uint64_t)c2 << 32) | ((uint64_t)len << 48);
}

// ---- feature key ------------------------------------------------------------------------------
// kind: 0..8 trigram templates; 16 morpheme preference; 17 word preference.  func: index of the
// owning scorer in the score program.  s0..s2: component strings (zero H2 when unused);
// a0, a1: small integers (tag ids, lengths, flags).
struct FKey {
    uint64_t k1, k2;
};

LT_HD FKey feature_key(uint32_t kind, uint32_t func, H2 s0, H2 s1, H2 s2, uint32_t a0, uint32_t a1) {
    uint64_t head = ((uint64_t)kind << 56) | ((uint64_t)func << 48) | ((uint64_t)a0 << 24) | (uint64_t)a1;
    uint64_t ka = head * 0x8CB92BA72F3D8DD7ull + s0.a * 0xE7037ED1A0B428DBull +
                  s1.a * 0x1D8E4E27C47D124Full + s2.a * 0xEB44ACCAB455D165ull;
    uint64_t kb = head * 0xA0761D6478BD642Full + s0.b * 0x2D358DCCAA6C78A5ull +
                  s1.b * 0x8BB84B93962EACC9ull + s2.b * 0x4B33A62ED433D4A3ull;
    FKey k;
    k.k1 = fmix64(ka);
    k.k2 = fmix64(kb ^ 0x589965CC75374CC3ull);
    if (k.k2 == 0) k.k2 = 1;
    return k;
}

constexpr uint32_t kKindMPref = 16;
constexpr uint32_t kKindWPref = 17;

}  // namespace lt
