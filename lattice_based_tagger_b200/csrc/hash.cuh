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
// slot hash: hash a combined with the length; fingerprint: hash b combined with the length
// (compared in full, so it needs no mixing).
LT_HD uint64_t dict_slot_hash(H2 h, uint32_t len) { return h.a + (uint64_t)len * 0xA24BAED4963EE407ull; }

// ---- cuckoo placement --------------------------------------------------------------------------
// The dictionary, the rule keys and the feature table are cuckoo tables: a key lives in one of
// TWO slots, both derived from the same 64-bit slot hash x.  A lookup loads both slots at once and
// never follows a chain, so a warp's probe costs one memory round trip whatever the other lanes hit.
// The slots come from two different 32-bit folds of x by multiply-shift (the top bits of a
// product depend on every bit of the fold): three 32-bit instructions each on the device instead
// of a 64-bit multiplication.  Tables hold at most 2^32 slots.
constexpr uint32_t kSlotMulA = 0x9E3779B1u, kSlotMulB = 0x85EBCA77u;   // odd
LT_HD uint64_t cuckoo_slot1(uint64_t x, uint32_t bits) {
    const uint32_t f = (uint32_t)x ^ (uint32_t)(x >> 32);
    return (uint64_t)((f * kSlotMulA) >> (32 - bits));
}
LT_HD uint64_t cuckoo_slot2(uint64_t x, uint32_t bits) {
    const uint32_t hi = (uint32_t)(x >> 32);
    const uint32_t f = (uint32_t)x + ((hi << 15) | (hi >> 17));
    return (uint64_t)((f * kSlotMulB) >> (32 - bits));
}
// Fingerprint 0 marks an empty slot: the table builder rejects a key whose fingerprint is 0 (2^-64
// per key), and a lookup whose fingerprint is 0 reads an empty slot as "absent" (its payload is 0).
LT_HD uint64_t dict_fp(H2 h, uint32_t len) { return h.b ^ ((uint64_t)len * 0x9FB21C651E98DF25ull); }

// ---- rule key: up to three code units, exact -------------------------------------------------
LT_HD uint64_t rule_key(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t len) {
    return (uint64_t)c0 | ((uint64_t)c1 << 16) | ((uint64_t)c2 << 32) | ((uint64_t)len << 48);
}

// ---- feature key ------------------------------------------------------------------------------
// kind: 0..8 trigram templates; 16 morpheme preference; 17 word preference.  func: index of the
// owning scorer in the score program.  s0..s2: component strings (zero H2 when unused);
// a0, a1: small integers (tag ids, lengths, flags).
//
// The key is ADDITIVE in its components:
//     ka = seed_a(kind, func) + head * Ta + s0.a * M0a + s1.a * M1a + s2.a * M2a      (mod 2^64)
//     kb = seed_b(kind, func) + head * Tb + s0.b * M0b + s1.b * M1b + s2.b * M2b
// so the beam kernel keeps the products of a word's hashes with the slot multipliers (once per
// lattice edge, once per surviving hypothesis) and forms a transition's keys with additions only.
// ka is the slot hash (cuckoo_slot1 / cuckoo_slot2), kb is the stored fingerprint.
struct FKey {
    uint64_t k1, k2;     // k1 = ka (slot source), k2 = kb (fingerprint; 0 is reserved for empty slots: the builder
                         // rejects such a key, a lookup with it can at worst read an empty slot as weight 0.0)
};

constexpr uint64_t kM0a = 0xE7037ED1A0B428DBull, kM1a = 0x1D8E4E27C47D124Full, kM2a = 0xEB44ACCAB455D165ull;
constexpr uint64_t kM0b = 0x2D358DCCAA6C78A5ull, kM1b = 0x8BB84B93962EACC9ull, kM2b = 0x4B33A62ED433D4A3ull;
constexpr uint64_t kTa = 0x8CB92BA72F3D8DD7ull, kTb = 0xA0761D6478BD642Full;

LT_HD H2 h2_mul(H2 h, uint64_t ma, uint64_t mb) { return H2{h.a * ma, h.b * mb}; }

// inverse of an odd number modulo 2^64 (Newton iteration): the slot multipliers are odd, so a product
// h * M0 can be turned into h * M1 by one multiplication with M1 / M0
constexpr uint64_t inv64(uint64_t a) {
    uint64_t x = a;                     // correct to 3 bits
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}
constexpr uint64_t kM1over0a = kM1a * inv64(kM0a), kM1over0b = kM1b * inv64(kM0b);
constexpr uint64_t kM2over0a = kM2a * inv64(kM0a), kM2over0b = kM2b * inv64(kM0b);
constexpr uint64_t kM2over1a = kM2a * inv64(kM1a), kM2over1b = kM2b * inv64(kM1b);
static_assert(kM0a * inv64(kM0a) == 1 && kM0b * inv64(kM0b) == 1, "inv64");
LT_HD H2 h2_add(H2 x, H2 y) { return H2{x.a + y.a, x.b + y.b}; }

LT_HD H2 feature_seed(uint32_t kind, uint32_t func) {
    const uint64_t t = ((uint64_t)kind << 8) | (uint64_t)func;
    return H2{fmix64(t * 0xD1B54A32D192ED03ull + 0x589965CC75374CC3ull),
              fmix64(t * 0xAEF17502108EF2D9ull + 0x1B03738712FAD5C9ull)};
}

LT_HD uint64_t feature_head(uint32_t a0, uint32_t a1) { return ((uint64_t)a0 << 24) | (uint64_t)a1; }

// key from an already summed component part `sum` = s0*M0 + s1*M1 + s2*M2 (pairwise)
LT_HD FKey feature_key_sum(H2 seed, uint64_t head, H2 sum) {
    FKey k;
    k.k1 = seed.a + head * kTa + sum.a;
    k.k2 = seed.b + head * kTb + sum.b;
    return k;
}

// same key for a head that fits 32 bits (tag ids, flags): a 32 x 64 multiply is cheaper on the device
LT_HD FKey feature_key_sum32(H2 seed, uint32_t head, H2 sum) {
    FKey k;
    k.k1 = seed.a + (uint64_t)head * kTa + sum.a;
    k.k2 = seed.b + (uint64_t)head * kTb + sum.b;
    return k;
}
LT_HD uint32_t feature_head32(uint32_t a0, uint32_t a1) { return (a0 << 24) | a1; }   // a0 < 256, a1 < 2^24

LT_HD FKey feature_key(uint32_t kind, uint32_t func, H2 s0, H2 s1, H2 s2, uint32_t a0, uint32_t a1) {
    H2 sum{s0.a * kM0a + s1.a * kM1a + s2.a * kM2a, s0.b * kM0b + s1.b * kM1b + s2.b * kM2b};
    return feature_key_sum(feature_seed(kind, func), feature_head(a0, a1), sum);
}

// slot hash of a key (cuckoo_slot1 / cuckoo_slot2 turn it into the key's two slots)
LT_HD uint64_t feature_slot_hash(uint64_t k1) { return k1; }

constexpr uint32_t kKindMPref = 16;
constexpr uint32_t kKindWPref = 17;

}  // namespace lt
