// beam.cuh — transition scoring + fixed-window beam search kernel.
//
// Replaces beam_search / Beam / Sequence (beam/beam.py:5-124) and the score functions
// (beam/score_funcs.py:18-144) with their feature templates (features/feature.py:76-121).
//
// One warp per sentence (atomic work queue).  Hypotheses are back-pointer entries in a shared-
// memory ring of the last 9 end positions (window 8, beam.py:30).  An entry holds its score and the
// hash products of its last two words that the NEXT transition's feature keys are sums of
// (hash.cuh: keys are additive), so a transition costs additions, not string hashing.
//
// Per end position e the warp
//   1. makes sure the edges of e's CSR bucket are PREPARED: per edge the word/morpheme hash
//      products and everything of the score that depends on the edge alone — RegularizationScore,
//      the preference scorers, templates 4 and 5.  Dictionary edges are prepared
//      32 consecutive edges at a time into a ring cache (the buckets of consecutive positions are
//      adjacent in HBM, so one full-warp pass serves many positions), unknown words 4 positions x 8
//      spans at a time;
//   2. counts edges per span and candidates per span; an unknown word after an unknown word is
//      allowed only from the window's first begin (beam.py:44-45), such candidates are never
//      enumerated;
//   3. enumerates candidates in the reference's generation order — begin ascending, parent rank
//      ascending, edge order ascending (beam.py:30-48) — 32 at a time, one per lane, and scores
//      each: score program in BeamScoreFunctions order, every fp64 add in the reference's
//      association; templates 0,1,2,7,8 are gathered from the feature table
//      (a cuckoo table: both slots of a key loaded at once), template 3/6 come from shared memory;
//   4. keeps the best `beam` candidates on the order-preserving integer image of the fp64 score
//      with ties resolved towards the earlier candidate, which is exactly the stable sort of
//      Beam.append (beam.py:83-86): rank counting (beam <= 16), a sorting network (<= 32) or
//      arg-max rounds (<= 64);
//   5. writes the survivors as new ring entries and one back-pointer each (shared memory when the
//      host finds room, HBM otherwise).
// The best path is recovered from the back-pointers and written as 16-byte edge records.
//
// The kernel is instantiated per top-K method, beam size, sentence-array size and score program
// (beam_kernel<MODE, KT, UC, PROG>): with those fixed every shared-memory array sits at a constant
// offset and the scorer loop unrolls with its template seeds as immediates — in this latency-bound
// kernel dynamically indexed constant loads and spilled address arithmetic were the largest costs.
#pragma once
#include "lattice.cuh"
#include "tables.cuh"

namespace lt {

constexpr int kRing = LT_WINDOW + 1;
constexpr int kEdgeRing = 64;                 // prepared dictionary edges: ring over the global edge index
constexpr int kBucketCached = 32;             // bucket-local edge indices below this are served from the ring
constexpr int kUnkBlock = 4;                  // end positions whose unknown words are prepared together
constexpr int kCacheSlots = kEdgeRing + kUnkBlock * LT_WINDOW;
constexpr int kRankMaxBeam = 16;              // beams up to this size select by rank counting, larger ones by sorting network
constexpr uint32_t kCtxMask = (1u << LT_TAG_NOUN) | (1u << LT_TAG_ADVERB) | (1u << LT_TAG_ADJECTIVE) | (1u << LT_TAG_VERB);

// entry meta bits
constexpr uint32_t kMetaTagMask = 0xFFu;
constexpr uint32_t kMetaHasI = 1u << 8;
constexpr uint32_t kMetaHasCtx = 1u << 9;
constexpr uint32_t kMetaUnkLenShift = 12;     // min(8, len_j), 4 bits

struct BeamArgs {
    const uint16_t* text;
    const int32_t* sent_off;
    int32_t n_sent;
    int32_t units;              // shared-memory elements per sentence array: >= longest sentence + 8, a multiple of 8
    int32_t beam;
    int32_t warps;              // warps per CTA
    const uint2* pos;           // [n_units] (first edge, edge count) per (sentence, end position)
    const lt_edge* edges;
    const int32_t* status;      // from the lattice pass
    const uint32_t* flags;      // lattice overflow flags (lattice.cuh)
    uint64_t* trail;            // [(n_units) * beam] back-pointers
    lt_edge* path_tmp;          // [n_units] best path, reversed, at the sentence's offset
    int32_t* path_len;          // [n_sent]
    double* scores;             // [n_sent]
    unsigned long long* counters;   // [3]=T [4]=F [5]=Bk [6]=W
    unsigned int* queue;
    const uint32_t* order;      // queue position -> sentence index (longest first), or nullptr
    int32_t trail_smem;         // back-pointers in shared memory (4 B) instead of `trail` (8 B)
    const H2* imp;              // imported lattices: 3 hashes (word, morph0, morph1) per LT_EDGE_EXPLICIT edge
    // all survivors of the last position (beam_search's return value, beam.py:59-61) instead of matures[0] only:
    int32_t kbest;              // 1: also write every survivor's path
    lt_edge* kb_tmp;            // [n_units * beam] path r of sentence s, reversed, at sent_off[s] * beam + r * (raw length)
    int32_t* kb_len;            // [n_sent * beam] words of survivor r (0 beyond the survivors)
    double* kb_scores;          // [n_sent * beam]
    int32_t* kb_count;          // [n_sent] survivors (1 for an empty sentence: [BOS, EOS])
};

// Prepared edges (shared memory, struct of arrays).  Slots [0, kEdgeRing): dictionary edges at
// (global edge index % kEdgeRing); slots kEdgeRing + ((e - 1) % kUnkBlock) * 8 + (span - 1): the
// unknown word of (end position e, span).
struct EdgeCache {
    H2* e0;            // word hash   * M0
    H2* g0;            // morph0 hash * M0 (dictionary edges only: an unknown word is its own morpheme)
    double* kval;      // 2 doubles per scorer: score-program values that depend on the edge only
    uint32_t* meta;    // tag0 | len << 8 (16 bits) | flags << 24
    uint32_t* present; // bit f*2: template 4 present, bit f*2+1: template 5 present (scorer f); bits 24..31: min(255, e - b)
};

// doubles per edge in kval (stride 2 * n_funcs): for scorer f: [2f] = REG / MPREF / WPREF value or
// template-4 weight, [2f+1] = template-5 weight

// The back-pointer trail takes 4 bytes per kept entry in shared memory when the host finds that it
// does not cost residency (BeamArgs::trail_smem); otherwise 8-byte records go to HBM.
// Per-warp shared memory, in this order so that most arrays sit at compile-time offsets:
//   fixed part      edge cache (e0, g0, meta, present), selection pool, per-span tables, counters
//   beam part       ring entries: score, p1, pp, c1, meta, non-unknown ranks        (kRing * beam each)
//   sentence part   ha, hb, CSR row, syllables                                       (units each)
//   kval            edge-only score values, 2 doubles per scorer and cache slot
//   trail           back-pointers, units * beam (when they live in shared memory)
// `units` >= longest sentence of the batch + 8 (a multiple of 8); the common sizes are template
// parameters of the kernel, so that every array but the trail sits at a constant offset.
constexpr size_t kBeamFixedBytes = (size_t)kCacheSlots * 16 + (size_t)kEdgeRing * 16 + 32 * 8 + 64 * 8 +
                                   (size_t)kCacheSlots * 8 + (5 * 16 + 32 + 8) * 4;
static_assert(kBeamFixedBytes % 16 == 0, "fixed part keeps 16-byte alignment");
__host__ __device__ inline size_t beam_ring_bytes(int beam) {
    return (((size_t)kRing * beam * (8 + 48 + 4 + 1)) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t beam_sentence_bytes(int units) { return (size_t)units * (8 + 8 + 8 + 2); }
// doubles of edge-only score values per cache slot: two per scorer
__host__ __device__ inline int beam_kval_doubles(int n_funcs, bool /*reg_tri*/) { return 2 * (n_funcs > 0 ? n_funcs : 1); }
__host__ __device__ inline size_t beam_warp_smem(int units, int beam, int kval_doubles, bool trail_smem) {
    const size_t bytes = kBeamFixedBytes + beam_ring_bytes(beam) + beam_sentence_bytes(units) + (size_t)kCacheSlots * 8 * (size_t)kval_doubles +
                         (trail_smem ? (size_t)units * beam * 4 : 0);
    return (bytes + 15) & ~(size_t)15;
}
// sentence-array sizes with their own kernel instantiation
__host__ __device__ inline int beam_units_class(int lcap) {
    const int need = lcap + 8;
    return need <= 64 ? 64 : (need <= 128 ? 128 : 0);
}

// candidate payload = shared-memory trail entry: (LT_WINDOW - span) << 27 | parent rank << 20 | bucket-local edge index
constexpr uint32_t kPayUnk = 0xFFFFFu;        // edge index of an unknown word

// HBM trail entry: edge reference (global edge index, or kTrailUnk) | span << 32 | parent rank << 40
constexpr uint32_t kTrailUnk = 0xFFFFFFFFu;

struct DenseView {
    const double* t3;
    const double* t4;
    const double* t6;
    const uint32_t* m3;
    const uint32_t* m4;
    const uint32_t* m6;
};

__device__ __forceinline__ DenseView dense_view(const unsigned char* blk, int nt) {
    DenseView d;
    d.t3 = reinterpret_cast<const double*>(blk);
    d.t4 = d.t3 + nt * nt;
    d.t6 = d.t4 + kT4Dense;
    d.m3 = reinterpret_cast<const uint32_t*>(d.t6 + 16);
    d.m4 = d.m3 + nt;
    d.m6 = d.m4 + 2;
    return d;
}

// numpy's ndarray.sum() association for the <= 9 surviving weights
__device__ __forceinline__ double numpy_order_sum9(const double (&v)[9], uint32_t present) {
    const int n = __popc(present);
    if (n < 8) {
        double s = 0.0;
        #pragma unroll
        for (int i = 0; i < 9; ++i)
            if ((present >> i) & 1u) s = __dadd_rn(s, v[i]);
        return s;
    }
    // exactly one of nine missing (n == 8) or none (n == 9): first eight present values -> lanes
    int missing = (n == 9) ? 9 : (__ffs(~present & 0x1FFu) - 1);
    double r[8];
    #pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (j < missing) ? v[j] : v[j + 1];
    double s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    if (n == 9) s = __dadd_rn(s, v[8]);
    return s;
}

struct EdgeView {
    int b, e;
    uint32_t len, tag0, tag1, rule, split, flags;
    H2 wk, mk, m1;      // hashes of word, morph0, morph1
};

__device__ __forceinline__ void unpack_edge(uint4 raw, EdgeView& k) {
    k.b = (int)(raw.x & 0xFFFFu);
    k.e = (int)(raw.x >> 16);
    k.len = raw.y & 0xFFFFu;
    k.tag0 = (raw.y >> 16) & 0xFFu;
    k.tag1 = (raw.y >> 24) & 0xFFu;
    k.rule = raw.z;
    k.split = raw.w & 0xFFFFu;
    k.flags = (raw.w >> 16) & 0xFFu;
}

__device__ __forceinline__ void unknown_edge(int b, int e, EdgeView& k) {
    k.b = b; k.e = e;
    k.len = (uint32_t)(e - b); k.tag0 = LT_TAG_UNK; k.tag1 = LT_NO_TAG; k.rule = LT_NO_RULE;
    k.split = 0; k.flags = LT_EDGE_UNK;
}

// IMP: the batch may hold an imported lattice (lt_lattice_import) — only the all-survivors kernels (KB = 1)
// are launched on one, so the throughput instantiations compile the test out.
template <int IMP>
__device__ __noinline__ void edge_hashes(const DevTables& T, const SentView& v, EdgeView& k, bool need_m1) {
    if (IMP != 0 && (k.flags & LT_EDGE_EXPLICIT)) {
        // imported lattice (lt_lattice_import): the strings of this word were hashed on the host
        const H2* h = v.imp + 3 * (size_t)k.rule;
        k.wk = h[0];
        k.mk = h[1];
        k.m1 = h[2];
        return;
    }
    k.wk = sub_hash(T, v, k.b, k.e);
    k.mk = k.wk;
    k.m1 = H2{0, 0};
    if (k.flags & LT_EDGE_LEMMA) {
        const int p = k.b + (int)k.split;
        if (k.rule == LT_NO_RULE) {
            k.mk = sub_hash(T, v, k.b, p + 1);
            if (need_m1) k.m1 = sub_hash(T, v, p + 1, k.e);
        } else {
            RuleRec rec = rule_load(T, k.rule);
            H2 pre = (p > k.b) ? sub_hash(T, v, k.b, p) : H2{0, 0};
            k.mk = h2_concat(pre, rec.stem, pow_at(T, rec.stem_len));
            if (need_m1) {
                int from = p + ((k.flags & LT_EDGE_SKIP2) ? 2 : 1);
                H2 suf{0, 0};
                uint32_t sl = 0;
                if (from < k.e) { suf = sub_hash(T, v, from, k.e); sl = (uint32_t)(k.e - from); }
                k.m1 = h2_concat(rec.eomi, suf, pow_at(T, sl));
            }
        }
    }
}

// Everything of a transition's score that depends on the edge alone, for scorer f:
//   REG / MPREF / WPREF: a = the scorer's value;  TRIGRAM: a = template 4 weight, b2 = template 5
//   weight, presence bits 0 / 1.
__device__ __forceinline__ uint32_t edge_score_body(const DevTables& T, const unsigned char* dense_blk, const EdgeView& k,
                                                    H2 e0, H2 g0, int f, int kind, H2 seed4, H2 seed5, H2 seed_pref,
                                                    double& a, double& b2) {
    uint32_t present = 0;
    const uint32_t tk = k.tag0;
    const lt_func& fn = T.funcs[f];
    a = 0.0;
    b2 = 0.0;
    if (kind == LT_FUNC_REG) {
        // score_funcs.py:65-73
        if (tk == LT_TAG_UNK) a = __dmul_rn(fn.p[0], __dadd_rn((double)k.len, 0.1));
        else a = __dmul_rn(fn.p[1], (double)k.len);
        a = __dadd_rn(0.0, a);
        if (k.len == 1 && tk == LT_TAG_NOUN) a = __dadd_rn(a, fn.p[2]);
    } else if (kind == LT_FUNC_MPREF) {
        // score_funcs.py:84-88
        FKey k0 = feature_key_sum32(seed_pref, feature_head32(tk, 0), g0);
        FeatProbe s0 = feat_first(T, k0);
        if (k.tag1 != LT_NO_TAG) {
            FKey k1 = feature_key_sum32(seed_pref, feature_head32(k.tag1, 0), h2_mul(k.m1, kM0a, kM0b));
            FeatProbe s1 = feat_first(T, k1);
            feat_resolve(T, k1, s1, b2);
        }
        feat_resolve(T, k0, s0, a);
        if (k.tag1 != LT_NO_TAG) a = __dadd_rn(a, b2);
        b2 = 0.0;
    } else if (kind == LT_FUNC_WPREF) {
        // score_funcs.py:99-100
        FKey k0 = feature_key_sum32(seed_pref, feature_head32(tk, 0), e0);
        FeatProbe s0 = feat_first(T, k0);
        feat_resolve(T, k0, s0, a);
    } else {
        // templates 4 (wk.len) and 5 (wk.word, wk.tag0, wk.is_l), features/feature.py:100,104
        const DenseView D = dense_view(dense_blk, T.n_tags);
        FKey q5 = feature_key_sum32(seed5, feature_head32(tk, (k.flags & LT_EDGE_IS_L) ? 1u : 0u), e0);
        FeatProbe s5 = feat_first(T, q5);
        if (k.len >= (uint32_t)kT4Dense) {
            FKey q4 = feature_key_sum(seed4, feature_head(k.len, 0), H2{0, 0});
            FeatProbe s4 = feat_first(T, q4);
            if (feat_resolve(T, q4, s4, a)) present |= 1u;
        } else if ((D.m4[k.len >> 5] >> (k.len & 31)) & 1u) {
            a = D.t4[k.len];
            present |= 1u;
        }
        if (feat_resolve(T, q5, s5, b2)) present |= 2u;
    }
    return present;
}

// any score program: kind, dense block and seeds of scorer f read from the tables
__device__ __noinline__ uint32_t edge_score(const DevTables& T, const unsigned char* dense_smem, const EdgeView& k,
                                            H2 e0, H2 g0, int f, double& a, double& b2) {
    const int kind = T.funcs[f].kind;
    const unsigned char* dense_blk = dense_smem;
    if (kind == LT_FUNC_TRIGRAM) dense_blk += (size_t)T.func_dense[f] * dense_block_bytes(T.n_tags);
    return edge_score_body(T, dense_blk, k, e0, g0, f, kind, T.seeds[f][4], T.seeds[f][5], T.seeds[f][9], a, b2);
}

// numpy's association from eight surviving weights on: rare, so the nine weights are
// simply gathered again and summed by numpy_order_sum9.
__device__ __noinline__ double trigram_sum_tree(const DevTables& T, const unsigned char* dense_blk, int NT, int f, FKey q0, FKey q1,
                                                FKey q2, FKey q7, FKey q8, uint32_t tj, uint32_t tk, uint32_t epresent,
                                                double val4, double val5, bool j_unk, uint32_t ul, bool has_i, bool ctx8) {
    const DenseView D = dense_view(dense_blk, NT);
    double w[9];
    uint32_t present = 0;
    #pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = 0.0;
    if (feat_resolve(T, q0, feat_first(T, q0), w[0])) present |= 1u << 0;
    if (feat_resolve(T, q1, feat_first(T, q1), w[1])) present |= 1u << 1;
    if (feat_resolve(T, q2, feat_first(T, q2), w[2])) present |= 1u << 2;
    if ((D.m3[tj] >> tk) & 1u) { w[3] = D.t3[tj * NT + tk]; present |= 1u << 3; }
    if ((epresent >> (2 * f)) & 1u) { w[4] = val4; present |= 1u << 4; }
    if ((epresent >> (2 * f + 1)) & 1u) { w[5] = val5; present |= 1u << 5; }
    if (j_unk && ((D.m6[0] >> ul) & 1u)) { w[6] = D.t6[ul]; present |= 1u << 6; }
    if (has_i && feat_resolve(T, q7, feat_first(T, q7), w[7])) present |= 1u << 7;
    if (ctx8 && feat_resolve(T, q8, feat_first(T, q8), w[8])) present |= 1u << 8;
    return numpy_order_sum9(w, present);
}

// order-preserving integer image of an fp64 score (larger score -> larger key); 0 is "no candidate"
__device__ __forceinline__ uint64_t sortable(double s) {
    uint64_t bits = (uint64_t)__double_as_longlong(s);
    return bits ^ ((bits >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
}
__device__ __forceinline__ double unsortable(uint64_t key) {
    uint64_t bits = key ^ ((key >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
    return __longlong_as_double((long long)bits);
}

#ifndef LT_BEAM_MINB
#define LT_BEAM_MINB 2
#endif
#ifndef LT_BEAM_MAXW
#define LT_BEAM_MAXW 8         // largest CTA in warps (with LT_BEAM_MINB: the register budget the kernels are compiled for)
#endif
#ifndef LT_BEAM_REG_WARPS
#define LT_BEAM_REG_WARPS 16   // warps per SM that budget allows (the host's residency estimate)
#endif
constexpr int kBeamMaxWarpsC = LT_BEAM_MAXW;
#ifndef LT_PROBE_SPLIT
#define LT_PROBE_SPLIT 1       // 1: generic kernels issue the loads of templates 7 and 8 after templates 0..2 are consumed
#endif
// (launch bounds: 128 registers per thread for every instantiation.  Compiling the small-beam instantiations for
// 5 resident 4-warp CTAs — 96 registers, with the shared-memory diet that makes room for the fifth — spills ~120
// bytes per thread and runs 16 % slower than 4 CTAs at 128 registers: profiles/README.md, r2d.)
constexpr int beam_max_threads(int, int, int, int) { return kBeamMaxWarpsC * 32; }
constexpr int beam_min_blocks(int, int, int, int) { return LT_BEAM_MINB; }
constexpr int kBeamWarps = 4;                 // preferred warps per CTA of the beam kernel
constexpr int kBeamMaxWarps = kBeamMaxWarpsC;              // largest CTA (128 registers per thread either way: 8 warps x 2 CTAs = 4 warps x 4 CTAs)

// Edge prep: hash products and the edge-only part of the score program into cache slot `slot`.
template <int PROG, int IMP>
__device__ __forceinline__ void prep_edge(const DevTables& T, const SentView& v, const unsigned char* dense_smem,
                                          EdgeView& k, bool need_m1, int nf, int kvs, const EdgeCache& C, uint32_t slot) {
    edge_hashes<IMP>(T, v, k, need_m1);
    const H2 e0 = h2_mul(k.wk, kM0a, kM0b), g0 = h2_mul(k.mk, kM0a, kM0b);
    uint32_t present = 0;
    if (PROG == 1) {
        // (RegularizationScore, SimpleTrigramFeatureScore): both scorers inline, the trigram seeds as immediates,
        // the regulariser computed while the template-5 probe is in flight
        double reg, unused, w4, w5;
        const uint32_t tri = edge_score_body(T, dense_smem, k, e0, g0, 1, LT_FUNC_TRIGRAM, feature_seed(4u, 1u), feature_seed(5u, 1u),
                                             H2{0, 0}, w4, w5);
        edge_score_body(T, dense_smem, k, e0, g0, 0, LT_FUNC_REG, H2{0, 0}, H2{0, 0}, H2{0, 0}, reg, unused);
        (void)unused;
        present = tri << 2;
        C.kval[slot * 4 + 0] = reg;
        C.kval[slot * 4 + 1] = 0.0;
        C.kval[slot * 4 + 2] = w4;
        C.kval[slot * 4 + 3] = w5;
    } else {
        #pragma unroll 1
        for (int f = 0; f < nf; ++f) {
            double a, b2;
            present |= edge_score(T, dense_smem, k, e0, g0, f, a, b2) << (2 * f);
            C.kval[slot * kvs + 2 * f] = a;
            C.kval[slot * kvs + 2 * f + 1] = b2;
        }
    }
    C.e0[slot] = e0;
    if (slot < (uint32_t)kEdgeRing) C.g0[slot] = g0;
    C.meta[slot] = k.tag0 | (k.len << 8) | (k.flags << 24);
    const uint32_t span = (uint32_t)(k.e - k.b);
    C.present[slot] = present | ((span < 255u ? span : 255u) << 24);
}

// MODE selects the top-K of a position's candidates:
//   2  rank by counting (beam <= kRankMaxBeam): every candidate counts the pool entries that beat it
//   1  32-lane bitonic sorting network per chunk + bitonic merge with the kept list (beam <= 32)
//   0  rounds of warp arg-max with two kept entries per lane (beam 33..64)
// KT: the beam size, UC: the sentence-array size when known at compile time (array offsets become
// constants, which is what keeps the kernel's address arithmetic out of registers); 0 = A.beam / A.units.
// PROG: 1 = the score program is exactly (RegularizationScore, SimpleTrigramFeatureScore) — the
// scorer loop of a candidate unrolls, the template seeds become immediates; 0 = read it from T.
// KB: 1 = every survivor of the last position is written out as well (lt_beam_kbest); 0 compiles that out of
// the instantiations the throughput path runs.
template <int MODE, int KT, int UC, int PROG, int KB = 0>
__global__ void __launch_bounds__(beam_max_threads(KT, UC, PROG, KB), beam_min_blocks(KT, UC, PROG, KB)) beam_kernel(const __grid_constant__ DevTables T, const __grid_constant__ BeamArgs A) {
    constexpr int KR = (MODE == 0) ? 2 : 1;      // kept entries per lane
    LT_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int K = KT ? KT : A.beam;
    const int NT = T.n_tags;

    // CTA-shared dense tables (tag x tag matrix, length vectors)
    const size_t dense_bytes = ((size_t)T.n_tri * dense_block_bytes(NT) + 15) & ~(size_t)15;
    unsigned char* dense_smem = smem_raw;
    lt_pdl_trigger();
    // (the score tables are constant: staging them does not wait for the lattice kernel)
    for (size_t i = threadIdx.x * 4; i < (size_t)T.n_tri * dense_block_bytes(NT); i += blockDim.x * 4)
        *reinterpret_cast<uint32_t*>(dense_smem + i) = *reinterpret_cast<const uint32_t*>(T.dense + i);
    __syncthreads();
    lt_pdl_wait();          // the lattice of this batch is complete
    if (A.flags[kFlagEdgeOverflow] | A.flags[kFlagStageOverflow]) {
        // lattice incomplete: the host grows the buffer and reruns; the path lengths the scan / pack kernels
        // behind this launch read must still be defined
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= A.n_sent; i += (int64_t)gridDim.x * blockDim.x)
            A.path_len[i] = 0;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) A.path_len[A.n_sent] = 0;      // the scan runs over n_sent + 1 entries

    const bool trail_smem = A.trail_smem != 0;
    const int units = UC ? UC : A.units;
    const int RK = kRing * K;
    const int kvs = beam_kval_doubles(T.n_funcs, PROG == 1);   // kval stride
    unsigned char* wbase = smem_raw + dense_bytes + (size_t)warp * beam_warp_smem(units, K, kvs, trail_smem);
    // fixed part
    EdgeCache C;
    C.e0 = reinterpret_cast<H2*>(wbase);
    C.g0 = C.e0 + kCacheSlots;
    uint64_t* s_newkey = reinterpret_cast<uint64_t*>(C.g0 + kEdgeRing);   // [32] selected keys by rank
    uint64_t* s_pool = s_newkey + 32;               // [64] selection pool: kept entries, then the chunk's lanes
    C.meta = reinterpret_cast<uint32_t*>(s_pool + 64);
    C.present = C.meta + kCacheSlots;
    uint32_t* s_cnt = C.present + kCacheSlots;      // [9] edges per span
    uint32_t* s_gstart = s_cnt + 16;                // [9] bucket-local index of the span's first edge
    uint32_t* s_tlist = s_gstart + 16;              // [8] spans that generate candidates, in generation order: first candidate << 4 | span
    uint32_t* s_nbeam = s_tlist + 16;               // [kRing] entries per ring slot
    uint32_t* s_nnon = s_nbeam + 16;                // [kRing] entries per ring slot that do not end in an unknown word
    uint32_t* s_newpay = s_nnon + 16;               // [32] selected payloads by rank
    uint32_t* s_acc = s_newpay + 32;                // [4] work counters of this warp: T, F, Bk, W (+ 4 spare)
    // beam part
    double* e_score = reinterpret_cast<double*>(wbase + kBeamFixedBytes);
    H2* e_p1 = reinterpret_cast<H2*>(e_score + RK);             // wj * M1
    H2* e_pp = e_p1 + RK;                                       // wj * M1 + wi * M2
    H2* e_c1 = e_pp + RK;                                       // contextual morph * M1
    uint32_t* e_meta = reinterpret_cast<uint32_t*>(e_c1 + RK);
    uint8_t* s_nonunk = reinterpret_cast<uint8_t*>(e_meta + RK);   // [kRing * K] ranks of the non-unknown entries, ascending
    // sentence part
    uint64_t* ha = reinterpret_cast<uint64_t*>(wbase + kBeamFixedBytes + beam_ring_bytes(K));
    uint64_t* hb = ha + units;
    uint2* spos = reinterpret_cast<uint2*>(hb + units);
    uint16_t* ch = reinterpret_cast<uint16_t*>(spos + units);
    C.kval = reinterpret_cast<double*>(ch + units);
    uint32_t* s_trail = reinterpret_cast<uint32_t*>(C.kval + kCacheSlots * kvs);   // [units * K] when trail_smem

    if (lane < 4) s_acc[lane] = 0;

    bool need_m1 = false;
    for (int f = 0; f < T.n_funcs; ++f) need_m1 |= (T.funcs[f].kind == LT_FUNC_MPREF);
    const int nf = PROG == 1 ? 2 : T.n_funcs;

    while (true) {
        unsigned int s = 0;
        if (lane == 0) s = atomicAdd(A.queue, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= (unsigned)A.n_sent) break;
        if (A.order) s = __ldg(A.order + s);
        const int s0 = __ldg(A.sent_off + s), s1 = __ldg(A.sent_off + s + 1);
        const int st = __ldg(A.status + s);
        // ---- stage syllables, prefix hashes and the sentence's CSR row ----
        int L = 0;
        const int raw_len = (st == LT_SENT_OK) ? (s1 - s0) : 0;
        for (int base = 0; base < raw_len; base += 32) {
            const int idx = s0 + base + lane;
            const bool valid = base + lane < raw_len;
            const uint32_t c = valid ? (uint32_t)__ldg(A.text + idx) : 0x20u;
            const bool keep = valid && (c != 0x20u);
            const unsigned km = __ballot_sync(kFull, keep);
            const int pos = L + __popc(km & lt_mask);
            if (keep) ch[pos] = (uint16_t)c;
            L += __popc(km);
        }
        __syncwarp();
        for (int i = lane; i < L; i += 32) spos[i] = __ldg(A.pos + s0 + i);
        prefix_hashes_inline(ch, L, lane, ha, hb);
        SentView v{ch, ha, hb, nullptr, KB ? A.imp : nullptr};

        if (L == 0 && lane == 0) { A.path_len[s] = 0; A.scores[s] = 0.0; }

        // beam[0] = [BOS] (beam.py:21-23)
        if (lane == 0) {
            e_score[0] = 0.0;
            e_p1[0] = h2_mul(T.bos, kM1a, kM1b);
            e_pp[0] = H2{0, 0};
            e_c1[0] = H2{0, 0};
            e_meta[0] = (uint32_t)LT_TAG_BOS;
            s_nbeam[0] = 1;
            s_nnon[0] = 1;
            s_nonunk[0] = 0;
        }
        __syncwarp();

        uint32_t ring_lo = 0, ring_hi = 0;      // global edge indices [ring_lo, ring_hi) are prepared

        for (int e = 1; e <= L; ++e) {
            const int slot_e = e % kRing;
            const uint2 bucket = spos[e - 1];
            const uint32_t es = bucket.x, ne = bucket.y;
            const int jmax = (e < LT_WINDOW) ? e : LT_WINDOW;

            // ---- 1. EDGE PREP of dictionary edges, 32 consecutive edges at a time ----
            // Buckets of consecutive end positions are adjacent in HBM (one reservation per sentence, edges
            // ranked by end), so one pass normally prepares the edges of many positions ahead.
            if (ne > 0) {
                const uint32_t need_hi = es + (ne < (uint32_t)kBucketCached ? ne : (uint32_t)kBucketCached);
                if (es < ring_lo || es > ring_hi || need_hi > ring_hi) {
                    uint32_t start = ring_hi;
                    if (es < ring_lo || es > ring_hi) { start = es; ring_lo = es; }
                    // how far beyond this bucket the sentence's edges continue without a gap
                    const int p = e + lane;                       // 0-based index of end position e + 1 + lane
                    const uint2 nb = (p < L) ? spos[p] : make_uint2(0u, 0u);
                    uint32_t incl = nb.y;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(kFull, incl, d);
                        if (lane >= d) incl += t;
                    }
                    const bool gap = (p >= L) || (nb.y != 0 && nb.x != es + ne + (incl - nb.y));
                    const unsigned gaps = __ballot_sync(kFull, gap);
                    const int run = gaps ? __ffs(gaps) - 1 : 32;     // positions ahead that continue the range
                    const uint32_t ahead = run ? __shfl_sync(kFull, incl, run - 1) : 0u;
                    const uint32_t avail = es + ne + ahead - start;
                    const uint32_t n = avail < 32u ? avail : 32u;
                    if ((uint32_t)lane < n) {
                        const uint32_t gi = start + lane;
                        EdgeView k;
                        unpack_edge(ldg16(A.edges + gi), k);
                        prep_edge<PROG, KB>(T, v, dense_smem, k, need_m1, nf, kvs, C, gi & (kEdgeRing - 1));
                    }
                    ring_hi = start + n;
                    if (ring_hi - ring_lo > (uint32_t)kEdgeRing) ring_lo = ring_hi - kEdgeRing;
                    __syncwarp();
                }
            }
            // ---- the unknown words of kUnkBlock end positions at a time, lanes = (position, span) ----
            if (((e - 1) & (kUnkBlock - 1)) == 0) {
                const int pe = e + (lane >> 3);
                const int j = (lane & 7) + 1;
                if (pe <= L && j <= pe) {
                    EdgeView k;
                    unknown_edge(pe - j, pe, k);
                    prep_edge<PROG, KB>(T, v, dense_smem, k, need_m1, nf, kvs, C, (uint32_t)(kEdgeRing + lane));
                }
                __syncwarp();
            }
            const uint32_t unk_base = (uint32_t)(kEdgeRing + (((e - 1) & (kUnkBlock - 1)) << 3) - 1);   // + span

            // ---- 2. edges per span (bucket sorted by begin ascending = span descending) ----
            if (lane >= 1 && lane <= LT_WINDOW) s_cnt[lane] = 0;
            __syncwarp();
            for (uint32_t base = 0; base < ne; base += 32) {
                const uint32_t idx = base + lane;
                const unsigned act = __ballot_sync(kFull, idx < ne);
                if (idx < ne) {
                    uint32_t span;
                    if (idx < (uint32_t)kBucketCached) {
                        span = C.present[(es + idx) & (kEdgeRing - 1)] >> 24;
                    } else {
                        const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(A.edges + es + idx));
                        span = (x >> 16) - (x & 0xFFFFu);
                    }
                    const unsigned same = __match_any_sync(act, span);
                    if (span <= (uint32_t)LT_WINDOW && (same & lt_mask) == 0) {
                        // first lane of the span's group in this chunk (a group may continue from the previous chunk)
                        const uint32_t old = s_cnt[span];
                        if (old == 0) s_gstart[span] = idx;
                        s_cnt[span] = old + __popc(same);
                    }
                }
                __syncwarp();
            }
            // lane t < 8 owns span LT_WINDOW - t: generation order (begin ascending) is lane order
            uint32_t span_nc = 0, span_c0 = 0;      // candidates of the lane's span, index of its first candidate
            uint32_t N;
            {
                const int jj = LT_WINDOW - lane;
                if (lane < LT_WINDOW && jj <= jmax) {
                    const uint32_t c = s_cnt[jj];
                    const int ps = (e - jj) % kRing;
                    // an unknown word may follow an unknown word only from the window's first begin (beam.py:44-45):
                    // elsewhere only the parents that do not end in an unknown word generate a candidate
                    span_nc = c ? s_nbeam[ps] * c : ((jj < jmax) ? s_nnon[ps] : s_nbeam[ps]);
                }
                uint32_t incl = span_nc;
                #pragma unroll
                for (int d = 1; d < LT_WINDOW; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, incl, d);
                    if (lane >= d) incl += t;
                }
                span_c0 = incl - span_nc;
                N = __shfl_sync(kFull, incl, LT_WINDOW - 1);
                const unsigned gen = __ballot_sync(kFull, span_nc > 0);
                if (span_nc > 0) s_tlist[__popc(gen & lt_mask)] = (span_c0 << 4) | (uint32_t)jj;
            }
            __syncwarp();

            // ---- 3 + 4. candidates in generation order, running top-K ----
            uint64_t keep_key[KR];
            uint32_t keep_pay[KR];
            #pragma unroll
            for (int r = 0; r < KR; ++r) { keep_key[r] = 0; keep_pay[r] = 0; }
            uint32_t nk = 0;      // kept entries so far (MODE 2)

            for (uint32_t c0 = 0; c0 < N; c0 += 32) {
                const uint32_t c = c0 + lane;
                bool valid = c < N;
                // the span of candidate c: spans that start inside this chunk vote their first lane
                const bool gen = span_nc > 0;
                const unsigned starts = __reduce_or_sync(kFull, (gen && span_c0 >= c0 && span_c0 < c0 + 32u) ? (1u << (span_c0 - c0)) : 0u);
                const uint32_t before = __popc(__ballot_sync(kFull, gen && span_c0 < c0));
                int j = 0;
                uint32_t rem = 0;
                if (valid) {
                    const uint32_t tl = s_tlist[before + __popc(starts & (0xFFFFFFFFu >> (31 - lane))) - 1];
                    j = (int)(tl & 15u);
                    rem = c - (tl >> 4);
                }
                uint64_t ckey = 0;
                uint32_t cpay = 0;
                uint32_t cand_F = 0;        // feature tuples this candidate generates
                if (valid) {
                    const uint32_t cj = s_cnt[j];
                    const bool unk_edge = (cj == 0);
                    const int pbase = ((e - j) % kRing) * K;
                    uint32_t prank, eidx = 0;
                    if (unk_edge) {
                        prank = (j < jmax) ? (uint32_t)s_nonunk[pbase + rem] : rem;
                    } else {
                        prank = (cj == 1u) ? rem : rem / cj;
                        eidx = rem - prank * cj;
                    }
                    const int pslot = pbase + (int)prank;
                    const uint32_t pmeta = e_meta[pslot];
                    const uint32_t tj = pmeta & kMetaTagMask;
                    const uint32_t bidx = unk_edge ? 0u : s_gstart[j] + eidx;      // bucket-local edge index
                    const uint32_t slot = unk_edge ? unk_base + (uint32_t)j : ((es + bidx) & (kEdgeRing - 1));
                    H2 e0, g0;
                    uint32_t emeta, epresent = 0;
                    const bool uncached = !unk_edge && bidx >= (uint32_t)kBucketCached;
                    EdgeView kfly;
                    if (uncached) {
                        // bucket larger than the cached part: prepare this edge on the fly
                        unpack_edge(ldg16(A.edges + es + bidx), kfly);
                        edge_hashes<KB>(T, v, kfly, need_m1);
                        e0 = h2_mul(kfly.wk, kM0a, kM0b);
                        g0 = h2_mul(kfly.mk, kM0a, kM0b);
                        emeta = kfly.tag0 | (kfly.len << 8) | (kfly.flags << 24);
                    } else {
                        e0 = C.e0[slot];
                        g0 = unk_edge ? e0 : C.g0[slot];
                        emeta = C.meta[slot];
                        epresent = C.present[slot];
                    }
                    const uint32_t tk = emeta & 0xFFu;
                    // (only a dictionary that files entries under the tag 'Unknown' can still meet beam.py:44-45 here)
                    valid = !(tj == LT_TAG_UNK && tk == LT_TAG_UNK && j < jmax);
                    const double pscore = e_score[pslot];
                    const H2 p1 = e_p1[pslot];
                    const bool has_i = (pmeta & kMetaHasI) != 0;
                    const bool j_unk = (tj == LT_TAG_UNK);
                    const bool ctx8 = ((kCtxMask >> tk) & 1u) && (pmeta & kMetaHasCtx);
                    double inc = 0.0;
                    #pragma unroll (PROG == 1 ? 2 : 1)
                    for (int f = 0; f < nf; ++f) {
                        const int kind = PROG == 1 ? (f == 0 ? LT_FUNC_REG : LT_FUNC_TRIGRAM) : T.funcs[f].kind;
                        double val, val5;
                        if (uncached) {
                            double uv, uv5;      // (by reference to a call: keep val / val5 themselves in registers)
                            epresent |= edge_score(T, dense_smem, kfly, e0, g0, f, uv, uv5) << (2 * f);
                            val = uv;
                            val5 = uv5;
                        } else {
                            val = C.kval[slot * kvs + 2 * f];
                            val5 = C.kval[slot * kvs + 2 * f + 1];
                        }
                        if (kind == LT_FUNC_TRIGRAM) {
                            // SimpleTrigramFeatureScore.score (score_funcs.py:137-144)
                            const unsigned char* dense_blk = dense_smem + (PROG == 1 ? 0 : (size_t)T.func_dense[f] * dense_block_bytes(NT));
                            const H2 sd0 = PROG == 1 ? feature_seed(0u, (uint32_t)f) : T.seeds[f][0];
                            const H2 sd1 = PROG == 1 ? feature_seed(1u, (uint32_t)f) : T.seeds[f][1];
                            const H2 sd2 = PROG == 1 ? feature_seed(2u, (uint32_t)f) : T.seeds[f][2];
                            const H2 sd7 = PROG == 1 ? feature_seed(7u, (uint32_t)f) : T.seeds[f][7];
                            const H2 sd8 = PROG == 1 ? feature_seed(8u, (uint32_t)f) : T.seeds[f][8];
                            const DenseView D = dense_view(dense_blk, NT);
                            cand_F += valid ? 6u + (j_unk ? 1u : 0u) + (has_i ? 1u : 0u) + (ctx8 ? 1u : 0u) : 0u;
                            const H2 pp = has_i ? e_pp[pslot] : H2{0, 0};
                            const H2 c1v = ctx8 ? e_c1[pslot] : H2{0, 0};
                            const uint32_t hk = feature_head32(tk, 0);
                            const FKey q0 = feature_key_sum32(sd0, hk, h2_add(e0, p1));
                            const FKey q1 = feature_key_sum32(sd1, hk, p1);
                            const FKey q2 = feature_key_sum32(sd2, feature_head32(tj, tk), e0);
                            const FKey q7 = feature_key_sum32(sd7, 0u, h2_add(e0, pp));
                            const FKey q8 = feature_key_sum32(sd8, 0u, h2_add(g0, c1v));
                            // all first-slot loads in flight before any is consumed
                            const FeatProbe s0 = feat_first(T, q0);
                            const FeatProbe s1 = feat_first(T, q1);
                            const FeatProbe s2 = feat_first(T, q2);
                            FeatProbe s7, s8;
                            // generic score programs issue templates 7 / 8 after 0..2 are consumed (all five at once
                            // spill there); the specialised kernel has the registers for a single round trip
                            constexpr bool kSplitProbes = LT_PROBE_SPLIT && PROG != 1;
                            if constexpr (!kSplitProbes) {
                                if (has_i) s7 = feat_first(T, q7);
                                if (ctx8) s8 = feat_first(T, q8);
                            }
                            // running left-to-right sum = numpy's order while fewer than 8 weights survive
                            double acc = 0.0, w;
                            int n = 0;
                            if (feat_resolve(T, q0, s0, w)) { acc = __dadd_rn(acc, w); ++n; }
                            if (feat_resolve(T, q1, s1, w)) { acc = __dadd_rn(acc, w); ++n; }
                            if (feat_resolve(T, q2, s2, w)) { acc = __dadd_rn(acc, w); ++n; }
                            if constexpr (kSplitProbes) {
                                if (has_i) s7 = feat_first(T, q7);
                                if (ctx8) s8 = feat_first(T, q8);
                            }
                            if ((D.m3[tj] >> tk) & 1u) { acc = __dadd_rn(acc, D.t3[tj * NT + tk]); ++n; }
                            if ((epresent >> (2 * f)) & 1u) { acc = __dadd_rn(acc, val); ++n; }
                            if ((epresent >> (2 * f + 1)) & 1u) { acc = __dadd_rn(acc, val5); ++n; }
                            const uint32_t ul = (pmeta >> kMetaUnkLenShift) & 0xFu;
                            if (j_unk && ((D.m6[0] >> ul) & 1u)) { acc = __dadd_rn(acc, D.t6[ul]); ++n; }
                            if (has_i && feat_resolve(T, q7, s7, w)) { acc = __dadd_rn(acc, w); ++n; }
                            if (ctx8 && feat_resolve(T, q8, s8, w)) { acc = __dadd_rn(acc, w); ++n; }
                            if (n >= 8)   // numpy switches to an 8-lane tree: redo the gather and add in that order (rare)
                                acc = trigram_sum_tree(T, dense_blk, NT, f, q0, q1, q2, q7, q8, tj, tk, epresent, val, val5, j_unk, ul,
                                                       has_i, ctx8);
                            val = n ? acc : 0.0;
                        }
                        inc = __dadd_rn(inc, val);          // score += f(...), score_funcs.py:51-53
                    }
                    double newscore = __dadd_rn(pscore, inc);        // Sequence.add, beam.py:115
                    newscore = __dadd_rn(newscore, 0.0);             // -0.0 sorts as 0.0
                    ckey = valid ? sortable(newscore) : 0ull;
                    // payload doubles as the generation ordinal: begin ascending (= span descending), parent
                    // rank ascending, edge order ascending (beam.py:30-48)
                    cpay = ((uint32_t)(LT_WINDOW - j) << 27) | (prank << 20) | (unk_edge ? kPayUnk : bidx);
                }
                // work counters: scored transitions and generated feature tuples of this chunk
                const uint32_t chunk_T = __popc(__ballot_sync(kFull, ckey != 0));
                {
                    const uint32_t chunk_F = __reduce_add_sync(kFull, cand_F);
                    if (lane == 0) { s_acc[0] += chunk_T; s_acc[1] += chunk_F; }
                }
                if constexpr (MODE == 2) {
                    // ---- top-K by rank counting: an entry's rank = number of pool entries that beat it ----
                    // pool = kept entries (earlier candidates, they win ties) + this chunk's lanes in generation order
                    const uint32_t nc = (N - c0 < 32u) ? (N - c0) : 32u;
                    const uint64_t kkey = keep_key[0];
                    s_pool[32 + lane] = ckey;
                    if ((uint32_t)lane < nk) s_pool[lane] = kkey;
                    __syncwarp();
                    uint32_t gt_c = 0, gt_k = 0;
                    if (nk == 0) {
                        #pragma unroll 4
                        for (uint32_t l = 0; l < nc; ++l) gt_c += (s_pool[32 + l] > ckey) ? 1u : 0u;
                    } else {
                        #pragma unroll 4
                        for (uint32_t l = 0; l < nc; ++l) {
                            const uint64_t o = s_pool[32 + l];
                            gt_c += (o > ckey) ? 1u : 0u;
                            gt_k += (o > kkey) ? 1u : 0u;
                        }
                        const uint64_t cm1 = ckey - 1;        // a kept entry with an equal key is the earlier candidate
                        #pragma unroll 4
                        for (uint32_t r = 0; r < nk; ++r) gt_c += (s_pool[r] > cm1) ? 1u : 0u;
                    }
                    // equal keys inside the chunk: the earlier lane first
                    gt_c += __popc(__match_any_sync(kFull, ckey) & lt_mask);
                    const uint32_t nvalid = chunk_T;
                    if (ckey != 0 && gt_c < (uint32_t)K) { s_newkey[gt_c] = ckey; s_newpay[gt_c] = cpay; }
                    if ((uint32_t)lane < nk && lane + gt_k < (uint32_t)K) { s_newkey[lane + gt_k] = kkey; s_newpay[lane + gt_k] = keep_pay[0]; }
                    __syncwarp();
                    nk = (nk + nvalid < (uint32_t)K) ? nk + nvalid : (uint32_t)K;
                    keep_key[0] = ((uint32_t)lane < nk) ? s_newkey[lane] : 0ull;
                    keep_pay[0] = ((uint32_t)lane < nk) ? s_newpay[lane] : 0u;
                    __syncwarp();
                } else if constexpr (MODE == 1) {
                    // ---- top-K by sorting network: sort the chunk (best first), then merge with the kept list ----
                    // order: larger key first; equal keys: smaller payload (= earlier candidate) first
                    uint64_t bk = ckey;
                    uint32_t bp = cpay;
                    #pragma unroll
                    for (int size = 2; size <= 32; size <<= 1) {
                        #pragma unroll
                        for (int stride = size >> 1; stride > 0; stride >>= 1) {
                            const uint64_t ok = __shfl_xor_sync(kFull, bk, stride);
                            const uint32_t op = __shfl_xor_sync(kFull, bp, stride);
                            const bool mine_first = (bk > ok) || (bk == ok && bp < op);
                            const bool lower = (lane & stride) == 0;
                            const bool descending = (lane & size) == 0;      // this block sorts best-first
                            const bool keep_mine = (lower == descending) ? mine_first : !mine_first;
                            if (!keep_mine) { bk = ok; bp = op; }
                        }
                    }
                    if (c0 == 0) {
                        keep_key[0] = bk;
                        keep_pay[0] = bp;
                    } else {
                        // kept list is best-first in lanes 0..31; against the reversed chunk the lane-wise
                        // winners form a bitonic sequence holding the 32 best of the union
                        const uint64_t rk = __shfl_sync(kFull, bk, 31 - lane);
                        const uint32_t rp = __shfl_sync(kFull, bp, 31 - lane);
                        // kept entries are earlier candidates: they win ties
                        if (rk > keep_key[0]) { keep_key[0] = rk; keep_pay[0] = rp; }
                        #pragma unroll
                        for (int stride = 16; stride > 0; stride >>= 1) {
                            const uint64_t ok = __shfl_xor_sync(kFull, keep_key[0], stride);
                            const uint32_t op = __shfl_xor_sync(kFull, keep_pay[0], stride);
                            const bool mine_first = (keep_key[0] > ok) || (keep_key[0] == ok && keep_pay[0] < op);
                            const bool lower = (lane & stride) == 0;
                            if (lower != mine_first) { keep_key[0] = ok; keep_pay[0] = op; }
                        }
                    }
                    if (lane >= K) { keep_key[0] = 0; keep_pay[0] = 0; }
                } else {
                    // ---- beams of 33..64: the kept list is two sorted runs of 32 (ranks 0..31 in keep[0], 32..63 in
                    // keep[1]); a chunk is sorted by the same network as above, merged into the first run, and the 32
                    // entries that lose there are merged into the second.  One total order everywhere: larger key first,
                    // equal keys by payload = generation ordinal (beam.py:85 is a stable sort).
                    auto first_of = [](uint64_t ak, uint32_t ap, uint64_t bk2, uint32_t bp2) { return (ak > bk2) || (ak == bk2 && ap < bp2); };
                    // best-first order of a BITONIC sequence held one element per lane
                    auto clean = [&](uint64_t& k0, uint32_t& p0) {
                        #pragma unroll
                        for (int stride = 16; stride > 0; stride >>= 1) {
                            const uint64_t ok = __shfl_xor_sync(kFull, k0, stride);
                            const uint32_t op = __shfl_xor_sync(kFull, p0, stride);
                            const bool mine_first = first_of(k0, p0, ok, op);
                            const bool lower = (lane & stride) == 0;
                            if (lower != mine_first) { k0 = ok; p0 = op; }
                        }
                    };
                    bool skip = false;
                    if (c0 != 0 && K > 0) {
                        // a later chunk changes nothing unless one of its candidates beats the K-th kept entry
                        const int last = K - 1;
                        const uint64_t thr_k = __shfl_sync(kFull, last >= 32 ? keep_key[KR - 1] : keep_key[0], last & 31);
                        skip = thr_k != 0 && __ballot_sync(kFull, ckey > thr_k) == 0u;
                    }
                    if (!skip) {
                        uint64_t bk = ckey;
                        uint32_t bp = cpay;
                        #pragma unroll
                        for (int size = 2; size <= 32; size <<= 1) {
                            #pragma unroll
                            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                                const uint64_t ok = __shfl_xor_sync(kFull, bk, stride);
                                const uint32_t op = __shfl_xor_sync(kFull, bp, stride);
                                const bool mine_first = first_of(bk, bp, ok, op);
                                const bool lower = (lane & stride) == 0;
                                const bool descending = (lane & size) == 0;
                                const bool keep_mine = (lower == descending) ? mine_first : !mine_first;
                                if (!keep_mine) { bk = ok; bp = op; }
                            }
                        }
                        if (c0 == 0) {
                            keep_key[0] = bk;
                            keep_pay[0] = bp;
                            #pragma unroll
                            for (int r = 1; r < KR; ++r) { keep_key[r] = 0; keep_pay[r] = 0; }
                        } else {
                            // run 0 against the reversed chunk: lane-wise winners = the best 32 of both, losers = the rest
                            // (both bitonic); the losers then meet run 1 the same way
                            #pragma unroll
                            for (int r = 0; r < KR; ++r) {
                                const uint64_t rk = __shfl_sync(kFull, bk, 31 - lane);
                                const uint32_t rp = __shfl_sync(kFull, bp, 31 - lane);
                                const bool theirs = first_of(rk, rp, keep_key[r], keep_pay[r]);
                                bk = theirs ? keep_key[r] : rk;          // the loser of the pair moves on
                                bp = theirs ? keep_pay[r] : rp;
                                if (theirs) { keep_key[r] = rk; keep_pay[r] = rp; }
                                clean(keep_key[r], keep_pay[r]);
                                if (r + 1 < KR) clean(bk, bp);
                            }
                        }
                        #pragma unroll
                        for (int r = 0; r < KR; ++r)
                            if (r * 32 + lane >= K) { keep_key[r] = 0; keep_pay[r] = 0; }
                    }
                }
            }

            // ---- 5. survivors -> ring entries + trail ----
            int nl = 0, nn = 0;
            #pragma unroll
            for (int r = 0; r < KR; ++r) {
                const bool have = keep_key[r] != 0;
                uint32_t tag0 = LT_TAG_UNK;
                if (have) {
                    const int rank = r * 32 + lane;
                    const uint32_t kp = keep_pay[r];
                    const int j = LT_WINDOW - (int)((kp >> 27) & 0xFu);
                    const uint32_t prank = (kp >> 20) & 0x7Fu;
                    const uint32_t bidx = kp & 0xFFFFFu;
                    const bool unk_edge = (s_cnt[j] == 0);
                    const int pslot = ((e - j) % kRing) * K + (int)prank;
                    // the survivor's edge was prepared for this position: its word / morph products with M0
                    // are in the cache, and x * M0 -> x * M1 is one multiplication by M1 / M0
                    const bool cached = unk_edge || bidx < (uint32_t)kBucketCached;
                    H2 wk1, wk2, mk1;
                    uint32_t len, eref = kTrailUnk;
                    if (!unk_edge) eref = es + bidx;
                    if (cached) {
                        const uint32_t cslot = unk_edge ? unk_base + (uint32_t)j : (eref & (kEdgeRing - 1));
                        const H2 e0 = C.e0[cslot], g0 = unk_edge ? e0 : C.g0[cslot];
                        const uint32_t em = C.meta[cslot];
                        tag0 = em & 0xFFu;
                        len = (em >> 8) & 0xFFFFu;
                        wk1 = h2_mul(e0, kM1over0a, kM1over0b);
                        wk2 = h2_mul(e0, kM2over0a, kM2over0b);
                        mk1 = h2_mul(g0, kM1over0a, kM1over0b);
                    } else {
                        EdgeView k;
                        unpack_edge(ldg16(A.edges + eref), k);
                        edge_hashes<KB>(T, v, k, false);
                        tag0 = k.tag0;
                        len = k.len;
                        wk1 = h2_mul(k.wk, kM1a, kM1b);
                        wk2 = h2_mul(k.wk, kM2a, kM2b);
                        mk1 = h2_mul(k.mk, kM1a, kM1b);
                    }
                    const uint32_t pmeta = e_meta[pslot];
                    const uint32_t tj = pmeta & kMetaTagMask;
                    const bool k_ctx = (tag0 < 32) && ((kCtxMask >> tag0) & 1u);
                    const bool j_ctx = (tj < 32) && ((kCtxMask >> tj) & 1u);
                    const int dst = slot_e * K + rank;
                    e_score[dst] = unsortable(keep_key[r]);
                    e_p1[dst] = wk1;
                    e_pp[dst] = h2_add(wk1, h2_mul(e_p1[pslot], kM2over1a, kM2over1b));     // wk * M1 + wj * M2
                    e_c1[dst] = k_ctx ? mk1 : (j_ctx ? e_c1[pslot] : H2{0, 0});
                    const uint32_t ul = len < 8u ? len : 8u;
                    e_meta[dst] = tag0 | kMetaHasI | ((k_ctx || j_ctx) ? kMetaHasCtx : 0u) | (ul << kMetaUnkLenShift);
                    if (trail_smem) s_trail[(e - 1) * K + rank] = kp;
                    else A.trail[(size_t)(s0 + e - 1) * K + rank] = (uint64_t)eref | ((uint64_t)j << 32) | ((uint64_t)prank << 40);
                }
                // ranks of the entries that may be followed by an unknown word, ascending
                const unsigned m_have = __ballot_sync(kFull, have);
                const unsigned m_non = __ballot_sync(kFull, have && tag0 != LT_TAG_UNK);
                if (have && tag0 != LT_TAG_UNK) s_nonunk[slot_e * K + nn + __popc(m_non & lt_mask)] = (uint8_t)(r * 32 + lane);
                nl += __popc(m_have);
                nn += __popc(m_non);
            }
            if (lane == 0) { s_nbeam[slot_e] = (uint32_t)nl; s_nnon[slot_e] = (uint32_t)nn; s_acc[2] += (uint32_t)nl; }
            __syncwarp();
        }

        // ---- best path: matures[0] (tagger.py:78); with A.kbest every survivor (beam.py:59-61) ----
        // lane 0 follows the back-pointers (a dependent chain) and lists (end, edge reference); then the
        // lanes fetch the edge records side by side
        if constexpr (KB != 0) {
            const int nsurv = (L > 0) ? (int)s_nbeam[L % kRing] : 1;
            if (lane == 0) A.kb_count[s] = (st == LT_SENT_OK) ? nsurv : 0;
            for (int r = lane; r < K; r += 32) {
                A.kb_len[(size_t)s * K + r] = 0;
                A.kb_scores[(size_t)s * K + r] = 0.0;
            }
            __syncwarp();
        }
        if (L > 0) {
            uint64_t* s_path = ha;              // the prefix hashes are no longer needed
            const int nout = (KB != 0) ? (int)s_nbeam[L % kRing] : 1;
            for (int r0 = 0; r0 < nout; ++r0) {
                int W = 0;
                if (lane == 0) {
                    const double final_score = e_score[(L % kRing) * K + r0];
                    if (r0 == 0) A.scores[s] = final_score;
                    if constexpr (KB != 0) A.kb_scores[(size_t)s * K + r0] = final_score;
                    int e = L, r = r0;
                    while (e > 0) {
                        uint32_t eref, span;
                        if (trail_smem) {
                            const uint32_t t = s_trail[(e - 1) * K + r];
                            span = (uint32_t)LT_WINDOW - (t >> 27);
                            const uint32_t bidx = t & kPayUnk;
                            eref = (bidx == kPayUnk) ? kTrailUnk : spos[e - 1].x + bidx;
                            r = (int)((t >> 20) & 0x7Fu);
                        } else {
                            const uint64_t t = A.trail[(size_t)(s0 + e - 1) * K + r];
                            eref = (uint32_t)t;
                            span = (uint32_t)((t >> 32) & 0xFFu);
                            r = (int)((t >> 40) & 0xFFu);
                        }
                        s_path[W] = (uint64_t)eref | ((uint64_t)e << 32) | ((uint64_t)span << 48);
                        ++W;
                        e -= (int)span;
                    }
                    if (r0 == 0) {
                        A.path_len[s] = W;
                        s_acc[3] += (uint32_t)W;
                    }
                    if constexpr (KB != 0) A.kb_len[(size_t)s * K + r0] = W;
                }
                W = __shfl_sync(kFull, W, 0);
                __syncwarp();
                for (int w = lane; w < W; w += 32) {
                    const uint64_t t = s_path[w];
                    const uint32_t eref = (uint32_t)t;
                    lt_edge ed;
                    if (eref == kTrailUnk) {
                        const uint32_t e = (uint32_t)(t >> 32) & 0xFFFFu, span = (uint32_t)(t >> 48);
                        ed.b = (uint16_t)(e - span); ed.e = (uint16_t)e; ed.len = (uint16_t)span;
                        ed.tag0 = LT_TAG_UNK; ed.tag1 = LT_NO_TAG; ed.rule = LT_NO_RULE; ed.split = 0;
                        ed.flags = LT_EDGE_UNK; ed.reserved = 0;
                    } else {
                        ed = A.edges[eref];
                    }
                    if (r0 == 0) A.path_tmp[s0 + w] = ed;
                    if constexpr (KB != 0) A.kb_tmp[(size_t)s0 * K + (size_t)r0 * (s1 - s0) + w] = ed;
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    // counters (per warp and launch each stays far below 2^32)
    __syncwarp();
    if (lane < 4) atomicAdd(A.counters + 3 + lane, (unsigned long long)s_acc[lane]);
}

}  // namespace lt
