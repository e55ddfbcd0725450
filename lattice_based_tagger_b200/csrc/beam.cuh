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
//   1. reads the CSR bucket of e (sorted by begin) and counts edges per span;
//   2. EDGE PREP, lanes = edges: per edge the word/morpheme hash products and everything of the
//      score that depends on the edge alone — RegularizationScore, the preference scorers,
//      templates 4 and 5 — into a shared-memory edge cache;
//   3. enumerates candidates in the reference's generation order — begin ascending, parent rank
//      ascending, edge order ascending (beam.py:30-48) — 32 at a time, one per lane, and scores
//      each: score program in BeamScoreFunctions order, every fp64 add in the reference's
//      association; templates 0,1,2,7,8 are gathered from the HBM feature table
//      with all first-slot loads in flight together, template 3/6 come from shared memory;
//   4. keeps the best `beam` candidates: repeated warp arg-max (redux.sync on the order-preserving
//      integer image of the fp64 score) with ties resolved towards the earlier candidate, which is
//      exactly the stable sort of Beam.append (beam.py:83-86);
//   5. writes the survivors as new ring entries and one 8-byte back-pointer each to the HBM trail.
// The best path is recovered from the trail and written as 16-byte edge records.
#pragma once
#include "lattice.cuh"
#include "tables.cuh"

namespace lt {

constexpr int kRing = LT_WINDOW + 1;
constexpr int kECache = 24;                   // cached window edges per position (+ 8 unknown spans)
constexpr int kECacheAll = kECache + LT_WINDOW;
constexpr int kRoundsMaxBeam = 6;             // beams above this use the sorting-network top-K (one sentence per warp)
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
    int32_t lcap;
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
};

// per-edge cache entry (shared memory, struct of arrays)
struct EdgeCache {
    H2* e0;            // word hash   * M0
    H2* g0;            // morph0 hash * M0
    double* kval;      // kRingVals doubles per edge: score-program values that depend on the edge only
    uint32_t* meta;    // tag0 | len << 8 (16 bits) | flags << 24
    uint32_t* eref;    // global edge index (kTrailUnk for unknown words)
    uint32_t* present; // bit f*2: template 4 present, bit f*2+1: template 5 present (per scorer f)
};

// doubles per edge in kval (stride 2 * n_funcs): for scorer f: [2f] = REG / MPREF / WPREF value or
// template-4 weight, [2f+1] = template-5 weight

__host__ __device__ inline size_t beam_warp_smem(int lcap, int beam, int n_funcs) {
    size_t units = (size_t)lcap + 8;
    size_t bytes = units * 8 * 2;                          // ha, hb
    bytes += units * 8;                                    // pos (uint2)
    bytes += (size_t)kRing * beam * (8 + 64);              // score, p1, pp, j2, c1
    bytes += (size_t)kECacheAll * (16 + 16 + 16 * (n_funcs > 0 ? n_funcs : 1));   // e0, g0, kval
    bytes += units * 2;                                    // chars
    bytes = (bytes + 7) & ~(size_t)7;
    bytes += (size_t)kRing * beam * 4;                     // meta
    bytes += (size_t)kECacheAll * 12;                      // cache meta, eref, present
    bytes += 64 * 4;                                       // per-span tables + ring sizes
    return (bytes + 15) & ~(size_t)15;
}

// trail entry: edge reference (global edge index, or kTrailUnk) | span << 32 | parent rank << 40
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

__device__ __noinline__ void edge_hashes(const DevTables& T, const SentView& v, EdgeView& k, bool need_m1) {
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
__device__ __noinline__ uint32_t edge_score(const DevTables& T, const unsigned char* dense_smem, const EdgeView& k,
                                            H2 e0, H2 g0, int f, double& a, double& b2) {
    uint32_t present = 0;
    const uint32_t tk = k.tag0;
    const lt_func& fn = T.funcs[f];
    a = 0.0;
    b2 = 0.0;
    if (fn.kind == LT_FUNC_REG) {
        // score_funcs.py:65-73
        if (tk == LT_TAG_UNK) a = __dmul_rn(fn.p[0], __dadd_rn((double)k.len, 0.1));
        else a = __dmul_rn(fn.p[1], (double)k.len);
        a = __dadd_rn(0.0, a);
        if (k.len == 1 && tk == LT_TAG_NOUN) a = __dadd_rn(a, fn.p[2]);
    } else if (fn.kind == LT_FUNC_MPREF) {
        // score_funcs.py:84-88
        FKey k0 = feature_key_sum32(T.seeds[f][9], feature_head32(tk, 0), g0);
        FeatProbe s0 = feat_first(T, k0);
        if (k.tag1 != LT_NO_TAG) {
            FKey k1 = feature_key_sum32(T.seeds[f][9], feature_head32(k.tag1, 0), h2_mul(k.m1, kM0a, kM0b));
            FeatProbe s1 = feat_first(T, k1);
            feat_resolve(T, k1, s1, b2);
        }
        feat_resolve(T, k0, s0, a);
        if (k.tag1 != LT_NO_TAG) a = __dadd_rn(a, b2);
        b2 = 0.0;
    } else if (fn.kind == LT_FUNC_WPREF) {
        // score_funcs.py:99-100
        FKey k0 = feature_key_sum32(T.seeds[f][9], feature_head32(tk, 0), e0);
        FeatProbe s0 = feat_first(T, k0);
        feat_resolve(T, k0, s0, a);
    } else {
        // templates 4 (wk.len) and 5 (wk.word, wk.tag0, wk.is_l), features/feature.py:100,104
        const DenseView D = dense_view(dense_smem + (size_t)T.func_dense[f] * dense_block_bytes(T.n_tags), T.n_tags);
        FKey q5 = feature_key_sum32(T.seeds[f][5], feature_head32(tk, (k.flags & LT_EDGE_IS_L) ? 1u : 0u), e0);
        FeatProbe s5 = feat_first(T, q5);
        if (k.len >= (uint32_t)kT4Dense) {
            FKey q4 = feature_key_sum(T.seeds[f][4], feature_head(k.len, 0), H2{0, 0});
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

// numpy's association from eight surviving weights on: rare, so the nine weights are
// simply gathered again and summed by numpy_order_sum9.
__device__ __noinline__ double trigram_sum_tree(const DevTables& T, const DenseView& D, int NT, int f, FKey q0, FKey q1,
                                                FKey q2, FKey q7, FKey q8, uint32_t tj, uint32_t tk, uint32_t epresent,
                                                double val4, double val5, bool j_unk, uint32_t ul, bool has_i, bool ctx8) {
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
#define LT_BEAM_MINB 4
#endif
constexpr int kBeamWarps = 4;                 // warps per CTA of the beam kernel

// Prefix hashes by a scan over a GROUP of G lanes: H[0] = 0, H[i+1] = H[i] * B + (c_i + 1).
template <int G>
__device__ __forceinline__ void prefix_hashes_group(const uint16_t* ch, int L, int Lmax, int gl, unsigned gmask,
                                                    uint64_t* ha, uint64_t* hb) {
    H2 carry{0, 0};
    if (gl == 0) {
        ha[0] = 0;
        hb[0] = 0;
    }
    constexpr uint64_t A1 = kBaseA, A2 = A1 * A1, A4 = A2 * A2, A8 = A4 * A4, A16 = A8 * A8;
    constexpr uint64_t B1 = kBaseB, B2 = B1 * B1, B4 = B2 * B2, B8 = B4 * B4, B16 = B8 * B8;
    for (int base = 0; base < Lmax; base += G) {
        const int i = base + gl;
        const uint64_t v = (i < L) ? (uint64_t)ch[i] + 1u : 0u;
        uint64_t sa = v, sb = v, ta, tb;
        ta = __shfl_up_sync(0xFFFFFFFFu, sa, 1, G);  tb = __shfl_up_sync(0xFFFFFFFFu, sb, 1, G);
        if (gl >= 1)  { sa += ta * A1;  sb += tb * B1; }
        ta = __shfl_up_sync(0xFFFFFFFFu, sa, 2, G);  tb = __shfl_up_sync(0xFFFFFFFFu, sb, 2, G);
        if (gl >= 2)  { sa += ta * A2;  sb += tb * B2; }
        ta = __shfl_up_sync(0xFFFFFFFFu, sa, 4, G);  tb = __shfl_up_sync(0xFFFFFFFFu, sb, 4, G);
        if (gl >= 4)  { sa += ta * A4;  sb += tb * B4; }
        ta = __shfl_up_sync(0xFFFFFFFFu, sa, 8, G);  tb = __shfl_up_sync(0xFFFFFFFFu, sb, 8, G);
        if (gl >= 8)  { sa += ta * A8;  sb += tb * B8; }
        if (G == 32) {
            ta = __shfl_up_sync(0xFFFFFFFFu, sa, 16, G); tb = __shfl_up_sync(0xFFFFFFFFu, sb, 16, G);
            if (gl >= 16) { sa += ta * A16; sb += tb * B16; }
        }
        uint64_t pa = 1, pb = 1;   // B^(gl+1)
        {
            uint64_t xa = A1, xb = B1;
            const int k = gl + 1;
            #pragma unroll
            for (int bit = 0; bit < 6; ++bit) {
                if (k & (1 << bit)) { pa *= xa; pb *= xb; }
                xa *= xa; xb *= xb;
            }
        }
        const uint64_t outa = carry.a * pa + sa, outb = carry.b * pb + sb;
        if (i < L) {
            ha[i + 1] = outa;
            hb[i + 1] = outb;
        }
        carry.a = __shfl_sync(0xFFFFFFFFu, outa, G - 1, G);
        carry.b = __shfl_sync(0xFFFFFFFFu, outb, G - 1, G);
    }
    __syncwarp();
}

// Maximum of a group-uniform value over the groups of a warp, so that both groups run the same
// number of loop iterations and stay converged (lock-step) — otherwise two half-warp groups would
// simply serialise.
template <int G>
__device__ __forceinline__ uint32_t warp_max(uint32_t x) {
    if (G == 32) return x;
    const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, 16);
    return x > y ? x : y;
}

// Collectives always name the full warp: a *_sync with a half-warp mask makes the hardware treat
// the two halves as separate convergence groups, which then run one after the other.
template <int G>
__device__ __forceinline__ uint32_t group_max(uint32_t x, int lane) {
    if (G == 32) return __reduce_max_sync(0xFFFFFFFFu, x);
    const bool hi_half = lane >= 16;
    const uint32_t m0 = __reduce_max_sync(0xFFFFFFFFu, hi_half ? 0u : x);
    const uint32_t m1 = __reduce_max_sync(0xFFFFFFFFu, hi_half ? x : 0u);
    return hi_half ? m1 : m0;
}
// ballot restricted to the caller's group, bits at absolute lane positions
#define LT_GBALLOT(pred) (__ballot_sync(0xFFFFFFFFu, (pred)) & gmask)

// KR = ceil(beam / G) kept entries per lane; G = lanes per sentence (32: one sentence per warp,
// 16: two sentences per warp — small beams generate about 16 candidates per position, so half a
// warp per sentence doubles the useful lanes of every phase).
template <int KR, int G, bool SORTNET>   // SORTNET: sorting-network top-K (G == 32, KR == 1), else arg-max rounds
__global__ void __launch_bounds__(kBeamWarps * 32, LT_BEAM_MINB) beam_kernel(const __grid_constant__ DevTables T, const __grid_constant__ BeamArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);                       // lane within the sentence's group
    const int gshift = lane & ~(G - 1);                  // first lane of the group
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << gshift);
    const int group = (threadIdx.x >> 5) * (32 / G) + (lane / G);
    const int K = A.beam;
    const int NT = T.n_tags;

    // CTA-shared dense tables (tag x tag matrix, length vectors)
    const size_t dense_bytes = ((size_t)T.n_tri * dense_block_bytes(NT) + 15) & ~(size_t)15;
    unsigned char* dense_smem = smem_raw;
    for (size_t i = threadIdx.x * 4; i < (size_t)T.n_tri * dense_block_bytes(NT); i += blockDim.x * 4)
        *reinterpret_cast<uint32_t*>(dense_smem + i) = *reinterpret_cast<const uint32_t*>(T.dense + i);
    __syncthreads();
    if (A.flags[kFlagEdgeOverflow] | A.flags[kFlagStageOverflow]) return;   // lattice incomplete: the host grows the buffer and reruns

    const size_t units = (size_t)A.lcap + 8;
    unsigned char* wbase = smem_raw + dense_bytes + (size_t)group * beam_warp_smem(A.lcap, K, T.n_funcs);
    uint64_t* ha = reinterpret_cast<uint64_t*>(wbase);
    uint64_t* hb = ha + units;
    uint2* spos = reinterpret_cast<uint2*>(hb + units);
    double* e_score = reinterpret_cast<double*>(spos + units);
    H2* e_p1 = reinterpret_cast<H2*>(e_score + kRing * K);     // wj * M1
    H2* e_pp = e_p1 + kRing * K;                                // wj * M1 + wi * M2
    H2* e_j2 = e_pp + kRing * K;                                // wj * M2
    H2* e_c1 = e_j2 + kRing * K;                                // contextual morph * M1
    EdgeCache C;
    C.e0 = e_c1 + kRing * K;
    C.g0 = C.e0 + kECacheAll;
    C.kval = reinterpret_cast<double*>(C.g0 + kECacheAll);
    const int kvs = 2 * (T.n_funcs > 0 ? T.n_funcs : 1);   // kval stride
    uint16_t* ch = reinterpret_cast<uint16_t*>(C.kval + (size_t)kECacheAll * kvs);
    uintptr_t after = (reinterpret_cast<uintptr_t>(ch + units) + 7) & ~(uintptr_t)7;
    uint32_t* e_meta = reinterpret_cast<uint32_t*>(after);
    C.meta = e_meta + kRing * K;
    C.eref = C.meta + kECacheAll;
    C.present = C.eref + kECacheAll;
    uint32_t* s_cnt = C.present + kECacheAll;       // [9] edges per span
    uint32_t* s_gstart = s_cnt + 16;                // [9] first bucket-local index of the span's group
    uint32_t* s_ncand = s_gstart + 16;              // [9] candidates of the span
    uint32_t* s_nbeam = s_ncand + 16;               // [kRing] entries per ring slot

    unsigned long long acc_T = 0, acc_F = 0, acc_B = 0, acc_W = 0;

    bool need_m1 = false;
    for (int f = 0; f < T.n_funcs; ++f) need_m1 |= (T.funcs[f].kind == LT_FUNC_MPREF);
    const int nf = T.n_funcs;

    while (true) {
        // the warp takes one sentence per group; a group without a (taggable) sentence idles with L = 0
        unsigned int s = 0;
        if (lane == 0) s = atomicAdd(A.queue, (unsigned)(32 / G));
        s = __shfl_sync(kFull, s, 0);
        if (s >= (unsigned)A.n_sent) break;
        s += (unsigned)(lane / G);
        const bool have = s < (unsigned)A.n_sent;
        if (have && A.order) s = __ldg(A.order + s);
        const int s0 = have ? __ldg(A.sent_off + s) : 0, s1 = have ? __ldg(A.sent_off + s + 1) : 0;
        const int st = have ? __ldg(A.status + s) : LT_SENT_BAD_SPACE;
        // ---- stage syllables, prefix hashes and the sentence's CSR row ----
        int L = 0;
        const int raw_len = (st == LT_SENT_OK) ? (s1 - s0) : 0;
        const int raw_max = (int)warp_max<G>((uint32_t)raw_len);
        for (int base = 0; base < raw_max; base += G) {
            const int idx = s0 + base + gl;
            const bool valid = base + gl < raw_len;
            const uint32_t c = valid ? (uint32_t)__ldg(A.text + idx) : 0x20u;
            const bool keep = valid && (c != 0x20u);
            const unsigned km = LT_GBALLOT(keep) >> gshift;
            const int pos = L + __popc(km & ((1u << gl) - 1u));
            if (keep) ch[pos] = (uint16_t)c;
            L += __popc(km);
        }
        __syncwarp();
        for (int i = gl; i < L; i += G) spos[i] = __ldg(A.pos + s0 + i);
        const int Lmax = (int)warp_max<G>((uint32_t)L);
        prefix_hashes_group<G>(ch, L, Lmax, gl, gmask, ha, hb);
        SentView v{ch, ha, hb, nullptr};

        if (L == 0 && have && gl == 0) { A.path_len[s] = 0; A.scores[s] = 0.0; }

        // beam[0] = [BOS] (beam.py:21-23)
        if (gl == 0) {
            e_score[0] = 0.0;
            e_p1[0] = h2_mul(T.bos, kM1a, kM1b);
            e_j2[0] = h2_mul(T.bos, kM2a, kM2b);
            e_pp[0] = H2{0, 0};
            e_c1[0] = H2{0, 0};
            e_meta[0] = (uint32_t)LT_TAG_BOS;
            s_nbeam[0] = 1;
        }
        __syncwarp();

        for (int e = 1; e <= Lmax; ++e) {
            const bool on = e <= L;                     // this group's sentence still has positions
            const int slot_e = e % kRing;
            const uint2 bucket = on ? spos[e - 1] : make_uint2(0u, 0u);
            const uint32_t es = bucket.x, ne = bucket.y;
            const int jmax = (e < LT_WINDOW) ? e : LT_WINDOW;
            const uint32_t ne_max = warp_max<G>(ne);

            // ---- 1. edges per span (bucket sorted by begin ascending = span descending) ----
            uint32_t my_cnt = 0;            // lane j (1..8) counts span j
            uint4 raw0 = make_uint4(0, 0, 0, 0);
            for (uint32_t base = 0; base < ne_max; base += G) {
                const uint32_t idx = base + gl;
                int span = 0;
                if (idx < ne) {
                    const uint4 raw = ldg16(A.edges + es + idx);
                    if (base == 0) raw0 = raw;
                    span = (int)(raw.x >> 16) - (int)(raw.x & 0xFFFFu);
                }
                #pragma unroll
                for (int j = 1; j <= LT_WINDOW; ++j) {
                    const uint32_t c = __popc(LT_GBALLOT(span == j));
                    if (gl == j) my_cnt += c;
                }
            }
            // lane j: group start = ne - sum_{j' <= j} cnt[j']: spans descend along the bucket
            {
                const uint32_t c = (gl >= 1 && gl <= LT_WINDOW) ? my_cnt : 0u;
                uint32_t incl = c;      // inclusive prefix over lanes 1..j
                #pragma unroll
                for (int d = 1; d < 16; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, incl, d, G);
                    if (gl >= d) incl += t;
                }
                if (gl >= 1 && gl <= LT_WINDOW) {
                    s_cnt[gl] = c;
                    s_gstart[gl] = ne - incl;                 // bucket-local index of the group's first edge
                    const uint32_t np = (on && gl <= jmax) ? s_nbeam[(e - gl) % kRing] : 0u;
                    s_ncand[gl] = np * (c ? c : 1u);
                }
            }
            __syncwarp();
            const uint32_t in_window = ne - s_gstart[LT_WINDOW];    // = sum of cnt[1..8]
            const uint32_t first_in = ne - in_window;               // bucket-local index of the first window edge
            uint32_t N = 0;
            #pragma unroll
            for (int j = 1; j <= LT_WINDOW; ++j) N += s_ncand[j];

            // ---- 2. edge prep into the cache: window edges, then the unknown word of every empty span ----
            {
                const uint32_t n_cached = in_window < (uint32_t)kECache ? in_window : (uint32_t)kECache;
                const uint32_t prep_max = warp_max<G>(n_cached + LT_WINDOW);
                for (uint32_t base = 0; base < prep_max; base += G) {
                    const uint32_t ci = base + gl;
                    bool active = false;
                    EdgeView k;
                    uint32_t slot = 0, eref = kTrailUnk;
                    // the first G bucket entries are still in the registers of lane (bucket index)
                    const uint32_t bidx = first_in + ci;
                    uint4 r0;
                    r0.x = __shfl_sync(kFull, raw0.x, bidx & (G - 1), G);
                    r0.y = __shfl_sync(kFull, raw0.y, bidx & (G - 1), G);
                    r0.z = __shfl_sync(kFull, raw0.z, bidx & (G - 1), G);
                    r0.w = __shfl_sync(kFull, raw0.w, bidx & (G - 1), G);
                    if (ci < n_cached) {
                        unpack_edge((bidx < (uint32_t)G) ? r0 : ldg16(A.edges + es + bidx), k);
                        slot = ci;
                        eref = es + bidx;
                        active = true;
                    } else {
                        const int j = (int)(ci - n_cached) + 1;
                        if (on && j <= jmax && s_cnt[j] == 0) {
                            unknown_edge(e - j, e, k);
                            slot = (uint32_t)(kECache + j - 1);
                            active = true;
                        }
                    }
                    if (active) {
                        edge_hashes(T, v, k, need_m1);
                        const H2 e0 = h2_mul(k.wk, kM0a, kM0b), g0 = h2_mul(k.mk, kM0a, kM0b);
                        uint32_t present = 0;
                        #pragma unroll 1
                        for (int f = 0; f < nf; ++f) {
                            double a, b2;
                            present |= edge_score(T, dense_smem, k, e0, g0, f, a, b2) << (2 * f);
                            C.kval[slot * kvs + 2 * f] = a;
                            C.kval[slot * kvs + 2 * f + 1] = b2;
                        }
                        C.e0[slot] = e0;
                        C.g0[slot] = g0;
                        C.meta[slot] = k.tag0 | (k.len << 8) | (k.flags << 24);
                        C.eref[slot] = eref;
                        C.present[slot] = present;
                    }
                }
            }
            __syncwarp();

            // ---- 3 + 4. candidates in generation order, running top-K ----
            uint64_t keep_key[KR];
            uint32_t keep_pay[KR];
            #pragma unroll
            for (int r = 0; r < KR; ++r) { keep_key[r] = 0; keep_pay[r] = 0; }

            const uint32_t N_max = warp_max<G>(N);
            for (uint32_t c0 = 0; c0 < N_max; c0 += G) {
                const uint32_t c = c0 + gl;
                bool valid = c < N;
                int j = 0;
                uint32_t rem = c;
                if (valid) {
                    #pragma unroll
                    for (int jj = LT_WINDOW; jj >= 1; --jj) {
                        if (j == 0) {
                            const uint32_t nc = s_ncand[jj];
                            if (rem < nc) j = jj; else rem -= nc;
                        }
                    }
                }
                uint64_t ckey = 0;
                uint32_t cpay = 0;
                if (valid) {
                    const uint32_t cj = s_cnt[j];
                    const bool unk_edge = (cj == 0);
                    const uint32_t nedge = unk_edge ? 1u : cj;
                    const uint32_t prank = (nedge == 1u) ? rem : rem / nedge, eidx = rem - prank * nedge;
                    const int pslot = ((e - j) % kRing) * K + (int)prank;
                    const uint32_t pmeta = e_meta[pslot];
                    const uint32_t tj = pmeta & kMetaTagMask;
                    // cache slot of the edge
                    const uint32_t widx = s_gstart[j] - first_in + eidx;     // index among window edges
                    const uint32_t slot = unk_edge ? (uint32_t)(kECache + j - 1) : widx;
                    H2 e0, g0;
                    uint32_t emeta, epresent = 0;
                    const bool uncached = !unk_edge && widx >= (uint32_t)kECache;
                    EdgeView kfly;
                    if (uncached) {
                        // bucket larger than the cache: prepare this edge on the fly
                        unpack_edge(ldg16(A.edges + es + s_gstart[j] + eidx), kfly);
                        edge_hashes(T, v, kfly, need_m1);
                        e0 = h2_mul(kfly.wk, kM0a, kM0b);
                        g0 = h2_mul(kfly.mk, kM0a, kM0b);
                        emeta = kfly.tag0 | (kfly.len << 8) | (kfly.flags << 24);
                    } else {
                        e0 = C.e0[slot];
                        g0 = C.g0[slot];
                        emeta = C.meta[slot];
                        epresent = C.present[slot];
                    }
                    const uint32_t tk = emeta & 0xFFu;
                    // two unknown words in a row are only allowed from the window's first begin (beam.py:44-45)
                    if (tj == LT_TAG_UNK && tk == LT_TAG_UNK && j < jmax) {
                        valid = false;
                    } else {
                        const double pscore = e_score[pslot];
                        const H2 p1 = e_p1[pslot];
                        const bool has_i = (pmeta & kMetaHasI) != 0;
                        const bool j_unk = (tj == LT_TAG_UNK);
                        const bool ctx8 = ((kCtxMask >> tk) & 1u) && (pmeta & kMetaHasCtx);
                        double inc = 0.0;
                        #pragma unroll 1
                        for (int f = 0; f < nf; ++f) {
                            double val, val5;
                            if (uncached) {
                                epresent |= edge_score(T, dense_smem, kfly, e0, g0, f, val, val5) << (2 * f);
                            } else {
                                val = C.kval[slot * kvs + 2 * f];
                                val5 = C.kval[slot * kvs + 2 * f + 1];
                            }
                            if (T.funcs[f].kind == LT_FUNC_TRIGRAM) {
                                // SimpleTrigramFeatureScore.score (score_funcs.py:137-144)
                                const DenseView D = dense_view(dense_smem + (size_t)T.func_dense[f] * dense_block_bytes(NT), NT);
                                acc_F += 6u + (j_unk ? 1u : 0u) + (has_i ? 1u : 0u) + (ctx8 ? 1u : 0u);
                                const H2 pp = has_i ? e_pp[pslot] : H2{0, 0};
                                const H2 c1v = ctx8 ? e_c1[pslot] : H2{0, 0};
                                const uint32_t hk = feature_head32(tk, 0);
                                const FKey q0 = feature_key_sum32(T.seeds[f][0], hk, h2_add(e0, p1));
                                const FKey q1 = feature_key_sum32(T.seeds[f][1], hk, p1);
                                const FKey q2 = feature_key_sum32(T.seeds[f][2], feature_head32(tj, tk), e0);
                                const FKey q7 = feature_key_sum32(T.seeds[f][7], 0u, h2_add(e0, pp));
                                const FKey q8 = feature_key_sum32(T.seeds[f][8], 0u, h2_add(g0, c1v));
                                // all first-slot loads in flight before any is consumed
                                const FeatProbe s0 = feat_first(T, q0);
                                const FeatProbe s1 = feat_first(T, q1);
                                const FeatProbe s2 = feat_first(T, q2);
                                FeatProbe s7, s8;
                                if (has_i) s7 = feat_first(T, q7);
                                if (ctx8) s8 = feat_first(T, q8);
                                // running left-to-right sum = numpy's order while fewer than 8 weights survive
                                double acc = 0.0, w;
                                int n = 0;
                                if (feat_resolve(T, q0, s0, w)) { acc = __dadd_rn(acc, w); ++n; }
                                if (feat_resolve(T, q1, s1, w)) { acc = __dadd_rn(acc, w); ++n; }
                                if (feat_resolve(T, q2, s2, w)) { acc = __dadd_rn(acc, w); ++n; }
                                if ((D.m3[tj] >> tk) & 1u) { acc = __dadd_rn(acc, D.t3[tj * NT + tk]); ++n; }
                                if ((epresent >> (2 * f)) & 1u) { acc = __dadd_rn(acc, val); ++n; }
                                if ((epresent >> (2 * f + 1)) & 1u) { acc = __dadd_rn(acc, val5); ++n; }
                                const uint32_t ul = (pmeta >> kMetaUnkLenShift) & 0xFu;
                                if (j_unk && ((D.m6[0] >> ul) & 1u)) { acc = __dadd_rn(acc, D.t6[ul]); ++n; }
                                if (has_i && feat_resolve(T, q7, s7, w)) { acc = __dadd_rn(acc, w); ++n; }
                                if (ctx8 && feat_resolve(T, q8, s8, w)) { acc = __dadd_rn(acc, w); ++n; }
                                if (n >= 8)   // numpy switches to an 8-lane tree: redo the gather and add in that order (rare)
                                    acc = trigram_sum_tree(T, D, NT, f, q0, q1, q2, q7, q8, tj, tk, epresent, val, val5, j_unk, ul,
                                                           has_i, ctx8);
                                val = n ? acc : 0.0;
                            }
                            inc = __dadd_rn(inc, val);          // score += f(...), score_funcs.py:51-53
                        }
                        double newscore = __dadd_rn(pscore, inc);        // Sequence.add, beam.py:115
                        newscore = __dadd_rn(newscore, 0.0);             // -0.0 sorts as 0.0
                        ckey = sortable(newscore);
                        // payload doubles as the generation ordinal: begin ascending (= span descending), parent
                        // rank ascending, edge order ascending (beam.py:30-48)
                        cpay = ((uint32_t)(LT_WINDOW - j) << 27) | (prank << 20) | (unk_edge ? 0u : s_gstart[j] + eidx);
                        acc_T += 1;
                    }
                }
                if constexpr (SORTNET) {
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
                // ---- top-K of (kept so far) U (this chunk): K rounds of group arg-max ----
                // Priority on equal keys: kept entries (earlier candidates) by rank, then chunk lanes in order.
                uint64_t new_key[KR];
                uint32_t new_pay[KR];
                #pragma unroll
                for (int r = 0; r < KR; ++r) { new_key[r] = 0; new_pay[r] = 0; }
                uint64_t pool_key[KR + 1];
                #pragma unroll
                for (int r = 0; r < KR; ++r) pool_key[r] = keep_key[r];
                pool_key[KR] = ckey;
                // rounds needed: min(K, pool size), the same for both groups of the warp
                uint32_t pool_n = __popc(LT_GBALLOT(ckey != 0));
                #pragma unroll
                for (int r = 0; r < KR; ++r) pool_n += __popc(LT_GBALLOT(keep_key[r] != 0));
                const int rounds = (int)warp_max<G>(pool_n < (uint32_t)K ? pool_n : (uint32_t)K);
                for (int round = 0; round < rounds; ++round) {
                    // lane-local best: lower pool index wins ties
                    uint64_t best = pool_key[0];
                    int cls = 0;
                    #pragma unroll
                    for (int r = 1; r <= KR; ++r)
                        if (pool_key[r] > best) { best = pool_key[r]; cls = r; }
                    const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
                    const uint32_t mhi = group_max<G>(hi, lane);
                    const bool c1 = (hi == mhi) && (mhi != 0);              // mhi == 0: this group's pool is exhausted
                    const uint32_t mlo = group_max<G>(c1 ? lo : 0u, lane);
                    const bool c2 = c1 && (lo == mlo);
                    int win_cls = 0;
                    unsigned wm = 0;
                    #pragma unroll
                    for (int r = 0; r <= KR; ++r) {
                        const unsigned m = LT_GBALLOT(c2 && cls == r);
                        if (wm == 0 && m != 0) { wm = m; win_cls = r; }
                    }
                    const bool sel = wm != 0;                               // false: this group's pool is exhausted
                    const int src = sel ? __ffs(wm) - 1 : lane;             // absolute lane of the winner
                    uint32_t pay_mine = cpay;
                    #pragma unroll
                    for (int r = 0; r < KR; ++r)
                        if (win_cls == r) pay_mine = keep_pay[r];
                    const uint32_t wpay = __shfl_sync(kFull, pay_mine, src, G);
                    const uint64_t wkey = ((uint64_t)mhi << 32) | mlo;
                    if (sel && gl == (round % G)) {
                        #pragma unroll
                        for (int r = 0; r < KR; ++r)
                            if ((round / G) == r) { new_key[r] = wkey; new_pay[r] = wpay; }
                    }
                    if (sel && lane == src) {
                        #pragma unroll
                        for (int r = 0; r <= KR; ++r)
                            if (win_cls == r) pool_key[r] = 0;
                    }
                }
                #pragma unroll
                for (int r = 0; r < KR; ++r) { keep_key[r] = new_key[r]; keep_pay[r] = new_pay[r]; }
                }
            }

            // ---- 5. survivors -> ring entries + trail ----
            int nl = 0;
            #pragma unroll
            for (int r = 0; r < KR; ++r) {
                const unsigned m = LT_GBALLOT(keep_key[r] != 0);
                nl += __popc(m);
                if (keep_key[r] != 0) {
                    const int rank = r * G + gl;
                    const uint32_t kp = keep_pay[r];
                    const int j = LT_WINDOW - (int)((kp >> 27) & 0xFu);
                    const uint32_t prank = (kp >> 20) & 0x7Fu;
                    const bool unk_edge = (s_cnt[j] == 0);
                    const int pslot = ((e - j) % kRing) * K + (int)prank;
                    // the survivor's edge was prepared for this position: its word / morph products with M0
                    // are in the cache, and x * M0 -> x * M1 is one multiplication by M1 / M0
                    const uint32_t widx = (kp & 0xFFFFFu) - first_in;
                    const bool cached = unk_edge || widx < (uint32_t)kECache;
                    H2 wk1, wk2, mk1;
                    uint32_t tag0, len, eref = kTrailUnk;
                    if (cached) {
                        const uint32_t cslot = unk_edge ? (uint32_t)(kECache + j - 1) : widx;
                        const H2 e0 = C.e0[cslot], g0 = C.g0[cslot];
                        const uint32_t em = C.meta[cslot];
                        tag0 = em & 0xFFu;
                        len = (em >> 8) & 0xFFFFu;
                        eref = C.eref[cslot];
                        wk1 = h2_mul(e0, kM1over0a, kM1over0b);
                        wk2 = h2_mul(e0, kM2over0a, kM2over0b);
                        mk1 = h2_mul(g0, kM1over0a, kM1over0b);
                    } else {
                        EdgeView k;
                        eref = es + (kp & 0xFFFFFu);
                        unpack_edge(ldg16(A.edges + eref), k);
                        edge_hashes(T, v, k, false);
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
                    e_pp[dst] = h2_add(wk1, e_j2[pslot]);
                    e_j2[dst] = wk2;
                    e_c1[dst] = k_ctx ? mk1 : (j_ctx ? e_c1[pslot] : H2{0, 0});
                    const uint32_t ul = len < 8u ? len : 8u;
                    e_meta[dst] = tag0 | kMetaHasI | ((k_ctx || j_ctx) ? kMetaHasCtx : 0u) | (ul << kMetaUnkLenShift);
                    A.trail[(size_t)(s0 + e - 1) * K + rank] =
                        (uint64_t)eref | ((uint64_t)j << 32) | ((uint64_t)prank << 40);
                }
            }
            if (gl == 0) s_nbeam[slot_e] = (uint32_t)nl;
            acc_B += (gl == 0) ? (unsigned long long)nl : 0ull;
            __syncwarp();
        }

        // ---- best path: matures[0] (tagger.py:78) ----
        if (gl == 0 && L > 0) {
            A.scores[s] = e_score[(L % kRing) * K + 0];
            int e = L, r = 0, W = 0;
            while (e > 0) {
                const uint64_t t = A.trail[(size_t)(s0 + e - 1) * K + r];
                const uint32_t eref = (uint32_t)t;
                const int span = (int)((t >> 32) & 0xFFu);
                lt_edge ed;
                if (eref == kTrailUnk) {
                    ed.b = (uint16_t)(e - span); ed.e = (uint16_t)e; ed.len = (uint16_t)span;
                    ed.tag0 = LT_TAG_UNK; ed.tag1 = LT_NO_TAG; ed.rule = LT_NO_RULE; ed.split = 0;
                    ed.flags = LT_EDGE_UNK; ed.reserved = 0;
                } else {
                    ed = A.edges[eref];
                }
                A.path_tmp[s0 + W] = ed;
                ++W;
                r = (int)((t >> 40) & 0xFFu);
                e -= span;
            }
            A.path_len[s] = W;
            acc_W += (unsigned long long)W;
        }
        __syncwarp();
    }
    // counters
    #pragma unroll
    for (int d = 16; d; d >>= 1) {
        acc_T += __shfl_xor_sync(kFull, acc_T, d);
        acc_F += __shfl_xor_sync(kFull, acc_F, d);
        acc_B += __shfl_xor_sync(kFull, acc_B, d);
        acc_W += __shfl_xor_sync(kFull, acc_W, d);
    }
    if (lane == 0) {
        atomicAdd(A.counters + 3, acc_T);
        atomicAdd(A.counters + 4, acc_F);
        atomicAdd(A.counters + 5, acc_B);
        atomicAdd(A.counters + 6, acc_W);
    }
}

}  // namespace lt
