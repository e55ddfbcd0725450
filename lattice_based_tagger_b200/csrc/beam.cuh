// beam.cuh — transition scoring + fixed-window beam search kernel.
//
// Replaces beam_search / Beam / Sequence (beam/beam.py:5-124) and the score functions
// (beam/score_funcs.py:18-144) with their feature templates (features/feature.py:76-121).
//
// One warp per sentence (atomic work queue).  Hypotheses are back-pointer entries in a shared-
// memory ring of the last 9 end positions (window 8, beam.py:30): score, the hashes of the last
// two words and of the contextual morpheme, and a few tag bits — everything the next transition's
// features depend on.  Per end position e the warp
//   1. reads the CSR bucket of e and counts edges per begin (the bucket is sorted by begin);
//   2. enumerates candidates in the reference's generation order — begin ascending, parent rank
//      ascending, edge order ascending (beam.py:30-48) — 32 at a time, one per lane;
//   3. scores each lane's transition: the score program in BeamScoreFunctions order, every fp64
//      add in the reference's association, feature weights gathered from the
//      HBM feature table by hashed key with all first-slot loads in flight together, the tag x tag
//      matrix and the length vectors from shared memory;
//   4. keeps the best `beam` candidates in a sorted shared-memory list; insertion is strict
//      (a later equal score never displaces an earlier one), which is exactly the stable sort of
//      Beam.append (beam.py:83-86);
//   5. writes the survivors as new ring entries and one 8-byte back-pointer each to the HBM trail.
// The best path is recovered from the trail and written as 16-byte edge records.
#pragma once
#include "lattice.cuh"
#include "tables.cuh"

namespace lt {

constexpr int kRing = LT_WINDOW + 1;
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
    const uint32_t* end_off;
    const lt_edge* edges;
    const int32_t* status;      // from the lattice pass
    uint64_t* trail;            // [(n_units) * beam] back-pointers
    lt_edge* path_tmp;          // [n_units] best path, reversed, at the sentence's offset
    int32_t* path_len;          // [n_sent]
    double* scores;             // [n_sent]
    unsigned long long* counters;   // [3]=T [4]=F [5]=Bk [6]=W
    unsigned int* queue;
};

__host__ __device__ inline size_t beam_warp_smem(int lcap, int beam) {
    size_t units = (size_t)lcap + 8;
    size_t bytes = units * 8 * 2;                    // ha, hb
    bytes += (size_t)kRing * beam * (8 + 48);        // score, wj, wi, mc
    bytes += (size_t)beam * 8;                       // list keys
    bytes += units * 2;                              // chars
    bytes += (size_t)kRing * beam * 4;               // meta
    bytes += (size_t)beam * 4;                       // list payloads
    bytes += 16;                                     // ring sizes
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

struct ParentView {
    double score;
    H2 wj, wi, mc;
    uint32_t meta;
};

struct EdgeView {
    int b, e;
    uint32_t len, tag0, tag1, rule, split, flags;
    H2 wk, mk, m1;      // hashes of word, morph0, morph1
    uint32_t m1_valid;
};

__device__ __forceinline__ void edge_hashes(const DevTables& T, const SentView& v, EdgeView& k, bool need_m1) {
    k.wk = sub_hash(T, v, k.b, k.e);
    k.mk = k.wk;
    k.m1 = H2{0, 0};
    k.m1_valid = 0;
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
        k.m1_valid = 1;
    }
}

// increment of one transition: BeamScoreFunctions.score (score_funcs.py:50-54)
__device__ __forceinline__ double transition_increment(const DevTables& T, const unsigned char* dense_smem,
                                                       const ParentView& P, const EdgeView& k, uint32_t& nfeat) {
    double inc = 0.0;
    const uint32_t tj = P.meta & kMetaTagMask;
    const uint32_t tk = k.tag0;
    const H2 zero{0, 0};
    for (int f = 0; f < T.n_funcs; ++f) {
        const lt_func& fn = T.funcs[f];
        double val;
        if (fn.kind == LT_FUNC_REG) {
            // score_funcs.py:65-73
            if (tk == LT_TAG_UNK) val = __dmul_rn(fn.p[0], __dadd_rn((double)k.len, 0.1));
            else val = __dmul_rn(fn.p[1], (double)k.len);
            val = __dadd_rn(0.0, val);
            if (k.len == 1 && tk == LT_TAG_NOUN) val = __dadd_rn(val, fn.p[2]);
        } else if (fn.kind == LT_FUNC_MPREF) {
            // score_funcs.py:84-88
            FKey k0 = feature_key(kKindMPref, f, k.mk, zero, zero, tk, 0);
            uint4 s0 = feat_first(T, k0);
            double a = 0.0, b2 = 0.0;
            if (k.tag1 != LT_NO_TAG) {
                FKey k1 = feature_key(kKindMPref, f, k.m1, zero, zero, k.tag1, 0);
                uint4 s1 = feat_first(T, k1);
                feat_resolve(T, k1, s1, b2);
            }
            feat_resolve(T, k0, s0, a);
            val = (k.tag1 != LT_NO_TAG) ? __dadd_rn(a, b2) : a;
        } else if (fn.kind == LT_FUNC_WPREF) {
            // score_funcs.py:99-100
            FKey k0 = feature_key(kKindWPref, f, k.wk, zero, zero, tk, 0);
            uint4 s0 = feat_first(T, k0);
            val = 0.0;
            feat_resolve(T, k0, s0, val);
        } else {
            // SimpleTrigramFeatureScore.score (score_funcs.py:137-144) over trigram_encoder's templates
            const DenseView D = dense_view(dense_smem + (size_t)T.func_dense[f] * dense_block_bytes(T.n_tags), T.n_tags);
            double v[9];
            uint32_t present = 0;
            const bool has_i = (P.meta & kMetaHasI) != 0;
            const bool j_unk = (tj == LT_TAG_UNK);
            const bool ctx8 = ((kCtxMask >> tk) & 1u) && (tk < 32) && (P.meta & kMetaHasCtx);
            const bool t4_hashed = k.len >= (uint32_t)kT4Dense;
            nfeat += 6u + (j_unk ? 1u : 0u) + (has_i ? 1u : 0u) + (ctx8 ? 1u : 0u);
            // keys
            FKey q0 = feature_key(0, f, P.wj, k.wk, zero, tk, 0);
            FKey q1 = feature_key(1, f, P.wj, zero, zero, tk, 0);
            FKey q2 = feature_key(2, f, k.wk, zero, zero, tj, tk);
            FKey q5 = feature_key(5, f, k.wk, zero, zero, tk, (k.flags & LT_EDGE_IS_L) ? 1u : 0u);
            FKey q7 = feature_key(7, f, P.wi, P.wj, k.wk, 0, 0);
            FKey q8 = feature_key(8, f, P.mc, k.mk, zero, 0, 0);
            FKey q4 = feature_key(4, f, zero, zero, zero, k.len, 0);
            // all first-slot loads in flight before any is consumed
            uint4 s0 = feat_first(T, q0);
            uint4 s1 = feat_first(T, q1);
            uint4 s2 = feat_first(T, q2);
            uint4 s5 = feat_first(T, q5);
            uint4 s7 = has_i ? feat_first(T, q7) : make_uint4(0, 0, 0, 0);
            uint4 s8 = ctx8 ? feat_first(T, q8) : make_uint4(0, 0, 0, 0);
            uint4 s4 = t4_hashed ? feat_first(T, q4) : make_uint4(0, 0, 0, 0);
            #pragma unroll
            for (int i = 0; i < 9; ++i) v[i] = 0.0;
            if (feat_resolve(T, q0, s0, v[0])) present |= 1u << 0;
            if (feat_resolve(T, q1, s1, v[1])) present |= 1u << 1;
            if (feat_resolve(T, q2, s2, v[2])) present |= 1u << 2;
            if ((D.m3[tj] >> tk) & 1u) { v[3] = D.t3[tj * T.n_tags + tk]; present |= 1u << 3; }
            if (t4_hashed) {
                if (feat_resolve(T, q4, s4, v[4])) present |= 1u << 4;
            } else if ((D.m4[k.len >> 5] >> (k.len & 31)) & 1u) {
                v[4] = D.t4[k.len]; present |= 1u << 4;
            }
            if (feat_resolve(T, q5, s5, v[5])) present |= 1u << 5;
            if (j_unk) {
                const uint32_t ul = (P.meta >> kMetaUnkLenShift) & 0xFu;
                if ((D.m6[0] >> ul) & 1u) { v[6] = D.t6[ul]; present |= 1u << 6; }
            }
            if (has_i && feat_resolve(T, q7, s7, v[7])) present |= 1u << 7;
            if (ctx8 && feat_resolve(T, q8, s8, v[8])) present |= 1u << 8;
            val = present ? numpy_order_sum9(v, present) : 0.0;
        }
        inc = __dadd_rn(inc, val);
    }
    return inc;
}

__global__ void __launch_bounds__(256) beam_kernel(const DevTables T, const BeamArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int K = A.beam;
    const int NT = T.n_tags;

    // CTA-shared dense tables (tag x tag matrix, length vectors)
    const size_t dense_bytes = ((size_t)T.n_tri * dense_block_bytes(NT) + 15) & ~(size_t)15;
    unsigned char* dense_smem = smem_raw;
    for (size_t i = threadIdx.x * 4; i < (size_t)T.n_tri * dense_block_bytes(NT); i += blockDim.x * 4)
        *reinterpret_cast<uint32_t*>(dense_smem + i) = *reinterpret_cast<const uint32_t*>(T.dense + i);
    __syncthreads();

    const size_t units = (size_t)A.lcap + 8;
    unsigned char* wbase = smem_raw + dense_bytes + (size_t)warp * beam_warp_smem(A.lcap, K);
    uint64_t* ha = reinterpret_cast<uint64_t*>(wbase);
    uint64_t* hb = ha + units;
    double* e_score = reinterpret_cast<double*>(hb + units);
    H2* e_wj = reinterpret_cast<H2*>(e_score + kRing * K);
    H2* e_wi = e_wj + kRing * K;
    H2* e_mc = e_wi + kRing * K;
    double* l_key = reinterpret_cast<double*>(e_mc + kRing * K);
    uint16_t* ch = reinterpret_cast<uint16_t*>(l_key + K);
    uint32_t* e_meta = reinterpret_cast<uint32_t*>(ch + units);
    uint32_t* l_pay = e_meta + kRing * K;
    uint32_t* ring_n = l_pay + K;      // kRing entries used (as u8-in-u32: keep simple)
    // ring_n needs kRing words; beam_warp_smem reserves 16 bytes -> use bytes
    uint8_t* nbeam = reinterpret_cast<uint8_t*>(ring_n);

    unsigned long long acc_T = 0, acc_F = 0, acc_B = 0, acc_W = 0;
    bool need_m1 = false;
    for (int f = 0; f < T.n_funcs; ++f) need_m1 |= (T.funcs[f].kind == LT_FUNC_MPREF);

    while (true) {
        unsigned int s = 0;
        if (lane == 0) s = atomicAdd(A.queue, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= (unsigned)A.n_sent) break;
        const int s0 = __ldg(A.sent_off + s), s1 = __ldg(A.sent_off + s + 1);
        const int st = __ldg(A.status + s);
        if (st != LT_SENT_OK) {
            if (lane == 0) { A.path_len[s] = 0; A.scores[s] = 0.0; }
            continue;
        }
        // ---- stage syllables + prefix hashes (same arithmetic as the lattice kernel) ----
        int L = 0;
        for (int base = s0; base < s1; base += 32) {
            int idx = base + lane;
            bool valid = idx < s1;
            uint32_t c = valid ? (uint32_t)__ldg(A.text + idx) : 0x20u;
            bool keep = valid && (c != 0x20u);
            unsigned km = __ballot_sync(kFull, keep);
            int pos = L + __popc(km & ((1u << lane) - 1u));
            if (keep) ch[pos] = (uint16_t)c;
            L += __popc(km);
        }
        __syncwarp();
        prefix_hashes(ch, L, lane, ha, hb);
        SentView v{ch, ha, hb, nullptr};

        if (L == 0) {
            if (lane == 0) { A.path_len[s] = 0; A.scores[s] = 0.0; }
            continue;
        }

        // beam[0] = [BOS] (beam.py:21-23)
        if (lane == 0) {
            e_score[0] = 0.0;
            e_wj[0] = T.bos;
            e_wi[0] = H2{0, 0};
            e_mc[0] = H2{0, 0};
            e_meta[0] = (uint32_t)LT_TAG_BOS;
            nbeam[0] = 1;
        }
        __syncwarp();

        for (int e = 1; e <= L; ++e) {
            const int slot_e = e % kRing;
            const uint32_t es = __ldg(A.end_off + s0 + e - 1), ee = __ldg(A.end_off + s0 + e);
            const int jmax = (e < LT_WINDOW) ? e : LT_WINDOW;
            // edges per span (bucket sorted by begin ascending = span descending)
            uint32_t cnt[LT_WINDOW + 1];
            #pragma unroll
            for (int j = 0; j <= LT_WINDOW; ++j) cnt[j] = 0;
            for (uint32_t base = es; base < ee; base += 32) {
                uint32_t idx = base + lane;
                int span = 0;
                if (idx < ee) {
                    uint32_t be = __ldg(reinterpret_cast<const uint32_t*>(A.edges + idx));
                    span = (int)(be >> 16) - (int)(be & 0xFFFFu);
                }
                #pragma unroll
                for (int j = 1; j <= LT_WINDOW; ++j) cnt[j] += __popc(__ballot_sync(kFull, span == j));
            }
            // group starts and candidate counts, spans from jmax down to 1 (begin ascending)
            uint32_t gstart[LT_WINDOW + 1], ncand[LT_WINDOW + 1];
            uint32_t in_window = 0;
            #pragma unroll
            for (int j = 1; j <= LT_WINDOW; ++j) in_window += cnt[j];
            uint32_t run = ee - in_window;          // first edge inside the window
            uint32_t N = 0;
            #pragma unroll
            for (int j = LT_WINDOW; j >= 1; --j) {
                gstart[j] = run;
                run += cnt[j];
                uint32_t np = (j <= jmax) ? (uint32_t)nbeam[(e - j) % kRing] : 0u;
                uint32_t nedge = cnt[j] ? cnt[j] : 1u;
                ncand[j] = np * nedge;
                N += ncand[j];
            }

            int nl = 0;     // entries in the sorted list (warp-uniform)
            for (uint32_t c0 = 0; c0 < N; c0 += 32) {
                uint32_t c = c0 + lane;
                bool valid = c < N;
                int j = 0;
                uint32_t rem = c;
                if (valid) {
                    #pragma unroll
                    for (int jj = LT_WINDOW; jj >= 1; --jj) {
                        if (j == 0) {
                            if (rem < ncand[jj]) j = jj; else rem -= ncand[jj];
                        }
                    }
                }
                double newscore = 0.0;
                uint32_t pay = 0;
                if (valid) {
                    // select by dynamic j without local-memory arrays
                    uint32_t cj = 0, gs = 0;
                    #pragma unroll
                    for (int jj = 1; jj <= LT_WINDOW; ++jj)
                        if (jj == j) { cj = cnt[jj]; gs = gstart[jj]; }
                    const bool unk_edge = (cj == 0);
                    const uint32_t nedge = unk_edge ? 1u : cj;
                    const uint32_t prank = rem / nedge, eidx = rem - prank * nedge;
                    const int pslot = ((e - j) % kRing) * K + (int)prank;
                    ParentView P;
                    P.score = e_score[pslot];
                    P.wj = e_wj[pslot];
                    P.wi = e_wi[pslot];
                    P.mc = e_mc[pslot];
                    P.meta = e_meta[pslot];
                    EdgeView k;
                    k.b = e - j; k.e = e;
                    uint32_t eref;
                    if (unk_edge) {
                        k.len = (uint32_t)j; k.tag0 = LT_TAG_UNK; k.tag1 = LT_NO_TAG; k.rule = LT_NO_RULE;
                        k.split = 0; k.flags = LT_EDGE_UNK;
                        eref = kTrailUnk;
                    } else {
                        eref = gs + eidx;
                        uint4 raw = ldg16(A.edges + eref);
                        k.len = raw.y & 0xFFFFu;
                        k.tag0 = (raw.y >> 16) & 0xFFu;
                        k.tag1 = (raw.y >> 24) & 0xFFu;
                        k.rule = raw.z;
                        k.split = raw.w & 0xFFFFu;
                        k.flags = (raw.w >> 16) & 0xFFu;
                    }
                    // two unknown words in a row are only allowed from the window's first begin (beam.py:44-45)
                    const bool parent_unk = (P.meta & kMetaTagMask) == LT_TAG_UNK;
                    if (parent_unk && k.tag0 == LT_TAG_UNK && j < jmax) {
                        valid = false;
                    } else {
                        edge_hashes(T, v, k, need_m1);
                        uint32_t nfeat = 0;
                        double inc = transition_increment(T, dense_smem, P, k, nfeat);
                        newscore = __dadd_rn(P.score, inc);            // Sequence.add, beam.py:115
                        newscore = __dadd_rn(newscore, 0.0);           // -0.0 sorts as 0.0
                        pay = (unk_edge ? 0x80000000u : 0u) | ((uint32_t)j << 27) | (prank << 20) | (unk_edge ? 0u : (eref - es));
                        acc_T += 1;
                        acc_F += nfeat;
                    }
                }
                // ---- strict insertion into the sorted top-K list, lanes in generation order ----
                bool want = valid && (nl < K || newscore > l_key[K - 1]);
                unsigned m = __ballot_sync(kFull, want);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const double key = __shfl_sync(kFull, newscore, src);
                    const uint32_t kp = __shfl_sync(kFull, pay, src);
                    if (nl == K && !(key > l_key[K - 1])) continue;
                    // position = number of entries with key >= new key (equal keys stay in front)
                    int pos = 0;
                    for (int i0 = 0; i0 < nl; i0 += 32) {
                        int i = i0 + lane;
                        bool ge = (i < nl) && (l_key[i] >= key);
                        pos += __popc(__ballot_sync(kFull, ge));
                    }
                    const int last = (nl < K) ? nl : K - 1;     // index that receives the shifted tail
                    // shift [pos, last) one step down
                    double mk0 = 0.0, mk1 = 0.0;
                    uint32_t mp0 = 0, mp1 = 0;
                    const int i_a = pos + 1 + lane, i_b = pos + 33 + lane;
                    if (i_a <= last) { mk0 = l_key[i_a - 1]; mp0 = l_pay[i_a - 1]; }
                    if (i_b <= last) { mk1 = l_key[i_b - 1]; mp1 = l_pay[i_b - 1]; }
                    __syncwarp();
                    if (i_a <= last) { l_key[i_a] = mk0; l_pay[i_a] = mp0; }
                    if (i_b <= last) { l_key[i_b] = mk1; l_pay[i_b] = mp1; }
                    if (lane == 0) { l_key[pos] = key; l_pay[pos] = kp; }
                    if (nl < K) ++nl;
                    __syncwarp();
                }
            }

            // ---- survivors -> ring entries + trail ----
            for (int r0 = 0; r0 < nl; r0 += 32) {
                const int r = r0 + lane;
                if (r < nl) {
                    const uint32_t kp = l_pay[r];
                    const int j = (int)((kp >> 27) & 0xFu);
                    const uint32_t prank = (kp >> 20) & 0x7Fu;
                    const bool unk_edge = (kp >> 31) != 0;
                    const int pslot = ((e - j) % kRing) * K + (int)prank;
                    EdgeView k;
                    k.b = e - j; k.e = e;
                    uint32_t eref = kTrailUnk;
                    if (unk_edge) {
                        k.len = (uint32_t)j; k.tag0 = LT_TAG_UNK; k.tag1 = LT_NO_TAG; k.rule = LT_NO_RULE;
                        k.split = 0; k.flags = LT_EDGE_UNK;
                    } else {
                        eref = es + (kp & 0xFFFFFu);
                        uint4 raw = ldg16(A.edges + eref);
                        k.len = raw.y & 0xFFFFu;
                        k.tag0 = (raw.y >> 16) & 0xFFu;
                        k.tag1 = (raw.y >> 24) & 0xFFu;
                        k.rule = raw.z;
                        k.split = raw.w & 0xFFFFu;
                        k.flags = (raw.w >> 16) & 0xFFu;
                    }
                    edge_hashes(T, v, k, false);
                    const uint32_t pmeta = e_meta[pslot];
                    const uint32_t tj = pmeta & kMetaTagMask;
                    const bool k_ctx = (k.tag0 < 32) && ((kCtxMask >> k.tag0) & 1u);
                    const bool j_ctx = (tj < 32) && ((kCtxMask >> tj) & 1u);
                    const int dst = slot_e * K + r;
                    e_score[dst] = l_key[r];
                    e_wj[dst] = k.wk;
                    e_wi[dst] = e_wj[pslot];
                    e_mc[dst] = k_ctx ? k.mk : (j_ctx ? e_mc[pslot] : H2{0, 0});
                    uint32_t ul = k.len < 8u ? k.len : 8u;
                    e_meta[dst] = k.tag0 | kMetaHasI | ((k_ctx || j_ctx) ? kMetaHasCtx : 0u) | (ul << kMetaUnkLenShift);
                    A.trail[(size_t)(s0 + e - 1) * K + r] =
                        (uint64_t)eref | ((uint64_t)j << 32) | ((uint64_t)prank << 40);
                }
            }
            if (lane == 0) nbeam[slot_e] = (uint8_t)nl;
            acc_B += (lane == 0) ? (unsigned long long)nl : 0ull;
            __syncwarp();
        }

        // ---- best path: matures[0] (tagger.py:78) ----
        if (lane == 0) {
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
    }
    if (lane == 0) {
        atomicAdd(A.counters + 3, acc_T);
        atomicAdd(A.counters + 4, acc_F);
        atomicAdd(A.counters + 5, acc_B);
        atomicAdd(A.counters + 6, acc_W);
    }
}

}  // namespace lt
