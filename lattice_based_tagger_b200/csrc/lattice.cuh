// lattice.cuh — lattice construction kernel (dictionary lookup + lemmatizer + eojeol enumeration).
//
// Replaces, for a whole batch, sentence_lookup_as_begin_index -> sentence_lookup ->
// MorphemeLookup.lookup -> morpheme_lookup -> lr_lookup -> MorphemeDictionary.lookup ->
// analyze_morphology -> get_lemma_candidates of the reference
// (dictionary/lookup.py:344-369, :7-62, :99-132, :212-279, :171-210; dictionary/dictionary.py:304-315;
// dictionary/lemmatizer.py:5-112).
//
// One warp per sentence, sentences pulled from an atomic work queue.  The warp stages the
// sentence in shared memory: space-stripped syllables, eojeol starts, two prefix-hash arrays
// (warp scan) and, per syllable, the conjugation-rule lists of the 1/2/3-syllable keys that start
// there.  Every dictionary test is then one probe of the hashed dictionary with a substring hash
// composed from the prefix arrays (and rule stem/eomi hashes) — no characters are compared.
//
// Output is CSR keyed by END position: edges of sentence s that end at syllable e are
// edges[end_off[sent_off[s]+e-1] .. end_off[sent_off[s]+e]), ordered by begin position and, within
// one (b, e) span, in the reference's emission order (the only order beam_search can observe,
// SURVEY App. A Q5).  The kernel runs twice: COUNT fills end_cnt / beg_cnt, a device scan turns
// end_cnt into end_off, EMIT writes the 16-byte edge records.
//
// Work decomposition per eojeol (offset o, n syllables):
//   stage 1 (lr_lookup): lanes = tasks {whole, left_i, right_i}; a left task owns bucket o+i, the
//     whole + right tasks share bucket o+n in begin order, offsets from beg_cnt.
//   stage 2 (sub-word scan, only when stage 1 found nothing): lanes = end positions; each lane
//     walks its begins in ascending order, so its bucket is written sequentially.
#pragma once
#include "tables.cuh"

namespace lt {

constexpr int kLatWarps = 4;                 // warps per CTA of the lattice kernel
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kStage2Flag = 0x80000000u;

struct LatticeArgs {
    const uint16_t* text;       // raw UTF-16 units, spaces included
    const int32_t* sent_off;    // n_sent + 1
    int32_t n_sent;
    int32_t lcap;               // max raw units of one sentence (shared-memory sizing)
    uint32_t* end_cnt;          // [n_units + 1]  edges per (sentence, end position)
    uint32_t* beg_cnt;          // [n_units + 1]  stage-1 counts of the eojeol-end bucket by begin
    const uint32_t* end_off;    // exclusive scan of end_cnt (EMIT only)
    lt_edge* edges;             // EMIT only
    int32_t* sent_len;          // [n_sent] syllables
    int32_t* sent_edges;        // [n_sent] dictionary edges
    int32_t* status;            // [n_sent]
    unsigned long long* counters;   // [0]=L [1]=P [2]=E
    unsigned int* queue;        // work-queue cursor
};

__host__ __device__ inline size_t lattice_warp_smem(int lcap) {
    // chars u16, eoj u16, nend u8 (+pad), ha u64, hb u64, rref uint2[3], cnt u32
    size_t units = (size_t)lcap + 8;
    size_t bytes = units * 2 + units * 2 + units * 1;
    bytes = (bytes + 15) & ~(size_t)15;
    bytes += units * 8 * 2 + units * 8 * 3 + units * 4;
    return (bytes + 15) & ~(size_t)15;
}

struct SentView {
    const uint16_t* ch;
    const uint64_t* ha;
    const uint64_t* hb;
    const uint2* rref;      // rref[3 * p + (key_len - 1)] = (first rule, count | k3_first << 31)
};

__device__ __forceinline__ H2 sub_hash(const DevTables& T, const SentView& v, int b, int e) {
    return h2_sub(H2{v.ha[b], v.hb[b]}, H2{v.ha[e], v.hb[e]}, pow_at(T, (uint32_t)(e - b)));
}

__device__ __forceinline__ bool is_py_space(uint32_t c) {
    // str.split() separators in the BMP
    return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20) || c == 0x85 || c == 0xA0 || c == 0x1680 ||
           (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F ||
           c == 0x3000;
}

// Prefix hashes of the staged syllables by warp scan: H[0] = 0, H[i+1] = H[i] * B + (c_i + 1).
__device__ __forceinline__ void prefix_hashes(const uint16_t* ch, int L, int lane, uint64_t* ha, uint64_t* hb) {
    H2 carry{0, 0};
    if (lane == 0) {
        ha[0] = 0;
        hb[0] = 0;
    }
    constexpr uint64_t A1 = kBaseA, A2 = A1 * A1, A4 = A2 * A2, A8 = A4 * A4, A16 = A8 * A8, A32 = A16 * A16;
    constexpr uint64_t B1 = kBaseB, B2 = B1 * B1, B4 = B2 * B2, B8 = B4 * B4, B16 = B8 * B8, B32 = B16 * B16;
    for (int base = 0; base < L; base += 32) {
        int i = base + lane;
        uint64_t v = (i < L) ? (uint64_t)ch[i] + 1u : 0u;
        uint64_t sa = v, sb = v;
        uint64_t ta, tb;
        ta = __shfl_up_sync(kFull, sa, 1);  tb = __shfl_up_sync(kFull, sb, 1);
        if (lane >= 1)  { sa += ta * A1;  sb += tb * B1; }
        ta = __shfl_up_sync(kFull, sa, 2);  tb = __shfl_up_sync(kFull, sb, 2);
        if (lane >= 2)  { sa += ta * A2;  sb += tb * B2; }
        ta = __shfl_up_sync(kFull, sa, 4);  tb = __shfl_up_sync(kFull, sb, 4);
        if (lane >= 4)  { sa += ta * A4;  sb += tb * B4; }
        ta = __shfl_up_sync(kFull, sa, 8);  tb = __shfl_up_sync(kFull, sb, 8);
        if (lane >= 8)  { sa += ta * A8;  sb += tb * B8; }
        ta = __shfl_up_sync(kFull, sa, 16); tb = __shfl_up_sync(kFull, sb, 16);
        if (lane >= 16) { sa += ta * A16; sb += tb * B16; }
        // sa = sum_{j<=lane} v_j * B^(lane-j); prefix = carry * B^(lane+1) + sa
        uint64_t pa = 1, pb = 1;   // B^(lane+1)
        {
            uint64_t xa = A1, xb = B1;
            int k = lane + 1;
            #pragma unroll
            for (int bit = 0; bit < 6; ++bit) {
                if (k & (1 << bit)) { pa *= xa; pb *= xb; }
                xa *= xa; xb *= xb;
            }
        }
        uint64_t outa = carry.a * pa + sa, outb = carry.b * pb + sb;
        if (i < L) {
            ha[i + 1] = outa;
            hb[i + 1] = outb;
        }
        carry.a = __shfl_sync(kFull, outa, 31);
        carry.b = __shfl_sync(kFull, outb, 31);
        (void)A32; (void)B32;
    }
    __syncwarp();
}

// Stage the sentence: compaction, eojeol starts, prefix hashes.  Returns syllable count; n_eoj and
// bad (non-U+0020 whitespace seen) by reference.  eoj[n_eoj] = L.
__device__ __forceinline__ int stage_sentence(const uint16_t* __restrict__ text, int s0, int s1, int lane,
                                              uint16_t* ch, uint16_t* eoj, uint64_t* ha, uint64_t* hb,
                                              int& n_eoj, bool& bad) {
    int L = 0;
    n_eoj = 0;
    bad = false;
    uint32_t prev_last = 0x20;
    for (int base = s0; base < s1; base += 32) {
        int idx = base + lane;
        bool valid = idx < s1;
        uint32_t c = valid ? (uint32_t)__ldg(text + idx) : 0x20u;
        bool space = (c == 0x20u);
        if (valid && !space && is_py_space(c)) bad = true;
        uint32_t prev = __shfl_up_sync(kFull, c, 1);
        if (lane == 0) prev = prev_last;
        bool keep = valid && !space;
        bool start = keep && (prev == 0x20u);
        unsigned km = __ballot_sync(kFull, keep);
        unsigned sm = __ballot_sync(kFull, start);
        unsigned lt_mask = (1u << lane) - 1u;
        int pos = L + __popc(km & lt_mask);
        if (keep) ch[pos] = (uint16_t)c;
        if (start) eoj[n_eoj + __popc(sm & lt_mask)] = (uint16_t)pos;
        L += __popc(km);
        n_eoj += __popc(sm);
        prev_last = __shfl_sync(kFull, c, 31);
    }
    bad = __any_sync(kFull, bad);
    if (lane == 0) eoj[n_eoj] = (uint16_t)L;
    __syncwarp();
    prefix_hashes(ch, L, lane, ha, hb);
    return L;
}

// ---- lemmatizer ---------------------------------------------------------------------------------

// One (stem, eomi) candidate: the eomi must be a known Eomi; an Adjective stem is reported before
// a Verb stem (lemmatizer.py:44-50).  Returns the number of edges (0..2).
template <bool EMIT>
__device__ __forceinline__ int lemma_candidate(const DevTables& T, H2 stem, uint32_t stem_len, H2 eomi,
                                               uint32_t eomi_len, lt_edge proto, lt_edge*& out) {
    uint64_t pe = dict_probe(T, eomi, eomi_len);
    if (!((uint32_t)(pe >> 32) & kLemEomi)) return 0;
    uint32_t ps = (uint32_t)(dict_probe(T, stem, stem_len) >> 32);
    int n = 0;
    if (ps & kLemAdj) {
        if (EMIT) { proto.tag0 = LT_TAG_ADJECTIVE; *out++ = proto; }
        ++n;
    }
    if (ps & kLemVerb) {
        if (EMIT) { proto.tag0 = LT_TAG_VERB; *out++ = proto; }
        ++n;
    }
    return n;
}

// Rules of one key applied at split position p of the word [b, e): stem = word[:p-b] + rule.stem,
// eomi = rule.eomi + word[suffix_from - b:]  (lemmatizer.py:100-102, :107-111).
template <bool EMIT>
__device__ __forceinline__ int apply_rules(const DevTables& T, const SentView& v, uint2 ref, int b, int p, int e,
                                           int suffix_from, H2 pre, lt_edge proto, lt_edge*& out) {
    int count = (int)(ref.y & 0xFFFFu);
    if (count == 0) return 0;
    H2 suf{0, 0};
    uint32_t suf_len = 0;
    if (suffix_from < e) {
        suf = sub_hash(T, v, suffix_from, e);
        suf_len = (uint32_t)(e - suffix_from);
    }
    H2 pw_suf = pow_at(T, suf_len);
    int n = 0;
    for (int r = 0; r < count; ++r) {
        RuleRec rec = rule_load(T, ref.x + r);
        H2 stem = h2_concat(pre, rec.stem, pow_at(T, rec.stem_len));
        H2 eomi = h2_concat(rec.eomi, suf, pw_suf);
        proto.rule = ref.x + r;
        n += lemma_candidate<EMIT>(T, stem, (uint32_t)(p - b) + rec.stem_len, eomi, rec.eomi_len + suf_len, proto, out);
    }
    return n;
}

// All lemma edges of the word [b, e) in get_lemma_candidates order.  `ncand` accumulates the number
// of candidates the reference generates (2 dictionary probes each in the P counter).
template <bool EMIT>
__device__ int lemma_scan(const DevTables& T, const SentView& v, int b, int e, lt_edge proto, lt_edge*& out,
                          uint32_t& ncand) {
    int total = 0;
    proto.tag1 = LT_TAG_EOMI;
    for (int p = b; p < e; ++p) {
        proto.split = (uint16_t)(p - b);
        // plain split (not at the last syllable)
        if (p < e - 1) {
            proto.rule = LT_NO_RULE;
            proto.flags = (proto.flags & LT_EDGE_IS_L) | LT_EDGE_LEMMA;
            H2 stem = sub_hash(T, v, b, p + 1);
            H2 eomi = sub_hash(T, v, p + 1, e);
            total += lemma_candidate<EMIT>(T, stem, (uint32_t)(p + 1 - b), eomi, (uint32_t)(e - p - 1), proto, out);
            ++ncand;
        }
        uint2 r1 = v.rref[3 * p + 0];
        uint2 r2 = (p + 2 <= e) ? v.rref[3 * p + 1] : make_uint2(0u, 0u);
        uint2 r3 = (p + 3 <= e) ? v.rref[3 * p + 2] : make_uint2(0u, 0u);
        int c1 = (int)(r1.y & 0xFFFFu);
        if (!(c1 | (r2.y & 0xFFFFu) | (r3.y & 0xFFFFu))) continue;
        H2 pre = (p > b) ? sub_hash(T, v, b, p) : H2{0, 0};
        // one-syllable key: the whole rule list once per rule of the key (lemmatizer.py:100-102)
        if (c1) {
            proto.flags = (proto.flags & LT_EDGE_IS_L) | LT_EDGE_LEMMA;
            if (EMIT) {
                for (int rep = 0; rep < c1; ++rep) total += apply_rules<true>(T, v, r1, b, p, e, p + 1, pre, proto, out);
            } else {
                total += c1 * apply_rules<false>(T, v, r1, b, p, e, p + 1, pre, proto, out);
            }
            ncand += (uint32_t)(c1 * c1);
        }
        // {word[i:i+2], word[i:i+3]} in set order; the eomi continues at word[i+2:] for both
        proto.flags = (proto.flags & LT_EDGE_IS_L) | LT_EDGE_LEMMA | LT_EDGE_SKIP2;
        if (p == e - 1) {
            // both slices are the last syllable itself: its rules once more, empty suffix
            if (c1) {
                total += apply_rules<EMIT>(T, v, r1, b, p, e, e, pre, proto, out);
                ncand += (uint32_t)c1;
            }
        } else {
            bool k3_first = (r3.y >> 31) != 0;
            uint2 first = k3_first ? r3 : r2;
            uint2 second = k3_first ? r2 : r3;
            total += apply_rules<EMIT>(T, v, first, b, p, e, p + 2, pre, proto, out);
            total += apply_rules<EMIT>(T, v, second, b, p, e, p + 2, pre, proto, out);
            ncand += (first.y & 0xFFFFu) + (second.y & 0xFFFFu);
        }
    }
    return total;
}

// MorphemeDictionary.lookup on [b, e): tag hits in dictionary order, then lemma edges
// (dictionary.py:304-312).  `tagmask` is the dictionary payload of the substring.
template <bool EMIT>
__device__ int full_lookup(const DevTables& T, const SentView& v, int b, int e, bool is_l, uint32_t tagmask,
                           lt_edge*& out, uint32_t& ncand) {
    lt_edge proto;
    proto.b = (uint16_t)b;
    proto.e = (uint16_t)e;
    proto.len = (uint16_t)(e - b);
    proto.tag0 = 0;
    proto.tag1 = LT_NO_TAG;
    proto.rule = LT_NO_RULE;
    proto.split = 0;
    proto.flags = is_l ? LT_EDGE_IS_L : 0;
    proto.reserved = 0;
    int n = __popc(tagmask);
    if (EMIT && tagmask) {
        for (int k = 0; k < T.n_tag_order; ++k) {
            uint32_t t = T.tag_order[k];
            if ((tagmask >> t) & 1u) {
                proto.tag0 = (uint8_t)t;
                *out++ = proto;
            }
        }
    }
    n += lemma_scan<EMIT>(T, v, b, e, proto, out, ncand);
    return n;
}

// ---- the kernel -----------------------------------------------------------------------------------

template <bool EMIT>
__global__ void __launch_bounds__(kLatWarps * 32) lattice_kernel(const DevTables T, const LatticeArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const size_t units = (size_t)A.lcap + 8;
    unsigned char* base = smem_raw + (size_t)warp * lattice_warp_smem(A.lcap);
    uint16_t* ch = reinterpret_cast<uint16_t*>(base);
    uint16_t* eoj = ch + units;
    uint8_t* nend = reinterpret_cast<uint8_t*>(eoj + units);
    size_t off = (units * 5 + 15) & ~(size_t)15;
    uint64_t* ha = reinterpret_cast<uint64_t*>(base + off);
    uint64_t* hb = ha + units;
    uint2* rref = reinterpret_cast<uint2*>(hb + units);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(rref + 3 * units);

    unsigned long long acc_L = 0, acc_P = 0, acc_E = 0;

    while (true) {
        unsigned int s = 0;
        if (lane == 0) s = atomicAdd(A.queue, 1u);
        s = __shfl_sync(kFull, s, 0);
        if (s >= (unsigned)A.n_sent) break;
        const int s0 = __ldg(A.sent_off + s), s1 = __ldg(A.sent_off + s + 1);
        int n_eoj;
        bool bad;
        const int L = stage_sentence(A.text, s0, s1, lane, ch, eoj, ha, hb, n_eoj, bad);
        SentView v{ch, ha, hb, rref};

        // conjugation-rule lists of the keys starting at every syllable
        for (int p = lane; p < L; p += 32) {
            uint32_t c0 = ch[p], c1 = (p + 1 < L) ? ch[p + 1] : 0u, c2 = (p + 2 < L) ? ch[p + 2] : 0u;
            rref[3 * p + 0] = rule_probe(T, rule_key(c0, 0, 0, 1));
            rref[3 * p + 1] = (p + 1 < L) ? rule_probe(T, rule_key(c0, c1, 0, 2)) : make_uint2(0u, 0u);
            rref[3 * p + 2] = (p + 2 < L) ? rule_probe(T, rule_key(c0, c1, c2, 3)) : make_uint2(0u, 0u);
        }
        if (!EMIT) {
            for (int p = lane; p < s1 - s0 + 1; p += 32) cnt[p] = 0;
        }
        __syncwarp();

        uint32_t sent_total = 0;
        uint32_t ncand = 0;      // lemma candidates (lane-local)
        uint32_t nsub = 0;       // distinct substrings examined (lane 0 only)

        for (int w = 0; w < n_eoj; ++w) {
            const int o = eoj[w];
            const int n = eoj[w + 1] - o;
            const int oe = o + n;
            bool stage2;
            if (!EMIT) {
                // ---------------- stage 1, COUNT ----------------
                uint32_t total = 0;
                for (int ub = 0; ub < 2 * n; ub += 32) {
                    const int u = ub + lane;
                    const bool valid = (u < 2 * n) && (u != 1);
                    int b = o, e = oe;
                    if (u >= 2) {
                        if (u & 1) b = o + (u >> 1); else e = o + (u >> 1);
                    }
                    uint32_t tagmask = 0;
                    if (valid) tagmask = (uint32_t)dict_probe(T, sub_hash(T, v, b, e), (uint32_t)(e - b));
                    const uint32_t pmask = __shfl_xor_sync(kFull, tagmask, 1);
                    const bool left = !(u & 1);
                    bool special = false;
                    if (valid && u >= 2) {
                        special = left ? ((tagmask >> LT_TAG_NOUN) & 1u) && ((pmask >> LT_TAG_JOSA) & 1u)
                                       : ((pmask >> LT_TAG_NOUN) & 1u) && ((tagmask >> LT_TAG_JOSA) & 1u);
                    }
                    uint32_t c_full = 0;
                    if (valid && !special) {
                        lt_edge* none = nullptr;
                        c_full = (uint32_t)full_lookup<false>(T, v, b, e, b == o, tagmask, none, ncand);
                    }
                    const uint32_t pc = __shfl_xor_sync(kFull, c_full, 1);
                    uint32_t c = 0;
                    if (valid) {
                        if (u == 0) c = c_full;
                        else c = special ? 1u : ((c_full > 0 && pc > 0) ? c_full : 0u);
                    }
                    if (valid && c) {
                        if (u >= 2 && left) cnt[e - 1] = c;          // bucket o+i: this task only
                        else {
                            atomicAdd(&cnt[oe - 1], c);              // bucket o+n: whole + rights
                            A.beg_cnt[s0 + b] = c;
                        }
                    }
                    total += c;
                }
                #pragma unroll
                for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(kFull, total, d);
                stage2 = (total == 0);
                sent_total += total;
                if (lane == 0) {
                    nsub += (uint32_t)(2 * n - 1);
                    if (stage2) A.beg_cnt[s0 + o] = kStage2Flag;
                }
            } else {
                // ---------------- stage 1, EMIT ----------------
                const uint32_t head = __ldg(A.beg_cnt + s0 + o);
                stage2 = (head & kStage2Flag) != 0;
                if (!stage2) {
                    uint32_t carry = 0;     // edges already placed in bucket o+n
                    const uint32_t bucket_n = __ldg(A.end_off + s0 + oe - 1);
                    for (int ub = 0; ub < 2 * n; ub += 32) {
                        const int u = ub + lane;
                        const bool valid = (u < 2 * n) && (u != 1);
                        int b = o, e = oe;
                        if (u >= 2) {
                            if (u & 1) b = o + (u >> 1); else e = o + (u >> 1);
                        }
                        const bool left = (u >= 2) && !(u & 1);
                        uint32_t c = 0;
                        if (valid) {
                            if (left) c = __ldg(A.end_off + s0 + e) - __ldg(A.end_off + s0 + e - 1);
                            else c = __ldg(A.beg_cnt + s0 + b);
                        }
                        // exclusive prefix of the shared bucket's counts over lanes
                        uint32_t mine = (valid && !left) ? c : 0u;
                        uint32_t incl = mine;
                        #pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            uint32_t t = __shfl_up_sync(kFull, incl, d);
                            if (lane >= d) incl += t;
                        }
                        const uint32_t chunk_total = __shfl_sync(kFull, incl, 31);
                        uint32_t tagmask = 0;
                        if (valid && c) tagmask = (uint32_t)dict_probe(T, sub_hash(T, v, b, e), (uint32_t)(e - b));
                        // partner's tag mask decides the Noun+Josa special case
                        uint32_t pm_in = tagmask;
                        const uint32_t pmask = __shfl_xor_sync(kFull, pm_in, 1);
                        if (valid && c) {
                            lt_edge* out = A.edges + (left ? __ldg(A.end_off + s0 + e - 1)
                                                           : bucket_n + carry + (incl - mine));
                            bool special = false;
                            if (u >= 2) {
                                special = left ? ((tagmask >> LT_TAG_NOUN) & 1u) && ((pmask >> LT_TAG_JOSA) & 1u)
                                               : ((pmask >> LT_TAG_NOUN) & 1u) && ((tagmask >> LT_TAG_JOSA) & 1u);
                            }
                            if (special) {
                                lt_edge ed;
                                ed.b = (uint16_t)b; ed.e = (uint16_t)e; ed.len = (uint16_t)n;   // len = n (Q4)
                                ed.tag0 = left ? LT_TAG_NOUN : LT_TAG_JOSA; ed.tag1 = LT_NO_TAG;
                                ed.rule = LT_NO_RULE; ed.split = 0;
                                ed.flags = left ? LT_EDGE_IS_L : 0; ed.reserved = 0;
                                *out = ed;
                            } else {
                                uint32_t dummy = 0;
                                full_lookup<true>(T, v, b, e, b == o, tagmask, out, dummy);
                            }
                        }
                        carry += chunk_total;
                    }
                }
            }

            // ---------------- stage 2: sub-word scan ----------------
            if (stage2 && n >= 2) {
                const int M = (T.max_len > 0) ? T.max_len : n;
                // (i) which positions end a stand-alone Noun found by this scan
                for (int el = 2 + lane; el <= n; el += 32) {
                    bool any_noun = false;
                    int b_lo = el - M; if (b_lo < 1) b_lo = 1;
                    for (int bl = b_lo; bl < el; ++bl) {
                        uint32_t m = (uint32_t)dict_probe(T, sub_hash(T, v, o + bl, o + el), (uint32_t)(el - bl));
                        any_noun |= (m >> LT_TAG_NOUN) & 1u;
                    }
                    nend[o + el] = any_noun ? 1 : 0;
                }
                if (lane == 0) nend[o + 1] = 0;
                __syncwarp();
                // (ii) per end position, begins ascending
                uint32_t total2 = 0;
                for (int el = 2 + lane; el <= n; el += 32) {
                    const int e = o + el;
                    lt_edge* out = EMIT ? A.edges + __ldg(A.end_off + s0 + e - 1) : nullptr;
                    uint32_t c = 0;
                    int b_lo = el - M; if (b_lo < 1) b_lo = 1;
                    for (int bl = b_lo; bl < el; ++bl) {
                        const int b = o + bl;
                        uint32_t m = (uint32_t)dict_probe(T, sub_hash(T, v, b, e), (uint32_t)(e - b));
                        lt_edge proto;
                        proto.b = (uint16_t)b; proto.e = (uint16_t)e; proto.len = (uint16_t)(e - b);
                        proto.tag0 = 0; proto.tag1 = LT_NO_TAG; proto.rule = LT_NO_RULE; proto.split = 0;
                        proto.flags = 0; proto.reserved = 0;
                        // stand-alone tags in list order (lookup.py:104-105), then Josa after a Noun
                        const int order[5] = {LT_TAG_NOUN, LT_TAG_ADVERB, LT_TAG_EXCLAMATION, LT_TAG_DETERMINER, LT_TAG_NUMBER};
                        #pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if ((m >> order[k]) & 1u) {
                                if (EMIT) { proto.tag0 = (uint8_t)order[k]; *out++ = proto; }
                                ++c;
                            }
                        }
                        if (nend[b] && ((m >> LT_TAG_JOSA) & 1u)) {
                            if (EMIT) { proto.tag0 = LT_TAG_JOSA; *out++ = proto; }
                            ++c;
                        }
                        c += (uint32_t)lemma_scan<EMIT>(T, v, b, e, proto, out, ncand);
                    }
                    if (!EMIT) cnt[e - 1] = c;
                    total2 += c;
                }
                if (!EMIT) {
                    #pragma unroll
                    for (int d = 16; d; d >>= 1) total2 += __shfl_xor_sync(kFull, total2, d);
                    sent_total += total2;
                    if (lane == 0) {
                        // pairs (b, e), 1 <= b < e <= min(b+M, n), minus the (b, n) already examined by stage 1
                        uint32_t pairs = 0;
                        for (int bl = 1; bl < n; ++bl) pairs += (uint32_t)min(M, n - bl);
                        nsub += pairs - (uint32_t)min(M, n - 1);
                    }
                }
                __syncwarp();
            }
            __syncwarp();
        }

        if (!EMIT) {
            __syncwarp();
            for (int p = lane; p < s1 - s0; p += 32) A.end_cnt[s0 + p] = cnt[p];
            #pragma unroll
            for (int d = 16; d; d >>= 1) ncand += __shfl_xor_sync(kFull, ncand, d);
            if (lane == 0) {
                A.sent_len[s] = L;
                A.sent_edges[s] = (int32_t)sent_total;
                int st = LT_SENT_OK;
                if (bad) st = LT_SENT_BAD_SPACE;
                else if (L > 0 && sent_total == 0) st = LT_SENT_NO_EDGES;
                A.status[s] = st;
                acc_L += (unsigned long long)L;
                acc_P += (unsigned long long)nsub + 2ull * ncand;
                acc_E += sent_total;
            }
        }
        __syncwarp();
    }
    if (!EMIT && lane == 0) {
        atomicAdd(A.counters + 0, acc_L);
        atomicAdd(A.counters + 1, acc_P);
        atomicAdd(A.counters + 2, acc_E);
    }
}

}  // namespace lt
