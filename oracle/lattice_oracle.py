"""CPU oracle for the lattice tagger's decode path  —  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A pure-Python restatement of what `lattice_tagger.Tagger.tag` computes: eojeol lookup that builds
the morpheme lattice, additive scoring of (hypothesis, word) transitions, and the fixed-window
beam search.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module; the product path (`lattice_based_tagger_b200`)
never does and has no CPU fallback.

Parity is PINNED: `tests/test_oracle_vs_reference.py` compares this file with the imported
reference (`/root/reference`, when present) on randomised dictionaries, sentences and scorers,
and `tests/golden/*.json.gz` hold outputs of the reference itself (written by
`oracle/gen_golden.py`) that `tests/test_oracle_golden.py` replays everywhere.

Each function cites the reference lines it restates (paths relative to
`/root/reference/lattice_tagger/`).  Reference quirks are reproduced on purpose — SURVEY.md
Appendix A lists them (Q1..Q10).

An edge is the 9-tuple `(word, morph0, morph1, tag0, tag1, len, b, e, is_l)`, field for field the
reference's `Word` namedtuple, so results compare with `==` against either implementation.
"""

NOUN, JOSA, ADJECTIVE, VERB, EOMI = 'Noun', 'Josa', 'Adjective', 'Verb', 'Eomi'
ADVERB, EXCLAMATION, DETERMINER, NUMBER = 'Adverb', 'Exclamation', 'Determiner', 'Number'
BOS, EOS, UNK = 'BOS', 'EOS', 'Unknown'

DEFAULT_STANDALONES = (NOUN, ADVERB, EXCLAMATION, DETERMINER, NUMBER)   # dictionary/lookup.py:104-105
CONTEXTUAL = frozenset((NOUN, ADVERB, ADJECTIVE, VERB))                 # features/feature.py:88
WINDOW = 8                                                             # beam/beam.py:5 (max_len)

# field indices of an edge tuple
WORD, MORPH0, MORPH1, TAG0, TAG1, LEN, B, E, IS_L = range(9)


class Counters:
    """Work counters of SURVEY.md §8(d); they define the algorithmic byte counts of the roofline.

    L syllables, P dictionary string probes (distinct substrings examined per eojeol + 2 per
    lemma candidate), E dictionary edges, T scored transitions, F feature tuples generated,
    Bk kept beam entries, W words on the returned path.
    """
    __slots__ = ('L', 'P', 'E', 'T', 'F', 'Bk', 'W')

    def __init__(self):
        self.L = self.P = self.E = self.T = self.F = self.Bk = self.W = 0

    def as_dict(self):
        return {name: getattr(self, name) for name in self.__slots__}

    def bytes_lattice(self):
        return 2 * self.L + 16 * self.P + 16 * self.E

    def bytes_beam(self):
        return 16 * self.E + 16 * self.F + 8 * self.Bk + 8 + 4 * self.W


# --------------------------------------------------------------------------------------------
# dictionary view
# --------------------------------------------------------------------------------------------

class DictView:
    """The five attributes of a morpheme dictionary the decode path reads, plus the iteration
    order of the 2- vs 3-syllable conjugation set (Q3), which depends on the process's string
    hashing.  `k3_first` may be given to replay an order recorded in another process (golden
    fixtures); by default it is observed here exactly as the reference would see it."""

    def __init__(self, dictionary, k3_first=None):
        if not hasattr(dictionary, 'rules'):
            raise ValueError('dictionary must be MorphemeDictionary')    # dictionary/lookup.py:101-102
        self.tag_to_morphs = dictionary.tag_to_morphs
        self.rules = dictionary.rules
        self.verbs = dictionary.verbs
        self.adjectives = dictionary.adjectives
        self.eomis = dictionary.eomis
        self.k3_first = k3_first
        self.max_len = self._max_len(DEFAULT_STANDALONES)

    def _max_len(self, standalones):
        # MorphemeLookup._find_max_len, dictionary/lookup.py:123-132
        wanted = set(standalones) | {VERB, ADJECTIVE}
        best = 0
        for tag, morphs in self.tag_to_morphs.items():
            if tag in wanted:
                best = max(best, max(len(m) for m in morphs))    # ValueError on an empty set, as there
        return best

    def conj_pair(self, k2, k3):
        """Iteration order of the set {k2, k3} (dictionary/lemmatizer.py:107)."""
        if k2 == k3:
            return (k2,)
        if self.k3_first is not None and k2 in self.rules and k3 in self.rules:
            return (k3, k2) if self.k3_first.get(k3, False) else (k2, k3)
        return tuple({k2, k3})

    def check(self, morph, tag):
        return morph in self.tag_to_morphs.get(tag, ())           # dictionary/dictionary.py:238-239

    def tags_of(self, morph):
        return [t for t, morphs in self.tag_to_morphs.items() if morph in morphs]   # :241-242


def lemma_candidates(word, view, counters=None):
    """(stem, eomi) candidates of a surface form — dictionary/lemmatizer.py:90-112.

    Per position i: the plain split (not at the last syllable); the one-syllable rules of
    word[i], emitted once per rule of that key (nested duplicate loop, :100-102); then the rules
    of the 2- and 3-syllable slices in set order, both consuming only ONE more syllable
    (`r[1:]`, :109)."""
    rules = view.rules
    out = []
    last = len(word) - 1
    for i in range(len(word)):
        head, tail, before = word[:i + 1], word[i + 1:], word[:i]
        if i < last:
            out.append((head, tail))
        single = rules.get(word[i], ())
        for _ in single:
            for stem, eomi in single:
                out.append((before + stem, eomi + tail))
        for conj in view.conj_pair(word[i:i + 2], word[i:i + 3]):
            for stem, eomi in rules.get(conj, ()):
                out.append((before + stem, eomi + tail[1:]))
    if counters is not None:
        counters.P += 2 * len(out)
    return out


def lemmatize(word, view, counters=None):
    """Dictionary-checked analyses — dictionary/lemmatizer.py:43-51: the eomi must be known; an
    Adjective stem is reported before a Verb stem."""
    found = []
    for stem, eomi in lemma_candidates(word, view, counters):
        if eomi not in view.eomis:
            continue
        if stem in view.adjectives:
            found.append(((stem, ADJECTIVE), (eomi, EOMI)))
        if stem in view.verbs:
            found.append(((stem, VERB), (eomi, EOMI)))
    return found


def full_lookup(word, view, b, is_l, counters=None):
    """MorphemeDictionary.lookup — dictionary/dictionary.py:304-312: one single-morpheme edge
    per tag holding the string (dictionary iteration order, Q1), then the lemmatised analyses."""
    n = len(word)
    e = b + n
    edges = [(word, word, None, tag, None, n, b, e, is_l) for tag in view.tags_of(word)]
    for (m0, t0), (m1, t1) in lemmatize(word, view, counters):
        edges.append((word, m0, m1, t0, t1, n, b, e, is_l))
    return edges


def eojeol_lookup(eojeol, view, offset, counters=None):
    """MorphemeLookup.lookup -> morpheme_lookup(prefer_exact_match=True) — dictionary/lookup.py:116-121,
    :212-279, with lr_lookup(prefer_exact_match=False) (:171-210) inlined as its first stage."""
    n = len(eojeol)
    examined = set()

    # stage 1: whole eojeol, then every left/right split (lr_lookup, :191-210)
    examined.add((0, n))
    edges = full_lookup(eojeol, view, offset, True, counters)
    for i in range(1, n):
        left, right = eojeol[:i], eojeol[i:]
        examined.add((0, i))
        examined.add((i, n))
        if view.check(left, NOUN) and view.check(right, JOSA):
            # Noun + Josa special case: BOTH edges carry len = n (Q4, :201-202)
            edges.append((left, left, None, NOUN, None, n, offset, offset + i, True))
            edges.append((right, right, None, JOSA, None, n, offset + i, offset + n, False))
            continue
        lset = full_lookup(left, view, offset, True, counters)
        rset = full_lookup(right, view, offset + i, False, counters)
        if lset and rset:
            edges += lset
            edges += rset
    if edges:                                                    # :237-238
        if counters is not None:
            counters.P += len(examined)
        return edges

    # (:248-256 repeat the Noun+Josa test of stage 1 and cannot fire once stage 1 found nothing)

    # stage 2: sub-word scan, b starts at 1 so is_l is never set (:259-277)
    standalones = DEFAULT_STANDALONES
    noun_end = [False] * (n + 1)
    max_len = view.max_len if view.max_len > 0 else n            # :229-230
    for b in range(1, n):
        for e in range(b + 1, min(b + max_len, n) + 1):
            sub = eojeol[b:e]
            examined.add((b, e))
            for tag in standalones:
                if view.check(sub, tag):
                    edges.append((sub, sub, None, tag, None, e - b, offset + b, offset + e, False))
                    if tag == NOUN:
                        noun_end[e] = True
            if noun_end[b] and view.check(sub, JOSA):
                edges.append((sub, sub, None, JOSA, None, e - b, offset + b, offset + e, False))
            for (m0, t0), (m1, t1) in lemmatize(sub, view, counters):
                edges.append((sub, m0, m1, t0, t1, e - b, offset + b, offset + e, False))
    if counters is not None:
        counters.P += len(examined)
    return edges


def lr_lookup(eojeol, view, offset, prefer_exact_match=True):
    """LRLookup.lookup -> lr_lookup — dictionary/lookup.py:75-85, :171-210: the whole eojeol's
    analyses (alone, with prefer_exact_match, when there are any), then every left/right split
    whose two sides are both known, the Noun + Josa split with its `len = n` quirk (Q4)."""
    n = len(eojeol)
    edges = full_lookup(eojeol, view, offset, True)
    if prefer_exact_match and edges:
        return edges
    for i in range(1, n):
        left, right = eojeol[:i], eojeol[i:]
        if view.check(left, NOUN) and view.check(right, JOSA):
            edges.append((left, left, None, NOUN, None, n, offset, offset + i, True))
            edges.append((right, right, None, JOSA, None, n, offset + i, offset + n, False))
            continue
        lset = full_lookup(left, view, offset, True)
        rset = full_lookup(right, view, offset + i, False)
        if lset and rset:
            edges += lset
            edges += rset
    return edges


def word_lookup(eojeol, view, offset, prefer_exact_match=True):
    """WordLookup.lookup -> word_lookup — dictionary/lookup.py:87-97, :134-169: the whole eojeol's
    analyses (alone, with prefer_exact_match, when there are any), then the full lookup of EVERY
    substring, the whole eojeol included once more (so its analyses appear twice without
    prefer_exact_match); `is_l` marks substrings that start the eojeol."""
    n = len(eojeol)
    edges = full_lookup(eojeol, view, offset, True)
    if prefer_exact_match and edges:
        return edges
    for b in range(n):
        for e in range(b + 1, n + 1):
            edges += full_lookup(eojeol[b:e], view, offset + b, b == 0)
    return edges


def flatten_edges(edges):
    """flatten_words — dictionary/dictionary.py:114-167: a two-morpheme word becomes two
    one-morpheme words meeting at min(e, b + len(morph0)); a leading jamo of the second morpheme
    does not count as a syllable of its `len`."""
    out = []
    for w in edges:
        if w[TAG1] is None:
            out.append(w)
            continue
        len0, len1 = len(w[MORPH0]), len(w[MORPH1])
        m = min(w[E], w[B] + len0)
        if 'ㄱ' <= w[MORPH1][0] <= 'ㅎ':
            len1 -= 1
        out.append((w[MORPH0], w[MORPH0], None, w[TAG0], None, len0, w[B], m, w[IS_L]))
        out.append((w[MORPH1], w[MORPH1], None, w[TAG1], None, len1, m, w[E], False))
    return out


#: lookup modes: name -> eojeol lookup (eojeol, view, offset, counters) of the reference's EojeolLookup classes
LOOKUPS = {
    'morpheme': lambda eojeol, view, offset, counters=None: eojeol_lookup(eojeol, view, offset, counters),
    'lr': lambda eojeol, view, offset, counters=None: lr_lookup(eojeol, view, offset, True),
    'lr_all': lambda eojeol, view, offset, counters=None: lr_lookup(eojeol, view, offset, False),
    'word': lambda eojeol, view, offset, counters=None: word_lookup(eojeol, view, offset, True),
    'word_all': lambda eojeol, view, offset, counters=None: word_lookup(eojeol, view, offset, False),
}


def sentence_edges(sent, view, counters=None, lookup='morpheme', flatten=False):
    """Dictionary edges of a sentence in the order `sentence_lookup` produces them, without the
    BOS/EOS sentinels — dictionary/lookup.py:51-62.  Eojeols are `sent.split()`; offsets count
    syllables of the preceding eojeols.  `lookup` names the EojeolLookup class (LOOKUPS)."""
    fn = LOOKUPS[lookup]
    edges = []
    offset = 0
    for eojeol in sent.split():
        found = fn(eojeol, view, offset, counters)
        edges += flatten_edges(found) if flatten else found
        offset += len(eojeol)
    if counters is not None:
        counters.E += len(edges)
    return edges


def lattice_graph(sent, edges):
    """sentence_lookup_as_graph — dictionary/lookup.py:281-342: (nodes, links) with nodes =
    [BOS] + edges + [EOS] and links [from, to, 0] between a word and the words that begin at the
    closest non-empty begin index at or after its end.  IndexError without any edge, as there
    (`bindex` is `[]`, :362-363, and :326 indexes it)."""
    n = len(sent.replace(' ', ''))
    bos = (BOS, BOS, None, BOS, None, 0, 0, 0, False)
    eos = (EOS, EOS, None, EOS, None, 0, n, n, False)
    bindex = begin_index(sent, edges)

    def closest(begin):
        for i in range(begin, n):
            if bindex[i]:
                return i
        return -1

    links = [[bos, word, 0] for word in bindex[closest(0)]]
    for bucket in bindex:
        for src in bucket:
            nxt = closest(src[E])
            if nxt == -1:
                links.append([src, eos, 0])
            else:
                for dst in bindex[nxt]:
                    if src[LEN] == 0 and dst[LEN] == 0:
                        continue
                    links.append([src, dst, 0])
    return [bos] + list(edges) + [eos], links


def begin_index(sent, edges):
    """bindex of sentence_lookup_as_begin_index — dictionary/lookup.py:357-369: `[]` when the
    sentence has no dictionary edge at all, else one bucket per syllable keyed by edge.b."""
    if not edges:
        return []
    buckets = [[] for _ in range(len(sent.replace(' ', '')))]
    for edge in edges:
        buckets[edge[B]].append(edge)
    return buckets


# --------------------------------------------------------------------------------------------
# scoring
# --------------------------------------------------------------------------------------------

def feature_tuples(word_i, word_j, word_k):
    """The trigram templates — features/feature.py:94-121."""
    tk, tj = word_k[TAG0], word_j[TAG0]
    out = [(0, word_j[WORD], word_k[WORD], tk), (1, word_j[WORD], tk), (2, tj, word_k[WORD], tk),
           (3, tj, tk), (4, word_k[LEN]), (5, word_k[WORD], tk, word_k[IS_L])]
    if tj == UNK:
        out.append((6, min(8, word_j[LEN])))
    if word_i is not None:
        out.append((7, word_i[WORD], word_j[WORD], word_k[WORD]))
    if tk in CONTEXTUAL:
        if tj in CONTEXTUAL:
            out.append((8, word_j[MORPH0], word_k[MORPH0]))
        elif word_i is not None and word_i[TAG0] in CONTEXTUAL:
            out.append((8, word_i[MORPH0], word_k[MORPH0]))
    return out


def numpy_order_sum(values):
    """`ndarray.sum()` of a short contiguous fp64 vector, association spelled out (Q6, SURVEY
    §8c): fewer than 8 terms left to right from 0.0; from 8 terms on, eight running lanes
    combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and the remainder added left to right.
    At most 9 terms occur.  `tests/test_oracle_golden.py` re-asserts this against numpy."""
    n = len(values)
    if n < 8:
        total = 0.0
        for v in values:
            total += v
        return total
    lanes = list(values[:8])
    i = 8
    while i + 8 <= n:
        for j in range(8):
            lanes[j] += values[i + j]
        i += 8
    total = ((lanes[0] + lanes[1]) + (lanes[2] + lanes[3])) + ((lanes[4] + lanes[5]) + (lanes[6] + lanes[7]))
    for v in values[i:]:
        total += v
    return total


class ScoreProgram:
    """Ordered list of additive scorers, read off a `BeamScoreFunctions`-like object by attribute
    (works for the reference's classes and for `lattice_based_tagger_b200.beam`'s descriptors).
    Evaluation order and the running `score += f(...)` follow beam/score_funcs.py:50-54."""

    def __init__(self, score_funcs):
        self.steps = []
        for func in score_funcs.funcs:
            kind = type(func).__name__
            if kind == 'RegularizationScore':
                self.steps.append(('reg', (func.unknown_penalty, func.known_preference, func.syllable_penalty)))
            elif kind == 'MorphemePreferenceScore':
                self.steps.append(('mpref', func.tag_to_morph))
            elif kind == 'WordPreferenceScore':
                self.steps.append(('wpref', func.tag_to_word))
            elif kind == 'SimpleTrigramFeatureScore':
                if func.encoder is None:
                    raise AttributeError("'NoneType' object has no attribute 'encode_word'")
                coefficients = func.coefficients
                coefficients = coefficients.tolist() if hasattr(coefficients, 'tolist') else list(coefficients)
                self.steps.append(('trigram', (func.encoder.feature_dic, coefficients)))
            else:
                raise ValueError('unsupported score function %s' % kind)

    def increment(self, word_i, word_j, word_k, counters=None):
        score = 0
        for kind, arg in self.steps:
            if kind == 'reg':                                  # beam/score_funcs.py:65-73
                unknown_penalty, known_preference, syllable_penalty = arg
                value = 0
                if word_k[TAG0] == UNK:
                    value += unknown_penalty * (word_k[LEN] + 0.1)
                else:
                    value += known_preference * word_k[LEN]
                if word_k[LEN] == 1 and word_k[TAG0] == NOUN:
                    value += syllable_penalty
            elif kind == 'mpref':                              # :84-88
                value = arg.get(word_k[TAG0], {}).get(word_k[MORPH0], 0)
                if word_k[TAG1] is not None:
                    value += arg.get(word_k[TAG1], {}).get(word_k[MORPH1], 0)
            elif kind == 'wpref':                              # :99-100
                value = arg.get(word_k[TAG0], {}).get(word_k[WORD], 0)
            else:                                              # :137-144 + features/feature.py:28,57-60
                feature_dic, coefficients = arg
                tuples = feature_tuples(word_i, word_j, word_k)
                if counters is not None:
                    counters.F += len(tuples)
                picked = [coefficients[feature_dic[f]] for f in tuples if f in feature_dic]
                value = numpy_order_sum(picked) if picked else 0
            score += value
        return score


# --------------------------------------------------------------------------------------------
# beam search
# --------------------------------------------------------------------------------------------

class Hypothesis:
    """beam/beam.py:88-116 — word list, running score, count of trailing unknown words."""
    __slots__ = ('words', 'score', 'num_unk')

    def __init__(self, words, score, num_unk=0):
        self.words = words
        self.score = score
        self.num_unk = num_unk

    def extended(self, edge, increment):
        num_unk = self.num_unk + 1 if edge[TAG0] == UNK else 0
        return Hypothesis(self.words + [edge], self.score + increment, num_unk)


def beam_search(bindex, chars, program, beam_size=5, counters=None):
    """Left-to-right beam over end positions — beam/beam.py:5-61.

    For end e the begin b runs over the last `WINDOW` syllables; a (b, e) span without a
    dictionary edge gets one unknown word; an unknown word may not follow an unknown word except
    from the earliest begin of the window (:44-45); survivors are the `beam_size` best by a
    STABLE sort on -score (:85), so generation order breaks ties (Q7).  EOS adds 0 (:59-61)."""
    length = len(chars)
    bos = (BOS, BOS, None, BOS, None, 0, 0, 0, False)
    eos = (EOS, EOS, None, EOS, None, 0, length, length, False)
    beams = [[Hypothesis([bos], 0)]]
    for e in range(1, length + 1):
        grown = []
        b_min = max(0, e - WINDOW)
        for b in range(b_min, e):
            expansions = [edge for edge in bindex[b] if edge[E] == e]     # IndexError if bindex == []
            if not expansions:
                sub = chars[b:e]
                expansions = [(sub, sub, None, UNK, None, e - b, b, e, False)]
            for hyp in beams[b]:
                for edge in expansions:
                    if hyp.num_unk > 0 and edge[TAG0] == UNK and b_min < b:
                        continue
                    word_i = None if len(hyp.words) == 1 else hyp.words[-2]
                    increment = program.increment(word_i, hyp.words[-1], edge, counters)
                    if counters is not None:
                        counters.T += 1
                    grown.append(hyp.extended(edge, increment))
        kept = sorted(grown, key=lambda h: -h.score)[:beam_size]
        if counters is not None:
            counters.Bk += len(kept)
        beams.append(kept)
    return [h.extended(eos, 0) for h in beams[-1]]


class OracleTagger:
    """`Tagger(dictionary, score_funcs=...)` / `.tag(sent, beam_size)` — tagger/tagger.py:47-78."""

    def __init__(self, dictionary, score_funcs, k3_first=None, lookup='morpheme'):
        self.view = DictView(dictionary, k3_first)
        self.program = ScoreProgram(score_funcs) if score_funcs is not None else None
        self.lookup = lookup       # the reference's Tagger always uses MorphemeLookup (tagger.py:60)

    def lattice(self, sent, counters=None, flatten=False):
        return sentence_edges(sent, self.view, counters, self.lookup, flatten)

    def survivors(self, sent, beam_size=5, counters=None):
        chars = sent.replace(' ', '')
        if counters is not None:
            counters.L += len(chars)
        edges = sentence_edges(sent, self.view, counters, self.lookup)
        return beam_search(begin_index(sent, edges), chars, self.program, beam_size, counters)

    def tag(self, sent, beam_size=5, counters=None):
        best = self.survivors(sent, beam_size, counters)[0]
        if counters is not None:
            counters.W += len(best.words) - 2
        return best
