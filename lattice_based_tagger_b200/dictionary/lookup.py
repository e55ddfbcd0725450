"""Eojeol lookups and sentence lattices (host side of `csrc/lattice.cuh`).

The reference's lookup classes — `LRLookup`, `WordLookup`, `MorphemeLookup` and the functions
`sentence_lookup`, `sentence_lookup_as_begin_index`, `sentence_lookup_as_graph`
(`dictionary/lookup.py:7-369`) — with the same names, arguments and results.  The enumeration itself
runs on the device: an object here holds the options (`prefer_exact_match`, `flatten`, `max_len`)
that select a mode of the lattice kernel (`LT_LOOKUP_*`, include/lt_b200.h) and an `Engine` with the
dictionary's device tables, built on first use (or shared with the `Tagger` the object was given to).
One call looks one eojeol up; `lookup_batch` / `sentence_lookup_batch` are the calls the GPU is
built for.

Order of the returned words: the device emits edges grouped by (end, begin), each group in the
reference's order — the only order `beam_search` can observe.  `lookup_batch`
returns that order; `sentence_lookup*` and `lookup()` put the groups back into the reference's
enumeration order (`enumeration_order`, a permutation that follows from the spans alone).
"""

from .. import _native
from ..tagset import BOS, EOS
from .dictionary import Word
from .text import flatten_words


class EojeolLookup:
    """Base of the lookup classes (reference `lookup.py:64-73`)."""

    mode = _native.LT_LOOKUP_MORPHEME

    def __init__(self, dictionary, prefer_exact_match=True, flatten=False, device=0):
        if not hasattr(dictionary, 'rules'):
            raise ValueError('dictionary must be MorphemeDictionary')         # lookup.py:101-102
        self.dictionary = dictionary
        self.prefer_exact_match = prefer_exact_match
        self.flatten = flatten
        self.device = device
        self._engine = None
        self._owns_engine = False

    # -- device state --------------------------------------------------------------------------------
    def _attach(self, engine, max_len=None):
        self._release()
        self._engine = engine
        self._owns_engine = False

    def _release(self):
        if self._engine is not None and self._owns_engine:
            self._engine.close()
        self._engine = None

    def _ensure(self):
        if self._engine is None or self._engine.tables is None:
            from ..engine import Engine
            self._engine = Engine(self.dictionary, None, self.device)
            self._owns_engine = True
        self._engine.set_lookup(self.mode)
        return self._engine

    def refresh(self):
        """Recompile the device tables after the dictionary changed."""
        if self._owns_engine:
            self._release()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # -- the reference's API ---------------------------------------------------------------------------
    def __call__(self, eojeol, offset=0):
        return self.lookup(eojeol, offset)

    def lookup(self, eojeol, offset=0):
        """Words of one eojeol, `b` / `e` shifted by `offset`."""
        if ' ' in eojeol:
            raise ValueError('an eojeol holds no space')
        words = self.lookup_batch([eojeol], reference_order=True)[0]
        if offset:
            words = [w._replace(b=w.b + offset, e=w.e + offset) for w in words]
        return words

    # -- batched -----------------------------------------------------------------------------------
    def lookup_batch(self, sents, errors='raise', reference_order=None):
        """Dictionary words of every sentence (eojeols separated by spaces), without BOS / EOS.

        The device returns them grouped by (end, begin); `reference_order=True` puts them back into the
        order the reference's lookup enumerates them in (whole eojeol, then its splits / substrings).
        `flatten=True` needs that order — the split words land in other (begin, end) groups, and their
        order inside a group is the enumeration order — and therefore implies it."""
        sents = list(sents)
        found = self._ensure().lattice_words(sents, errors)
        if reference_order or (reference_order is None and self.flatten):
            found = [None if words is None else enumeration_order(sent, words, self.mode) for sent, words in zip(sents, found)]
        if self.flatten:
            found = [None if words is None else flatten_words(words) for words in found]
        return found


class MorphemeLookup(EojeolLookup):
    """reference `lookup.py:99-132`: whole eojeol and its left/right splits, else the sub-word scan
    over the stand-alone tags.  `standalones` and `max_len` are fixed to the reference's defaults
    (the device enumerates exactly what `Tagger.tag` uses, `tagger.py:60`)."""

    mode = _native.LT_LOOKUP_MORPHEME

    def __init__(self, dictionary, prefer_exact_match=True, standalones=None, max_len=-1, flatten=False, device=0):
        super().__init__(dictionary, prefer_exact_match, flatten, device)
        defaults = ['Noun', 'Adverb', 'Exclamation', 'Determiner', 'Number']
        if standalones is not None and list(standalones) != defaults:
            raise ValueError('the device lookup enumerates the default stand-alone tags %s only' % defaults)
        if not prefer_exact_match:
            raise ValueError('MorphemeLookup(prefer_exact_match=False) has no device implementation')
        self.standalones = defaults
        self._max_len_arg = max_len
        self.max_len = max_len if max_len > 0 else None       # derived from the dictionary on first use

    def _attach(self, engine, max_len=None):
        super()._attach(engine)
        self._check_max_len(max_len if max_len is not None else engine.tables.max_len)

    def _ensure(self):
        engine = super()._ensure()
        self._check_max_len(engine.tables.max_len)
        return engine

    def _check_max_len(self, derived):
        if self._max_len_arg > 0 and self._max_len_arg != derived:
            raise ValueError('MorphemeLookup(max_len=%d): the device scan uses the length derived from the '
                             'dictionary (%d)' % (self._max_len_arg, derived))
        self.max_len = derived


class LRLookup(EojeolLookup):
    """reference `lookup.py:75-85`, `lr_lookup` `:171-210`."""

    @property
    def mode(self):
        return _native.LT_LOOKUP_LR if self.prefer_exact_match else _native.LT_LOOKUP_LR_ALL


class WordLookup(EojeolLookup):
    """reference `lookup.py:87-97`, `word_lookup` `:134-169`."""

    @property
    def mode(self):
        return _native.LT_LOOKUP_WORD if self.prefer_exact_match else _native.LT_LOOKUP_WORD_ALL


class ExactLookup(EojeolLookup):
    """`MorphemeDictionary.lookup` of the whole string alone (reference `dictionary.py:304-312`)."""

    mode = _native.LT_LOOKUP_EXACT


def enumeration_order(sent, words, mode):
    """Device order (end, begin, order inside the span) -> the reference's enumeration order.

    Inside one (begin, end) span both orders agree; across spans the reference's order follows from
    the span alone.  Per eojeol [o, oe): `lr_lookup` / the first stage of `morpheme_lookup` list the
    whole eojeol, then for i = 1.. the left part [o, o+i) followed by the right part [o+i, oe)
    (`lookup.py:191-209`); the sub-word scan and `word_lookup`'s loop go by begin, then end
    (`lookup.py:259-277`, `:161-168`).  A first-stage result always holds a word that starts the
    eojeol, a scan result never does (it begins at 1), which tells the two apart; `word_lookup`'s
    initial whole-eojeol analyses come first and — without prefer_exact_match — once more inside the loop.
    """
    if not words:
        return words
    bounds = []
    o = 0
    for eojeol in sent.split():
        bounds.append((o, o + len(eojeol)))
        o += len(eojeol)
    by_eojeol = {bound: [] for bound in bounds}
    starts = [b for b, _ in bounds]
    import bisect
    for w in words:
        by_eojeol[bounds[bisect.bisect_right(starts, w.b) - 1]].append(w)
    word_mode = mode in (_native.LT_LOOKUP_WORD, _native.LT_LOOKUP_WORD_ALL)
    out = []
    for (o, oe) in bounds:
        group = by_eojeol[(o, oe)]
        if not group:
            continue
        if word_mode:
            whole = [w for w in group if w.b == o and w.e == oe]
            rest = [w for w in group if not (w.b == o and w.e == oe)]
            if mode == _native.LT_LOOKUP_WORD_ALL:
                first, again = whole[:len(whole) // 2], whole[len(whole) // 2:]
            else:
                first, again = (whole, []) if not rest else ([], whole)
            out += first
            out += sorted(rest + again, key=lambda w: (w.b, w.e))
        elif any(w.b == o for w in group):
            out += sorted(group, key=lambda w: 0 if (w.b == o and w.e == oe) else (2 * (w.e - o) - 1 if w.b == o else 2 * (w.b - o)))
        else:
            out += sorted(group, key=lambda w: (w.b, w.e))
    return out


def begin_index(n, words):
    """`bindex` of `sentence_lookup_as_begin_index` (`lookup.py:357-369`)."""
    if not words:
        return []
    bindex = [[] for _ in range(n)]
    for word in words:
        bindex[word.b].append(word)          # IndexError for a zero-width word at the sentence end, as there
    return bindex


def _with_sentinels(sent, words):
    n = len(sent.replace(' ', ''))
    return [Word(BOS, BOS, None, BOS, None, 0, 0, 0, False)] + list(words) + [Word(EOS, EOS, None, EOS, None, 0, n, n, False)]


def sentence_lookup(sent, eojeol_lookup):
    """[BOS] + dictionary words + [EOS], in the reference's order (reference `lookup.py:7-62`)."""
    return _with_sentinels(sent, eojeol_lookup.lookup_batch([sent], reference_order=True)[0])


def sentence_lookup_batch(sents, eojeol_lookup):
    sents = list(sents)
    return [_with_sentinels(s, w) for s, w in zip(sents, eojeol_lookup.lookup_batch(sents, reference_order=True))]


def sentence_lookup_as_begin_index(sent, eojeol_lookup):
    """(words, bindex); `bindex` is `[]` without any dictionary word (reference `lookup.py:344-369`)."""
    words = sentence_lookup(sent, eojeol_lookup)
    return words, begin_index(len(sent.replace(' ', '')), words[1:-1])


def sentence_lookup_as_graph(sent, eojeol_lookup):
    """(nodes, links): a word links to the words that begin at the closest non-empty begin index at
    or after its end, weight 0 (reference `lookup.py:281-342`).  A view of the device-built lattice;
    nothing is looked up here."""
    n = len(sent.replace(' ', ''))
    words, bindex = sentence_lookup_as_begin_index(sent, eojeol_lookup)

    def closest(begin):
        for i in range(begin, n):
            if bindex[i]:
                return i
        return -1

    bos, eos = words[0], words[-1]
    links = [[bos, word, 0] for word in bindex[closest(0)]]        # IndexError without any word, as there
    for bucket in bindex:
        for src in bucket:
            nxt = closest(src.e)
            if nxt == -1:
                links.append([src, eos, 0])
                continue
            for dst in bindex[nxt]:
                if src.len == 0 and dst.len == 0:
                    continue
                links.append([src, dst, 0])
    return words, links
