"""Beam search entry points and result objects (host side of `csrc/beam.cuh`).

`Sequence` is what `Tagger.tag` returns: the word list BOS .. EOS, the path score and the number
of trailing unknown words — the same three attributes as the reference's `Sequence`
(`beam/beam.py:88-124`).  `beam_search` has the reference's signature (`beam/beam.py:5`) and
return value (every survivor of the last position, EOS appended, best first); the search runs in
the device kernel on the caller's lattice, which travels as an imported lattice
(`lt_lattice_import`).  `Beam` is the reference's container of per-position hypothesis lists
(`beam/beam.py:63-86`), kept for code that builds beams by hand; the device keeps its own ring of
back-pointer entries instead.
"""

from ..tagset import BUILTIN_TAGS


class Sequence:
    def __init__(self, sequences, score, num_unk=0):
        self.sequences = sequences
        self.score = score
        self.num_unk = num_unk

    def __repr__(self):
        words = '[\n    {}\n  ]'.format('\n    '.join(str(w) for w in self.sequences))
        return 'Sequences(\n  words : {}\n  score : {}\n  num unks in tails : {}\n)'.format(
            words, self.score, self.num_unk)

    __str__ = __repr__

    def __eq__(self, other):
        if not isinstance(other, Sequence):
            return NotImplemented
        return self.sequences == other.sequences and self.score == other.score and self.num_unk == other.num_unk

    __hash__ = None


class Beam:
    """Hypothesis lists indexed by end position; `append` keeps the `k` best of a position by a
    stable sort on the score (reference `beam/beam.py:63-86`)."""

    def __init__(self, beam=None, k=5):
        self.k = k
        self.beam = beam if beam is not None else []

    def __getitem__(self, index):
        return self.beam[index]

    def __len__(self):
        return len(self.beam)

    def append(self, candidates):
        self.beam.append(sorted(candidates, key=lambda x: -x.score)[:self.k])


# engines of beam_search, one per score-function object (device tables are built from the scorers alone)
_ENGINES = {}
_MAX_ENGINES = 4


def _engine_for(score_functions, device):
    from ..compile import OTHER_TAG
    from ..dictionary.dictionary import MorphemeDictionary
    from ..engine import Engine
    key = (id(score_functions), device)
    hit = _ENGINES.get(key)
    if hit is not None and hit[0] is score_functions:
        return hit[1]
    # tags the scorers name that tagset.py does not know: the second component of (1, word, tag) / (5, word, tag, l)
    # features, the tag components of templates 2 and 3, the keys of the preference tables
    extra = []
    for func in getattr(score_functions, 'funcs', ()):
        encoder = getattr(func, 'encoder', None)
        for f in getattr(encoder, 'feature_dic', None) or ():
            if not isinstance(f, tuple) or not f:
                continue
            if f[0] == 0 and len(f) == 4:
                extra.append(f[3])
            elif f[0] == 1 and len(f) == 3:
                extra.append(f[2])
            elif f[0] == 2 and len(f) == 4:
                extra += [f[1], f[3]]
            elif f[0] == 3 and len(f) == 3:
                extra += [f[1], f[2]]
            elif f[0] == 5 and len(f) == 4:
                extra.append(f[2])
        for table in (getattr(func, 'tag_to_morph', None), getattr(func, 'tag_to_word', None)):
            extra += list(table or ())
    known = []
    for t in extra:
        if isinstance(t, str) and t not in BUILTIN_TAGS and t not in known:
            known.append(t)
    engine = Engine(MorphemeDictionary({}, {}), score_functions, device, extra_tags=list(known) + [OTHER_TAG])
    while len(_ENGINES) >= _MAX_ENGINES:
        _ENGINES.pop(next(iter(_ENGINES)))[1].close()
    _ENGINES[key] = (score_functions, engine)
    return engine


def beam_search(bindex, chars, score_functions, beam_size=5, max_len=8, debug=False, device=0):
    """The reference's `beam_search` (`beam/beam.py:5-61`) on the device: `bindex[b]` lists the words
    that begin at syllable b of `chars`; returns every survivor (a `Sequence` with EOS appended), best
    first.  `bindex == []` raises `IndexError` for a non-empty `chars`, as there.

    `max_len` is the window of the device kernel and cannot be changed (`Tagger.tag` cannot change it
    either, `tagger.py:75-76`).
    """
    if max_len != _WINDOW:
        raise ValueError('the device beam search has a fixed window of %d syllables (max_len)' % _WINDOW)
    if ' ' in chars:
        raise ValueError('`chars` is the space-stripped sentence')
    if score_functions is None:
        raise TypeError("'NoneType' object is not callable")
    if len(chars) > 0 and len(bindex) == 0:
        raise IndexError('list index out of range')
    words = [w for bucket in bindex for w in bucket]
    engine = _engine_for(score_functions, device)
    other = engine.tables.other_tag_id
    ids = engine.tables.tag_ids
    # a tag the scorers never mention behaves like any other unmentioned tag
    _ = [ids.setdefault(t, other) for w in words for t in (w.tag0, w.tag1) if t is not None and t not in ids]
    packed = engine.kbest_packed([chars], beam_size, imported=[words])
    return engine.unpack_kbest([chars], packed, beam_size, 'raise', engine._imported_words)[0]


_WINDOW = 8
