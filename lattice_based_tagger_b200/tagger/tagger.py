"""`Tagger` — the drop-in boundary of the decode path.

Same constructor and `tag()` signature as the reference (`tagger/tagger.py:47-48,68`); `tag()`
returns a `Sequence` whose `.sequences` is the word list BOS .. EOS, `.score` the path score and
`.num_unk` the trailing-unknown count.  Added for the GPU: `tag_batch()` (many sentences per
call — the unit the kernels are built for), `tag_batch_kbest()` (every beam survivor, the
return value of the reference's `beam_search`) and `lattice_batch()` (the lattice alone, i.e.
`sentence_lookup_as_begin_index`, `dictionary/lookup.py:344-369`).

All work happens in `liblt_b200.so`: the sentences go to the device as raw UTF-16 text, the
lattice and beam kernels run there, and packed 16-byte word records come back.  Nothing is
computed on the host besides packing strings and rebuilding `Word` tuples; without the library
or a CUDA device the constructor raises.
"""

from .. import _native
from ..dictionary import BaseMorphemeDictionary, Word
from ..dictionary.lookup import EojeolLookup, MorphemeLookup, begin_index
from ..engine import Engine, PackedSequences, pack_sentences  # noqa: F401  (pack_sentences is part of this module's surface)
from ..tagset import BOS, EOS


class Tagger:
    """
    >>> funcs = BeamScoreFunctions(RegularizationScore(unknown_penalty=-.1, known_preference=0.5))
    >>> tagger = Tagger(DemoMorphemeDictionary(), score_funcs=funcs)
    >>> tagger.tag('너무너무너무는 아이오아이의 노래 입니다').score
    15.5

    `lookup`: the reference accepts the argument and then always builds `MorphemeLookup`
    (`tagger.py:57-62`); any string does the same here.  An `EojeolLookup` object — `LRLookup`,
    `WordLookup`, `MorphemeLookup`, with their `prefer_exact_match` / `flatten` options — selects
    that enumeration on the device instead (the options the argument was meant for).
    """

    def __init__(self, dictionary='base', lookup='subword_lookup', encoder=None, score_funcs=None,
                 device=0, k3_first=None):
        if isinstance(dictionary, str):
            dictionary = BaseMorphemeDictionary()
        self.dictionary = dictionary
        self.score_funcs = score_funcs
        self.device = device
        self._k3_first = k3_first
        self._lookup_arg = lookup
        self._engine = None
        self.eojeol_lookup = None
        self.refresh()

    # -- device state ------------------------------------------------------------------------------
    def refresh(self):
        """(Re)compile the device tables — call after mutating the dictionary or the weights."""
        self.close()
        self._engine = Engine(self.dictionary, self.score_funcs, self.device, self._k3_first)
        if isinstance(self._lookup_arg, EojeolLookup):
            self.eojeol_lookup = self._lookup_arg
        else:
            self.eojeol_lookup = MorphemeLookup(self.dictionary, device=self.device)
        # the lookup object shares this tagger's tables (its own would be a second copy of the dictionary)
        self.eojeol_lookup._attach(self._engine, self._engine.tables.max_len)
        self._engine.set_lookup(self.eojeol_lookup.mode)

    def update_weights(self):
        """After changing the `coefficients` of the tagger's trigram scorers (same `feature_dic`): write them
        to the device tables in place — much cheaper than `refresh()`, which recompiles everything."""
        self._engine.tables.update_weights()

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # handles the C-ABI level tools (bench.py, tests) use directly
    @property
    def _lib(self):
        return self._engine._lib

    @property
    def _batch(self):
        return self._engine.batch

    @property
    def _tables(self):
        return self._engine.tables

    # -- the reference's API ---------------------------------------------------------------------------
    def tag(self, sent, beam_size=5, ensure_normalize=True, debug=False):
        if debug:
            self.trace(sent, beam_size)
        return self.tag_batch([sent], beam_size=beam_size)[0]

    # -- batched API -----------------------------------------------------------------------------------
    def tag_batch(self, sents, beam_size=5, errors='raise', lazy=True):
        """Tag many sentences in one device pass; returns one `Sequence` per sentence.

        A sentence the reference cannot tag — no dictionary edge at all, `IndexError` there
        (`lookup.py:362-363` + `beam.py:33`) — raises the same here, or yields `None` with
        `errors='none'`; so do the inputs this implementation rejects per sentence (whitespace
        other than U+0020, characters outside the BMP, sentences beyond the device limit:
        `ValueError`).  The result is a `PackedSequences`: a list whose `Sequence` / `Word` objects
        are built when an item is read (`lazy=False`: a plain list, all built at once).
        """
        if self.score_funcs is None:
            raise TypeError("'NoneType' object is not callable")      # what beam.py:47 raises
        sents = list(sents)
        if self.eojeol_lookup.flatten:
            best = [None if s is None else s[0] for s in self._kbest_flattened(sents, beam_size, errors)]
            return best
        packed = self._engine.tag_packed(sents, beam_size)
        if errors == 'raise':
            status = packed[3]
            if status.any():
                bad = int(status.nonzero()[0][0])
                self._engine._raise_for(int(status[bad]), bad)
        if not lazy:
            return self._engine.unpack(sents, packed, errors)
        return PackedSequences(self._engine, sents, packed, errors)

    def tag_batch_packed(self, sents, beam_size=5):
        """The C-ABI call alone: returns (path_off, path_edges, scores, status) numpy arrays."""
        return self._engine.tag_packed(list(sents), beam_size)

    def tag_batch_kbest(self, sents, beam_size=5, errors='raise'):
        """Every survivor of the beam per sentence, best first — what the reference's `beam_search`
        returns (`beam/beam.py:59-61`); `tag_batch(...)[i]` is `tag_batch_kbest(...)[i][0]`."""
        if self.score_funcs is None:
            raise TypeError("'NoneType' object is not callable")
        sents = list(sents)
        if self.eojeol_lookup.flatten:
            return self._kbest_flattened(sents, beam_size, errors)
        packed = self._engine.kbest_packed(sents, beam_size)
        return self._engine.unpack_kbest(sents, packed, beam_size, errors)

    def _kbest_flattened(self, sents, beam_size, errors):
        # flatten=True (flatten_words, dictionary.py:114-167): the device builds the lattice, the two-morpheme
        # words are split on the host, and the flattened lattice goes back as an imported one
        lattices = self.eojeol_lookup.lookup_batch(sents, errors=errors)
        usable = [(i, s, lat) for i, (s, lat) in enumerate(zip(sents, lattices)) if lat is not None]
        packed = self._engine.kbest_packed([s for _, s, _ in usable], beam_size, imported=[lat for _, _, lat in usable])
        got = self._engine.unpack_kbest([s for _, s, _ in usable], packed, beam_size, errors, self._engine._imported_words)
        out = [None] * len(sents)
        for (i, _, _), seqs in zip(usable, got):
            out[i] = seqs
        return out

    def trace(self, sent, beam_size=5):
        """`tag(debug=True)`: prints the hypotheses kept at every end position, best first.  (The
        reference prints every grown hypothesis before the beam is cut, `beam/beam.py:53-57`; the
        device keeps only the survivors, so those are what can be shown.)  The beam at end position
        e depends on nothing to its right, so it is the search result of the e-syllable prefix of
        the sentence's lattice — one k-best pass over all prefixes."""
        words = self.eojeol_lookup.lookup_batch([sent])[0]
        chars = sent.replace(' ', '')
        prefixes, lattices = [], []
        for e in range(1, len(chars) + 1):
            prefixes.append(chars[:e])
            lattices.append([w for w in words if w.e <= e])
        if not prefixes:
            return
        packed = self._engine.kbest_packed(prefixes, beam_size, imported=lattices)
        kept = self._engine.unpack_kbest(prefixes, packed, beam_size, 'none', self._engine._imported_words)
        for e, seqs in enumerate(kept, 1):
            print('\n{}\nEnd point = {}, len(kept) = {}\n'.format('-' * 40, e, len(seqs or ())))
            for seq in seqs or ():
                seq.sequences = seq.sequences[:-1]        # the hypotheses of an inner position carry no EOS
                seq.num_unk = 0
                for w in reversed(seq.sequences):
                    if w.tag0 != 'Unknown':
                        break
                    seq.num_unk += 1
                print(seq, end='\n\n')

    def unpack(self, sents, packed, errors='raise'):
        return self._engine.unpack(sents, packed, errors)

    def edges_to_words(self, chars, edges):
        """Packed `lt_edge` records -> `Word` tuples (include/lt_b200.h documents the encoding)."""
        return self._engine.records_to_words(chars, edges.tolist())

    def lattice_batch(self, sents):
        """`sentence_lookup_as_begin_index` for every sentence: list of (words, bindex).

        `words` = [BOS] + dictionary edges + [EOS]; edges come grouped by end position then begin
        position, each (begin, end) group in the reference's order (the order `beam_search`
        observes); `bindex` is `[]` for a sentence without any dictionary edge
        (`lookup.py:362-363`).
        """
        sents = list(sents)
        out = []
        for sent, real in zip(sents, self.eojeol_lookup.lookup_batch(sents)):
            m = len(sent.replace(' ', ''))
            words = [Word(BOS, BOS, None, BOS, None, 0, 0, 0, False)] + real
            words.append(Word(EOS, EOS, None, EOS, None, 0, m, m, False))
            out.append((words, begin_index(m, real)))
        return out

    def counters(self):
        """Work counters of the last batch: L, P, E, T, F, Bk, W."""
        return self._engine.counters()

    def timings(self):
        """Device times of the last batch by stage (first call only switches timing on)."""
        return self._engine.timings()

    def set_stage_timing(self, on):
        """Per-stage events on / off for the batches to come; off lets the kernels of a batch overlap their
        launch prologues with their predecessors' tails (programmatic dependent launch)."""
        self._engine.set_stage_timing(on)

    def info(self):
        """Workspace state: buffer capacities, reruns, kernel launches so far, launch shapes."""
        return self._engine.info()
