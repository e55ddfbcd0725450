"""`Tagger` — the drop-in boundary of the decode path.

Same constructor and `tag()` signature as the reference (`tagger/tagger.py:47-48,68`); `tag()`
returns a `Sequence` whose `.sequences` is the word list BOS .. EOS, `.score` the path score and
`.num_unk` the trailing-unknown count.  Added for the GPU: `tag_batch()` (many sentences per
call — the unit the kernels are built for) and `lattice_batch()` (the lattice alone, i.e.
`sentence_lookup_as_begin_index`, `dictionary/lookup.py:344-369`).

All work happens in `liblt_b200.so`: the sentences go to the device as raw UTF-16 text, the
lattice and beam kernels run there, and packed 16-byte word records come back.  Nothing is
computed on the host besides packing strings and rebuilding `Word` tuples; without the library
or a CUDA device the constructor raises.
"""

import ctypes

import numpy as np

from .. import _native
from ..beam import Sequence
from ..compile import CompiledTables
from ..dictionary import BaseMorphemeDictionary, Word
from ..tagset import BOS, EOS


_BOS_WORD = Word(BOS, BOS, None, BOS, None, 0, 0, 0, False)


class MorphemeLookup:
    """Descriptor of the eojeol lookup the tagger uses (reference `MorphemeLookup`,
    `dictionary/lookup.py:99-132`): `prefer_exact_match=True`, the default stand-alone tags and
    `max_len` derived from the dictionary.  The enumeration itself is `csrc/lattice.cuh`."""

    def __init__(self, dictionary, max_len, flatten=False):
        self.dictionary = dictionary
        self.prefer_exact_match = True
        self.standalones = ['Noun', 'Adverb', 'Exclamation', 'Determiner', 'Number']
        self.max_len = max_len
        self.flatten = flatten


def pack_sentences(sents):
    """list[str] -> (uint16 text, int32 offsets); spaces stay in the text (the kernels strip them)."""
    n = len(sents)
    lengths = np.fromiter((len(s) for s in sents), dtype=np.int64, count=n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    if offsets[-1] >= 2 ** 31:
        raise ValueError('batch holds %d code units; split it (limit 2^31)' % offsets[-1])
    raw = ''.join(sents).encode('utf-16-le', 'surrogatepass')
    text = np.frombuffer(raw, dtype='<u2')
    if text.size != int(offsets[-1]):
        bad = next(i for i, s in enumerate(sents) if len(s.encode('utf-16-le', 'surrogatepass')) != 2 * len(s))
        raise ValueError('sentence %d contains characters outside the Basic Multilingual Plane, '
                         'which the device text format (UTF-16 code units) does not support' % bad)
    if text.size == 0:
        text = np.zeros(1, dtype='<u2')
    return np.ascontiguousarray(text), offsets.astype(np.int32)


class Tagger:
    """
    >>> funcs = BeamScoreFunctions(RegularizationScore(unknown_penalty=-.1, known_preference=0.5))
    >>> tagger = Tagger(DemoMorphemeDictionary(), score_funcs=funcs)
    >>> tagger.tag('너무너무너무는 아이오아이의 노래 입니다').score
    15.5
    """

    def __init__(self, dictionary='base', lookup='subword_lookup', encoder=None, score_funcs=None,
                 device=0, k3_first=None):
        if isinstance(dictionary, str):
            dictionary = BaseMorphemeDictionary()
        self.dictionary = dictionary
        self.score_funcs = score_funcs
        self.device = device
        self._k3_first = k3_first
        self._lib = _native.load()
        self._tables = None
        self._batch = None
        self.eojeol_lookup = None
        self.refresh()

    # -- device state ------------------------------------------------------------------------------
    def refresh(self):
        """(Re)compile the device tables — call after mutating the dictionary or the weights."""
        self.close()
        self._tables = CompiledTables(self.dictionary, self.score_funcs, self.device, self._k3_first)
        self.eojeol_lookup = MorphemeLookup(self.dictionary, self._tables.max_len)
        handle = ctypes.c_void_p()
        _native.check(self._lib.lt_batch_create(self._tables.handle, ctypes.byref(handle)))
        self._batch = handle

    def close(self):
        if self._batch is not None:
            self._lib.lt_batch_destroy(self._batch)
            self._batch = None
        if self._tables is not None:
            self._tables.close()
            self._tables = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's API ---------------------------------------------------------------------------
    def tag(self, sent, beam_size=5, ensure_normalize=True, debug=False):
        return self.tag_batch([sent], beam_size=beam_size)[0]

    # -- batched API -----------------------------------------------------------------------------------
    def tag_batch(self, sents, beam_size=5, errors='raise'):
        """Tag many sentences in one device pass; returns one `Sequence` per sentence.

        A sentence the reference cannot tag — no dictionary edge at all, `IndexError` there
        (`lookup.py:362-363` + `beam.py:33`) — raises the same here, or yields `None` with
        `errors='none'`.
        """
        if self.score_funcs is None:
            raise TypeError("'NoneType' object is not callable")      # what beam.py:47 raises
        sents = list(sents)
        packed = self.tag_batch_packed(sents, beam_size)
        return self.unpack(sents, packed, errors)

    def tag_batch_packed(self, sents, beam_size=5):
        """The C-ABI call alone: returns (path_off, path_edges, scores, status) numpy arrays."""
        if not 1 <= beam_size <= _native.LT_MAX_BEAM:
            raise ValueError('beam_size must be in 1..%d' % _native.LT_MAX_BEAM)
        text, offsets = pack_sentences(sents)
        n = len(sents)
        cap = max(1, int(offsets[-1]))
        path_off = np.zeros(n + 1, dtype=np.int32)
        path_edges = np.zeros(cap, dtype=_native.EDGE_DTYPE)
        scores = np.zeros(max(1, n), dtype=np.float64)
        status = np.zeros(max(1, n), dtype=np.int32)
        _native.check(self._lib.lt_tag_batch_host(
            self._batch, _native.ptr(text), _native.ptr(offsets), n, int(beam_size),
            _native.ptr(path_off), _native.ptr(path_edges), cap, _native.ptr(scores), _native.ptr(status)))
        return path_off, path_edges[:int(path_off[n])], scores[:n], status[:n]

    def unpack(self, sents, packed, errors='raise'):
        path_off, path_edges, scores, status = packed
        # one conversion for the whole batch (per-sentence numpy slicing dominates otherwise)
        offs = path_off.tolist()
        records = path_edges.tolist()
        score_list = scores.tolist()
        status_list = status.tolist()
        out = []
        for i, sent in enumerate(sents):
            st = status_list[i]
            if st != _native.LT_SENT_OK:
                if errors == 'raise':
                    if st == _native.LT_SENT_NO_EDGES:
                        raise IndexError('list index out of range')   # as the reference does
                    raise ValueError('sentence %d contains whitespace other than U+0020' % i)
                out.append(None)
                continue
            chars = sent.replace(' ', '')
            n = len(chars)
            words = [_BOS_WORD]
            words += self._records_to_words(chars, records[offs[i]:offs[i + 1]])
            words.append(Word(EOS, EOS, None, EOS, None, 0, n, n, False))
            # adding EOS resets the trailing-unknown count (beam.py:113 with tag0 == EOS)
            out.append(Sequence(words, score_list[i], 0))
        return out

    def edges_to_words(self, chars, edges):
        """Packed `lt_edge` records -> `Word` tuples (include/lt_b200.h documents the encoding)."""
        return self._records_to_words(chars, edges.tolist())

    def _records_to_words(self, chars, records):
        names = self._tables.tag_names
        rules = self._tables.rules_flat
        lemma_flag, is_l_flag = _native.LT_EDGE_LEMMA, _native.LT_EDGE_IS_L
        new = tuple.__new__            # Word is a namedtuple: skips the per-call length check of _make
        words = []
        for b, e, length, tag0, tag1, rule, split, flags, _ in records:
            surface = chars[b:e]
            is_l = (flags & is_l_flag) != 0
            if flags & lemma_flag:
                if rule == _native.LT_NO_RULE:
                    morph0, morph1 = surface[:split + 1], surface[split + 1:]
                else:
                    stem, eomi = rules[rule]
                    skip = 2 if flags & _native.LT_EDGE_SKIP2 else 1
                    morph0, morph1 = surface[:split] + stem, eomi + surface[split + skip:]
                words.append(new(Word, (surface, morph0, morph1, names[tag0], names[tag1], length, b, e, is_l)))
            else:
                words.append(new(Word, (surface, surface, None, names[tag0], None, length, b, e, is_l)))
        return words

    def lattice_batch(self, sents):
        """`sentence_lookup_as_begin_index` for every sentence: list of (words, bindex).

        `words` = [BOS] + dictionary edges + [EOS]; edges come grouped by end position then begin
        position, each (begin, end) group in the reference's order (the order `beam_search`
        observes); `bindex` is `[]` for a sentence without any dictionary edge
        (`lookup.py:362-363`).
        """
        sents = list(sents)
        text, offsets = pack_sentences(sents)
        n = len(sents)
        n_units = int(offsets[-1])
        _native.check(self._lib.lt_lattice_host(self._batch, _native.ptr(text), _native.ptr(offsets), n))
        n_edges = ctypes.c_int64()
        _native.check(self._lib.lt_lattice_size(self._batch, ctypes.byref(n_edges)))
        edges = np.zeros(max(1, n_edges.value), dtype=_native.EDGE_DTYPE)
        end_off = np.zeros(n_units + 1, dtype=np.int64)
        _native.check(self._lib.lt_lattice_fetch(self._batch, _native.ptr(edges), edges.size, _native.ptr(end_off)))
        out = []
        for i, sent in enumerate(sents):
            chars = sent.replace(' ', '')
            lo, hi = int(end_off[offsets[i]]), int(end_off[offsets[i + 1]])
            real = self.edges_to_words(chars, edges[lo:hi])
            m = len(chars)
            words = [Word(BOS, BOS, None, BOS, None, 0, 0, 0, False)] + real
            words.append(Word(EOS, EOS, None, EOS, None, 0, m, m, False))
            if not real:
                out.append((words, []))
                continue
            bindex = [[] for _ in range(m)]
            for w in real:
                bindex[w.b].append(w)
            out.append((words, bindex))
        return out

    def counters(self):
        """Work counters of the last batch: L, P, E, T, F, Bk, W."""
        c = _native.lt_counters()
        _native.check(self._lib.lt_batch_counters(self._batch, ctypes.byref(c)))
        return c.as_dict()

    def timings(self):
        """Device times of the last batch by stage (first call only switches timing on)."""
        t = _native.lt_timings()
        _native.check(self._lib.lt_batch_timings(self._batch, ctypes.byref(t)))
        return t.as_dict()
