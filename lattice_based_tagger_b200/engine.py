"""`Engine` — one set of device tables plus a batch workspace, and the packing / unpacking between
Python strings / `Word` tuples and the C ABI's flat arrays (include/lt_b200.h).

Everything that computes runs in `liblt_b200.so`: an `Engine` only flattens its inputs, calls the
library and rebuilds `Word` / `Sequence` objects from the packed 16-byte records.  `Tagger`, the
eojeol-lookup classes (`dictionary/lookup.py`), `MorphemeDictionary.lookup / lemmatize`,
`analyze_morphology` and `beam_search` are thin layers over it.
"""

import ctypes
from collections.abc import Sequence as _SequenceABC

import numpy as np

from . import _native
from .compile import CompiledTables
from .tagset import BOS, EOS

_STATUS_MESSAGES = {
    _native.LT_SENT_BAD_SPACE: 'sentence %d contains whitespace other than U+0020',
    _native.LT_SENT_TOO_LONG: 'sentence %d is longer than the %d code units the device kernels hold',
    _native.LT_SENT_UNSUPPORTED_CHAR: 'sentence %d contains characters outside the Basic Multilingual Plane, '
                                      'which the device text format (UTF-16 code units) does not support',
}


def _word_class():
    from .dictionary.dictionary import Word
    return Word


def _sequence_class():
    from .beam.beam import Sequence
    return Sequence


def pack_sentences(sents, unsupported=None):
    """list[str] -> (uint16 text, int32 offsets); spaces stay in the text (the kernels strip them).

    A sentence with characters outside the BMP has no one-unit-per-character UTF-16 form: with
    `unsupported` (a list) its index is recorded there and it travels as an empty sentence, so that
    the rest of the batch is tagged; without it the call raises `ValueError`.
    """
    n = len(sents)
    raw = ''.join(sents).encode('utf-16-le', 'surrogatepass')
    lengths = np.fromiter(map(len, sents), dtype=np.int64, count=n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    if len(raw) != 2 * int(offsets[-1]):
        bad = [i for i, s in enumerate(sents) if len(s.encode('utf-16-le', 'surrogatepass')) != 2 * len(s)]
        if unsupported is None:
            raise ValueError(_STATUS_MESSAGES[_native.LT_SENT_UNSUPPORTED_CHAR] % bad[0])
        unsupported.extend(bad)
        sents = list(sents)
        for i in bad:
            sents[i] = ''
        raw = ''.join(sents).encode('utf-16-le', 'surrogatepass')
        lengths = np.fromiter(map(len, sents), dtype=np.int64, count=n)
        np.cumsum(lengths, out=offsets[1:])
    if offsets[-1] >= 2 ** 31:
        raise ValueError('batch holds %d code units; split it (limit 2^31)' % offsets[-1])
    text = np.frombuffer(raw, dtype='<u2')
    if text.size == 0:
        text = np.zeros(1, dtype='<u2')
    return np.ascontiguousarray(text), offsets.astype(np.int32)


class PackedSequences(_SequenceABC):
    """Result of `tag_batch`: behaves like `list[Sequence | None]`, but a sentence's `Sequence` (its
    `Word` tuples, BOS .. EOS) is only built when that item is read.  The packed arrays the device
    returned stay available as `.path_off`, `.path_edges`, `.scores`, `.status`."""

    def __init__(self, engine, sents, packed, errors):
        self._engine = engine
        self._sents = sents
        self.path_off, self.path_edges, self.scores, self.status = packed
        self._errors = errors
        self._cache = {}

    def __len__(self):
        return len(self._sents)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return [self[i] for i in range(*index.indices(len(self)))]
        n = len(self)
        if index < 0:
            index += n
        if not 0 <= index < n:
            raise IndexError('list index out of range')
        if index not in self._cache:
            self._cache[index] = self._engine.sequence_at(self._sents[index], self, index, self._errors)
        return self._cache[index]

    def __eq__(self, other):
        if isinstance(other, (list, tuple, PackedSequences)):
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        return NotImplemented

    def __repr__(self):
        return 'PackedSequences(%d sentences, %d words)' % (len(self), len(self.path_edges))

    def materialize(self):
        """All items as a plain list (one bulk conversion, faster than reading them one by one)."""
        return self._engine.unpack(self._sents, (self.path_off, self.path_edges, self.scores, self.status), self._errors)


class Engine:
    """Device tables of (dictionary, score functions) + one batch workspace."""

    def __init__(self, dictionary, score_funcs=None, device=0, k3_first=None, extra_tags=()):
        self._lib = _native.load()
        self.device = device
        self.tables = CompiledTables(dictionary, score_funcs, device, k3_first, extra_tags=extra_tags)
        handle = ctypes.c_void_p()
        _native.check(self._lib.lt_batch_create(self.tables.handle, ctypes.byref(handle)))
        self.batch = handle
        self._mode = _native.LT_LOOKUP_MORPHEME
        self.unit_limit = int(self._lib.lt_tables_max_sentence_units(self.tables.handle))

    def close(self):
        if getattr(self, 'batch', None) is not None:
            self._lib.lt_batch_destroy(self.batch)
            self.batch = None
        if getattr(self, 'tables', None) is not None:
            self.tables.close()
            self.tables = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- lookups ---------------------------------------------------------------------------------
    def set_lookup(self, mode):
        if mode != self._mode:
            _native.check(self._lib.lt_batch_set_lookup(self.batch, int(mode)))
            self._mode = mode

    # -- tagging -----------------------------------------------------------------------------------
    def tag_packed(self, sents, beam_size):
        """The C-ABI call alone: (path_off, path_edges, scores, status) numpy arrays."""
        if not 1 <= beam_size <= _native.LT_MAX_BEAM:
            raise ValueError('beam_size must be in 1..%d' % _native.LT_MAX_BEAM)
        unsupported = []
        text, offsets = pack_sentences(sents, unsupported)
        n = len(sents)
        cap = max(1, int(offsets[-1]))
        path_off = np.zeros(n + 1, dtype=np.int32)
        path_edges = np.empty(cap, dtype=_native.EDGE_DTYPE)
        scores = np.zeros(max(1, n), dtype=np.float64)
        status = np.zeros(max(1, n), dtype=np.int32)
        _native.check(self._lib.lt_tag_batch_host(
            self.batch, _native.ptr(text), _native.ptr(offsets), n, int(beam_size),
            _native.ptr(path_off), _native.ptr(path_edges), cap, _native.ptr(scores), _native.ptr(status)))
        for i in unsupported:
            status[i] = _native.LT_SENT_UNSUPPORTED_CHAR
        return path_off, path_edges[:int(path_off[n])], scores[:n], status[:n]

    def kbest_packed(self, sents, beam_size, imported=None):
        """All survivors: (n_best, path_off, path_edges, scores, status); `path_off` / `scores` are
        indexed by sentence * beam_size + rank.  With `imported` the lattice is the caller's."""
        if not 1 <= beam_size <= _native.LT_MAX_BEAM:
            raise ValueError('beam_size must be in 1..%d' % _native.LT_MAX_BEAM)
        n = len(sents)
        unsupported = []
        if imported is None:
            text, offsets = pack_sentences(sents, unsupported)
            _native.check(self._lib.lt_tag_batch_host_kbest(self.batch, _native.ptr(text), _native.ptr(offsets), n, int(beam_size)))
        else:
            self.import_lattice(sents, imported)
            _native.check(self._lib.lt_beam_kbest(self.batch, int(beam_size), None))
        total = ctypes.c_int64()
        _native.check(self._lib.lt_kbest_size(self.batch, ctypes.byref(total)))
        n_best = np.zeros(max(1, n), dtype=np.int32)
        path_off = np.zeros(n * beam_size + 1, dtype=np.int32)
        path_edges = np.empty(max(1, total.value), dtype=_native.EDGE_DTYPE)
        scores = np.zeros(max(1, n * beam_size), dtype=np.float64)
        status = np.zeros(max(1, n), dtype=np.int32)
        _native.check(self._lib.lt_kbest_fetch(self.batch, _native.ptr(n_best), _native.ptr(path_off), _native.ptr(path_edges),
                                               path_edges.size, _native.ptr(scores), _native.ptr(status)))
        for i in unsupported:
            status[i] = _native.LT_SENT_UNSUPPORTED_CHAR
            n_best[i] = 0
        return n_best[:n], path_off, path_edges[:total.value], scores[:n * beam_size], status[:n]

    # -- lattices ----------------------------------------------------------------------------------
    def lattice_packed(self, sents):
        """(edges, end_off, offsets, status): the edges of all sentences sorted by (sentence, end,
        begin, reference order) and the CSR index over (sentence offset + end - 1)."""
        unsupported = []
        text, offsets = pack_sentences(sents, unsupported)
        n = len(sents)
        n_units = int(offsets[-1])
        _native.check(self._lib.lt_lattice_host(self.batch, _native.ptr(text), _native.ptr(offsets), n))
        n_edges = ctypes.c_int64()
        _native.check(self._lib.lt_lattice_size(self.batch, ctypes.byref(n_edges)))
        edges = np.empty(max(1, n_edges.value), dtype=_native.EDGE_DTYPE)
        end_off = np.zeros(n_units + 1, dtype=np.int64)
        _native.check(self._lib.lt_lattice_fetch(self.batch, _native.ptr(edges), edges.size, _native.ptr(end_off)))
        status = np.zeros(max(1, n), dtype=np.int32)
        _native.check(self._lib.lt_lattice_status(self.batch, _native.ptr(status), None))
        for i in unsupported:
            status[i] = _native.LT_SENT_UNSUPPORTED_CHAR
        return edges[:n_edges.value], end_off, offsets, status[:n]

    def lattice_words(self, sents, errors='raise'):
        """Per sentence the dictionary edges as `Word`s, grouped by end position then begin position,
        each (begin, end) group in the reference's order."""
        sents = list(sents)
        edges, end_off, offsets, status = self.lattice_packed(sents)
        records = edges.tolist()
        ends = end_off.tolist()
        offs = offsets.tolist()
        out = []
        for i, sent in enumerate(sents):
            st = int(status[i])
            if st not in (_native.LT_SENT_OK, _native.LT_SENT_NO_EDGES):
                if errors == 'raise':
                    raise ValueError(self._status_message(st, i))
                out.append(None)
                continue
            chars = sent.replace(' ', '')
            out.append(self.records_to_words(chars, records[ends[offs[i]]:ends[offs[i + 1]]]))
        return out

    def import_lattice(self, sents, lattices):
        """Upload caller-built lattices (`lattices[i]` = iterable of `Word`s of sentence i, any order)
        for `lt_beam*`; returns the flat list of the imported `Word`s (index = record.rule)."""
        text, offsets = pack_sentences(sents)
        n = len(sents)
        n_units = int(offsets[-1])
        tag_ids = self.tables.tag_ids
        other = self.tables.other_tag_id
        flat, rec, strings = [], [], []
        end_off = np.zeros(n_units + 1, dtype=np.int64)
        for i, words in enumerate(lattices):
            m = len(sents[i].replace(' ', ''))
            base, top = int(offsets[i]), int(offsets[i + 1])
            usable = []
            for order, w in enumerate(words or ()):
                # only spans inside the sentence can be used by the search (beam.py:30-33); zero-width words
                # (flatten_words produces them) are never expanded
                if 0 <= w.b < w.e <= m:
                    usable.append((w.e, w.b, order, w))
            usable.sort(key=lambda t: t[:3])
            counts = np.zeros(top - base, dtype=np.int64)
            start = len(flat)
            for e, b, _, w in usable:
                idx = len(flat)
                flat.append(w)
                if not 0 <= w.len < 65536:
                    raise ValueError('Word.len %r is outside 0..65535' % (w.len,))
                tag0 = tag_ids.get(w.tag0, other)
                tag1 = 0xFF if w.tag1 is None else tag_ids.get(w.tag1, other)
                if tag0 is None or tag1 is None:
                    raise ValueError('tag %r / %r is unknown to the compiled tables' % (w.tag0, w.tag1))
                flags = (_native.LT_EDGE_EXPLICIT | (_native.LT_EDGE_IS_L if w.is_l else 0) |
                         (_native.LT_EDGE_LEMMA if w.tag1 is not None else 0))
                rec.append((b, e, int(w.len), tag0, tag1, idx, 0, flags, 0))
                strings += [w.word, w.morph0, w.morph1 or '']
                counts[e - 1] += 1
            if top > base:
                end_off[base + 1:top + 1] = start + np.cumsum(counts)
        edges = np.array(rec, dtype=_native.EDGE_DTYPE) if rec else np.zeros(1, dtype=_native.EDGE_DTYPE)
        from .compile import _encode_units
        units, str_off = _encode_units(strings)
        _native.check(self._lib.lt_lattice_import(self.batch, _native.ptr(text), _native.ptr(offsets), n, _native.ptr(edges),
                                                  _native.ptr(end_off), _native.ptr(units), _native.ptr(str_off), len(strings)))
        self._imported_words = flat
        return flat

    # -- unpacking -----------------------------------------------------------------------------------
    def _status_message(self, st, index):
        msg = _STATUS_MESSAGES.get(st, 'sentence %d could not be tagged')
        return msg % ((index, self.unit_limit) if st == _native.LT_SENT_TOO_LONG else index)

    def _raise_for(self, st, index):
        if st == _native.LT_SENT_NO_EDGES:
            raise IndexError('list index out of range')       # as the reference does (lookup.py:362-363 + beam.py:33)
        raise ValueError(self._status_message(st, index))

    def records_to_words(self, chars, records, imported=None):
        """Packed `lt_edge` tuples -> `Word` tuples (include/lt_b200.h documents the encoding)."""
        Word = _word_class()
        names = self.tables.tag_names
        rules = self.tables.rules_flat
        lemma_flag, is_l_flag, explicit = _native.LT_EDGE_LEMMA, _native.LT_EDGE_IS_L, _native.LT_EDGE_EXPLICIT
        new = tuple.__new__            # Word is a namedtuple: skips the per-call length check of _make
        words = []
        for b, e, length, tag0, tag1, rule, split, flags, _ in records:
            if flags & explicit:
                words.append(imported[rule])
                continue
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

    def _sequence(self, sent, records, score, imported=None):
        Word = _word_class()
        chars = sent.replace(' ', '')
        n = len(chars)
        words = [Word(BOS, BOS, None, BOS, None, 0, 0, 0, False)]
        words += self.records_to_words(chars, records, imported)
        words.append(Word(EOS, EOS, None, EOS, None, 0, n, n, False))
        # adding EOS resets the trailing-unknown count (beam.py:113 with tag0 == EOS)
        return _sequence_class()(words, score, 0)

    def sequence_at(self, sent, packed, index, errors):
        st = int(packed.status[index])
        if st != _native.LT_SENT_OK:
            if errors == 'raise':
                self._raise_for(st, index)
            return None
        lo, hi = int(packed.path_off[index]), int(packed.path_off[index + 1])
        return self._sequence(sent, packed.path_edges[lo:hi].tolist(), float(packed.scores[index]))

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
                    self._raise_for(st, i)
                out.append(None)
                continue
            out.append(self._sequence(sent, records[offs[i]:offs[i + 1]], score_list[i]))
        return out

    def unpack_kbest(self, sents, packed, beam_size, errors='raise', imported=None):
        """list (per sentence) of list of `Sequence` — every survivor, best first (beam.py:59-61)."""
        n_best, path_off, path_edges, scores, status = packed
        offs = path_off.tolist()
        records = path_edges.tolist()
        score_list = scores.tolist()
        out = []
        for i, sent in enumerate(sents):
            st = int(status[i])
            if st != _native.LT_SENT_OK:
                if errors == 'raise':
                    self._raise_for(st, i)
                out.append(None)
                continue
            base = i * beam_size
            out.append([self._sequence(sent, records[offs[base + r]:offs[base + r + 1]], score_list[base + r], imported)
                        for r in range(int(n_best[i]))])
        return out

    # -- instrumentation -------------------------------------------------------------------------------
    def counters(self):
        c = _native.lt_counters()
        _native.check(self._lib.lt_batch_counters(self.batch, ctypes.byref(c)))
        return c.as_dict()

    def timings(self):
        t = _native.lt_timings()
        _native.check(self._lib.lt_batch_timings(self.batch, ctypes.byref(t)))
        return t.as_dict()

    def set_stage_timing(self, on):
        """Per-stage CUDA events between the kernels of a batch on / off (off: programmatic dependent launches)."""
        _native.check(self._lib.lt_batch_set_stage_timing(self.batch, 1 if on else 0))

    def info(self):
        i = _native.lt_info()
        _native.check(self._lib.lt_batch_info(self.batch, ctypes.byref(i)))
        return i.as_dict()
