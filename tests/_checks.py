"""Parity checks shared by the GPU tests (`-m gpu`, the real library) and the CPU tests that run the
same kernel sources through the SIMT emulator (tests/_emu.py).  Each check builds a tagger from this
package, an oracle from the same objects, and compares bit for bit."""

import pytest

import lattice_based_tagger_b200 as pkg
from oracle import lattice_oracle as lo
from tests import _cases

LOOKUPS = {
    'morpheme': lambda d, flatten=False: pkg.dictionary.MorphemeLookup(d, flatten=flatten),
    'lr': lambda d, flatten=False: pkg.dictionary.LRLookup(d, prefer_exact_match=True, flatten=flatten),
    'lr_all': lambda d, flatten=False: pkg.dictionary.LRLookup(d, prefer_exact_match=False, flatten=flatten),
    'word': lambda d, flatten=False: pkg.dictionary.WordLookup(d, prefer_exact_match=True, flatten=flatten),
    'word_all': lambda d, flatten=False: pkg.dictionary.WordLookup(d, prefer_exact_match=False, flatten=flatten),
}


def lattice_key(edges):
    """reference order -> the order the device emits: stable by (end, begin)"""
    return sorted(edges, key=lambda w: (w[7], w[6]))


def make_case(seed, n_sent=24, prefs=None, max_sent_len=60):
    case = _cases.random_case(seed, n_sent=n_sent, features=True, prefs=(seed % 2 == 0) if prefs is None else prefs,
                              max_sent_len=max_sent_len)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)
    return case


def check_against_oracle(tagger, oracle, sents, beams, counters=False):
    lat = lo.Counters()
    for sent, (words, bindex) in zip(sents, tagger.lattice_batch(sents)):
        assert [tuple(w) for w in words[1:-1]] == lattice_key(oracle.lattice(sent, lat)), sent
    for k in beams:
        got = tagger.tag_batch(sents, beam_size=k, errors='none')
        work = lo.Counters()
        for sent, seq in zip(sents, got):
            try:
                want = oracle.tag(sent, k, work)
            except IndexError:
                assert seq is None, sent
                continue
            assert [tuple(w) for w in seq.sequences] == want.words, (sent, k)
            assert seq.score == want.score, (sent, k)
        if counters:
            dev = tagger.counters()
            # the device counts the work of the reference's control flow
            for name in ('T', 'F', 'Bk', 'W'):
                assert dev[name] == getattr(work, name), (name, k)
            assert dev['E'] == lat.E and dev['P'] == lat.P


def check_kbest(tagger, oracle, sents, beams):
    """Every survivor of the beam: words, exact scores, order (beam.py:59-61)."""
    for k in beams:
        got = tagger.tag_batch_kbest(sents, beam_size=k, errors='none')
        best = tagger.tag_batch(sents, beam_size=k, errors='none')
        for sent, seqs, top in zip(sents, got, best):
            try:
                want = oracle.survivors(sent, k)
            except IndexError:
                assert seqs is None and top is None, sent
                continue
            assert len(seqs) == len(want), (sent, k)
            for seq, hyp in zip(seqs, want):
                assert [tuple(w) for w in seq.sequences] == hyp.words, (sent, k)
                assert seq.score == hyp.score
            assert top == seqs[0]


def check_lookup_modes(case, beams=(1, 4)):
    """LRLookup / WordLookup / flatten lattices and the searches over them."""
    dictionary, funcs = _cases.build_objects(case, pkg)
    sents = case['sentences']
    for name, make in LOOKUPS.items():
        for flatten in (False, True):
            oracle = lo.OracleTagger(dictionary, funcs, lookup=name)
            lookup = make(dictionary, flatten)
            tagger = pkg.Tagger(dictionary, lookup=lookup, score_funcs=funcs)
            for sent, words in zip(sents, lookup.lookup_batch(sents)):
                # flatten=False: device order (end, begin); flatten=True: the reference's own order
                want = oracle.lattice(sent, flatten=True) if flatten else lattice_key(oracle.lattice(sent))
                assert [tuple(w) for w in words] == want, (name, flatten, sent)
            for sent in sents[:6]:
                # sentence_lookup: the reference's enumeration order, word for word
                got = pkg.dictionary.sentence_lookup(sent, lookup)[1:-1]
                assert [tuple(w) for w in got] == oracle.lattice(sent, flatten=flatten), (name, flatten, sent)
            for k in beams:
                got = tagger.tag_batch_kbest(sents, beam_size=k, errors='none')
                for sent, seqs in zip(sents, got):
                    edges = oracle.lattice(sent, flatten=flatten)
                    chars = sent.replace(' ', '')
                    try:
                        want = lo.beam_search(lo.begin_index(sent, edges), chars, oracle.program, k)
                    except IndexError:
                        # (flatten: the reference also fails on a zero-width word at the sentence end, an artefact of
                        # its bindex construction; the device drops words that cannot be expanded instead)
                        if not edges:
                            assert seqs is None, (name, flatten, sent)
                        continue
                    assert len(seqs) == len(want), (name, flatten, sent, k)
                    for seq, hyp in zip(seqs, want):
                        assert [tuple(w) for w in seq.sequences] == hyp.words, (name, flatten, sent, k)
                        assert seq.score == hyp.score
            tagger.close()


def check_host_api(case):
    """The reference's function-level API answered by the device: MorphemeDictionary.lookup / lemmatize,
    analyze_morphology, sentence_lookup*, beam_search on a caller-built bindex."""
    dictionary, funcs = _cases.build_objects(case, pkg)
    view = lo.DictView(dictionary)
    program = lo.ScoreProgram(funcs)
    sents = case['sentences']
    pieces = sorted({e for s in sents for e in s.split()})[:24]
    for piece in pieces:
        want = lo.full_lookup(piece, view, 3, False)
        assert [tuple(w) for w in dictionary.lookup(piece, b=3)] == want, piece
        assert dictionary.lemmatize(piece) == lo.lemmatize(piece, view)
        assert pkg.dictionary.analyze_morphology(piece, dictionary.verbs, dictionary.adjectives, dictionary.eomis,
                                                 dictionary.rules) == lo.lemmatize(piece, view)
    lookup = pkg.dictionary.LRLookup(dictionary, prefer_exact_match=False)
    for sent in sents[:8]:
        edges = lo.sentence_edges(sent, view, lookup='lr_all')
        words, bindex = pkg.dictionary.sentence_lookup_as_begin_index(sent, lookup)
        assert [tuple(w) for w in words[1:-1]] == edges
        chars = sent.replace(' ', '')
        try:
            want = lo.beam_search(lo.begin_index(sent, edges), chars, program, 3)
        except IndexError:
            with pytest.raises(IndexError):
                pkg.beam.beam_search(bindex, chars, funcs, beam_size=3)
            with pytest.raises(IndexError):
                pkg.dictionary.sentence_lookup_as_graph(sent, lookup)
            continue
        got = pkg.beam.beam_search(bindex, chars, funcs, beam_size=3)
        assert len(got) == len(want)
        for seq, hyp in zip(got, want):
            assert [tuple(w) for w in seq.sequences] == hyp.words, sent
            assert seq.score == hyp.score
        nodes, links = pkg.dictionary.sentence_lookup_as_graph(sent, lookup)
        want_nodes, want_links = lo.lattice_graph(sent, edges)
        assert [tuple(w) for w in nodes] == want_nodes
        assert [[tuple(a), tuple(b), w] for a, b, w in links] == want_links


def dense_case(seed, syllables=14):
    """One long eojeol whose every prefix and suffix is a dictionary word under many tags: the
    bucket of the eojeol's last syllable holds more edges than the beam kernel caches per end position,
    and most of them span more than the 8-syllable window."""
    import random
    rng = random.Random(seed)
    alphabet = ['가', '나', '다']
    word = ''.join(rng.choice(alphabet) for _ in range(syllables))
    tags = ['Noun', 'Adverb', 'Exclamation', 'Determiner', 'Number', 'Pronoun', 'Josa', 'Eomi', 'Verb', 'Adjective']
    tag_to_morphs = {t: set() for t in tags}
    for i in range(1, len(word)):
        for piece in (word[:i], word[i:]):
            for t in rng.sample(tags[:6], 5):
                tag_to_morphs[t].add(piece)
    for t in ('Josa', 'Eomi', 'Verb', 'Adjective'):
        tag_to_morphs[t].update({word[-1], word[-2:], word[:2]})
    case = {'seed': seed, 'tags': tags, 'tag_to_morphs': {t: sorted(m) for t, m in tag_to_morphs.items()},
            'rules': {word[3]: [(word[3], word[-1])], word[5:7]: [(word[5], word[-2:])]},
            'sentences': [word, word + ' ' + word[:5], word[2:] + word, word[:9] + ' ' + word[4:]],
            'funcs': [{'kind': 'reg', 'unknown_penalty': -0.5, 'known_preference': 0.5, 'syllable_penalty': -0.2},
                      {'kind': 'trigram'}],
            'feature_keys': [], 'coefficients': []}
    return case


def check_dense_case(seed, syllables, beams, min_bucket):
    """Lattice order, paths and scores of the dense case against the oracle; the case must hold a bucket of at
    least `min_bucket` edges with spans beyond the window."""
    import lattice_based_tagger_b200 as pkg
    case = dense_case(seed, syllables)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)
    dictionary, funcs = _cases.build_objects(case, pkg)
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    oracle = lo.OracleTagger(dictionary, funcs)
    sents = case['sentences']
    words, bindex = tagger.lattice_batch(sents[:1])[0]
    last = [w for w in words[1:-1] if w.e == len(sents[0])]
    assert len(last) >= min_bucket and max(w.e - w.b for w in last) > 8       # the case does what it is built for
    check_against_oracle(tagger, oracle, sents, beams)
