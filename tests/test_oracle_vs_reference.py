"""Pins the oracle (oracle/lattice_oracle.py) to the reference itself.

Runs only where the reference checkout exists (the build container); the committed golden
fixtures (tests/test_oracle_golden.py) carry the same evidence to machines without it.
"""

import pytest

from oracle import lattice_oracle as lo
from tests import _cases
from tests.conftest import import_reference


def _reference_survivors(ref, dictionary, funcs, sent, k):
    lookup = ref.dictionary.MorphemeLookup(dictionary, flatten=False)
    words, bindex = ref.dictionary.sentence_lookup_as_begin_index(sent, lookup)
    chars = sent.replace(' ', '')
    matures = ref.beam.beam_search(bindex, chars, funcs, beam_size=k)
    return words, matures


@pytest.mark.parametrize('seed', range(60))
def test_random_cases_match_reference(seed):
    ref = import_reference()
    case = _cases.random_case(seed, features=True, prefs=(seed % 3 == 0))
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)
    ref_dict, ref_funcs = _cases.build_objects(case, ref)
    view_tagger = lo.OracleTagger(ref_dict, ref_funcs)
    for sent in case['sentences']:
        for k in (1, 3, 5, 32):
            try:
                words, matures = _reference_survivors(ref, ref_dict, ref_funcs, sent, k)
            except IndexError:
                with pytest.raises(IndexError):
                    view_tagger.survivors(sent, k)
                continue
            assert [tuple(w) for w in words[1:-1]] == view_tagger.lattice(sent)
            mine = view_tagger.survivors(sent, k)
            assert len(mine) == len(matures)
            for got, want in zip(mine, matures):
                assert got.words == [tuple(w) for w in want.sequences]
                assert got.score == want.score
                assert got.num_unk == want.num_unk


def test_demo_sentence_matches_reference():
    ref = import_reference()
    dictionary = ref.dictionary.DemoMorphemeDictionary()
    funcs = ref.beam.BeamScoreFunctions(ref.beam.RegularizationScore(unknown_penalty=-.1, known_preference=0.5))
    sent = '너무너무너무는 아이오아이의 노래 입니다'
    want = ref.tagger.Tagger(dictionary, score_funcs=funcs).tag(sent)
    got = lo.OracleTagger(dictionary, funcs).tag(sent)
    assert got.words == [tuple(w) for w in want.sequences]
    assert got.score == want.score == 15.5


_LOOKUPS = {
    'lr': lambda ref, d, flatten: ref.dictionary.LRLookup(d, prefer_exact_match=True, flatten=flatten),
    'lr_all': lambda ref, d, flatten: ref.dictionary.LRLookup(d, prefer_exact_match=False, flatten=flatten),
    'word': lambda ref, d, flatten: ref.dictionary.WordLookup(d, prefer_exact_match=True, flatten=flatten),
    'word_all': lambda ref, d, flatten: ref.dictionary.WordLookup(d, prefer_exact_match=False, flatten=flatten),
    'morpheme': lambda ref, d, flatten: ref.dictionary.MorphemeLookup(d, flatten=flatten),
}


@pytest.mark.parametrize('seed', range(100, 124))
def test_alternative_lookups_match_reference(seed):
    """LRLookup / WordLookup (with and without prefer_exact_match), flatten=True and
    sentence_lookup_as_graph restated by the oracle, against the reference's own classes."""
    ref = import_reference()
    case = _cases.random_case(seed, features=True, prefs=False, max_sent_len=16)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)
    ref_dict, ref_funcs = _cases.build_objects(case, ref)
    for name, make in _LOOKUPS.items():
        tagger = lo.OracleTagger(ref_dict, ref_funcs, lookup=name)
        lookup = make(ref, ref_dict, False)
        flat = make(ref, ref_dict, True)
        for sent in case['sentences']:
            words, bindex = ref.dictionary.sentence_lookup_as_begin_index(sent, lookup)
            assert [tuple(w) for w in words[1:-1]] == tagger.lattice(sent), (name, sent)
            assert [tuple(w) for w in ref.dictionary.sentence_lookup(sent, flat)[1:-1]] == tagger.lattice(sent, flatten=True)
            chars = sent.replace(' ', '')
            for k in (1, 4):
                try:
                    matures = ref.beam.beam_search(bindex, chars, ref_funcs, beam_size=k)
                except IndexError:
                    with pytest.raises(IndexError):
                        tagger.survivors(sent, k)
                    continue
                mine = tagger.survivors(sent, k)
                assert len(mine) == len(matures)
                for got, want in zip(mine, matures):
                    assert got.words == [tuple(w) for w in want.sequences]
                    assert got.score == want.score
            try:
                nodes, links = ref.dictionary.sentence_lookup_as_graph(sent, lookup)
            except IndexError:
                with pytest.raises(IndexError):
                    lo.lattice_graph(sent, tagger.lattice(sent))
                continue
            got_nodes, got_links = lo.lattice_graph(sent, tagger.lattice(sent))
            assert got_nodes == [tuple(w) for w in nodes]
            assert got_links == [[tuple(a), tuple(b), wgt] for a, b, wgt in links]
