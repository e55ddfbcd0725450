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
