"""The oracle replayed against outputs of the reference itself (tests/golden, see
oracle/gen_golden.py).  Runs anywhere: needs neither the reference checkout nor a GPU."""

import numpy as np
import pytest

import lattice_based_tagger_b200 as pkg
from oracle import lattice_oracle as lo
from tests import _cases, _golden


@pytest.mark.parametrize('name', _golden.names())
def test_oracle_reproduces_golden(name):
    payload = _golden.load(name)
    case = payload['case']
    dictionary, funcs = _cases.build_objects(case, pkg)
    tagger = lo.OracleTagger(dictionary, funcs, k3_first=payload['k3_first'])
    for sent, entry in zip(case['sentences'], payload['expected']):
        assert tagger.lattice(sent) == [_golden.edge(w) for w in entry['lattice']]
        for k_str in entry['beams']:
            k = int(k_str)
            want = _golden.expected_survivors(entry, k)
            if want is None:
                with pytest.raises(IndexError):
                    tagger.survivors(sent, k)
                continue
            got = tagger.survivors(sent, k)
            if len(want) == 1 and k > 1:
                got = got[:1]
            assert len(got) == len(want)
            for hyp, (words, score, num_unk) in zip(got, want):
                assert hyp.words == words
                assert hyp.score == score
                assert hyp.num_unk == num_unk


def test_numpy_association_is_what_the_oracle_spells_out():
    # the reference sums <= 9 gathered fp64 weights with ndarray.sum()
    rng = np.random.default_rng(5)
    coef = rng.standard_normal(4096) * np.exp(rng.uniform(-20, 20, 4096))
    for n in range(1, 12):
        for _ in range(400):
            idx = rng.integers(0, coef.size, n)
            want = coef[np.asarray(list(idx), dtype=int)].sum()
            assert lo.numpy_order_sum([float(c) for c in coef[idx]]) == want


def test_demo_known_answer():
    # SURVEY App. C smoke value: RegularizationScore(-.1, .5) only -> 6 words, score 15.5
    dictionary = pkg.dictionary.DemoMorphemeDictionary()
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore(unknown_penalty=-.1, known_preference=0.5))
    best = lo.OracleTagger(dictionary, funcs).tag('너무너무너무는 아이오아이의 노래 입니다')
    assert best.score == 15.5
    assert [(w[0], w[3], w[5]) for w in best.words[1:-1]] == [
        ('너무너무너무', 'Noun', 7), ('는', 'Josa', 7), ('아이오아이', 'Noun', 6), ('의', 'Josa', 6),
        ('노래', 'Noun', 2), ('입니다', 'Adjective', 3)]
    assert best.words[6][1:5] == ('이', 'ㅂ니다', 'Adjective', 'Eomi')
