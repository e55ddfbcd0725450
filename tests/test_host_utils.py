"""Host utilities of the 'next' rows: tagged-text helpers, feature scanner, weight
format — compared with the reference where it is present, and against its docstring examples."""

import numpy as np
import pytest

import lattice_based_tagger_b200 as pkg
from lattice_based_tagger_b200.dictionary import flatten_words, str_to_morphtag, text_to_words
from lattice_based_tagger_b200.features import SimpleTrigramEncoder, scan_dictionary, scan_features
from lattice_based_tagger_b200.trainer import dump_params, load_params
from tests.conftest import import_reference

PAIRS = [
    ('너무너무너무 는  아이오아이 의  노래  입니다',
     '너무너무너무/Noun 는/Josa  아이오아이/Noun 의/Josa  노래/Noun  이/Adjective+ㅂ니다/Eomi'),
    ('빙수  고명 으로  얹는  삶은  단팥 과  찰떡  젤리  포장 도  나와 있다',
     '빙수/Noun  고명/Noun 으로/Josa  얹/Verb+는/Eomi  삶/Verb+은/Eomi  단팥/Noun 과/Josa  찰떡/Noun  '
     '젤리/Noun  포장/Noun 도/Josa  나오/Verb+아/Eomi 있/Verb+다/Eomi'),
    ('봤어  영화관 가면  늘  보는  정도 인데  뭘',
     '보/Verb+았어/Eomi  영화관/Noun 가/Verb+면/Eomi  늘/Adverb  보/Verb+는/Eomi  정도/Noun 인데/Josa  무엇/Pronoun+을/Josa'),
    ('broken  pair', 'only/Noun'),
]


def test_docstring_examples():
    # dictionary/dictionary.py:30-41 and :116-128 of the reference
    assert str_to_morphtag('이/Adjective+ㅂ니다/Eomi') == [['이', 'Adjective'], ['ㅂ니다', 'Eomi']]
    words = text_to_words(*PAIRS[0])
    assert [(w.word, w.tag0, w.b, w.e, w.is_l) for w in words[1:-1]] == [
        ('너무너무너무', 'Noun', 0, 6, True), ('는', 'Josa', 6, 7, False), ('아이오아이', 'Noun', 7, 12, True),
        ('의', 'Josa', 12, 13, False), ('노래', 'Noun', 13, 15, True), ('입니다', 'Adjective', 15, 18, True)]
    flat = flatten_words(words)
    assert [(w.word, w.tag0, w.len, w.b, w.e) for w in flat[6:8]] == [('이', 'Adjective', 1, 15, 16), ('ㅂ니다', 'Eomi', 2, 16, 18)]
    assert str(words[6]) == 'Word(입니다, 이/Adjective + ㅂ니다/Eomi, len=3, b=15, e=18, L)'


def test_matches_reference():
    ref = import_reference()
    for word_text, morph_text in PAIRS[:3]:
        want = ref.dictionary.text_to_words(word_text, morph_text)
        assert [tuple(w) for w in text_to_words(word_text, morph_text)] == [tuple(w) for w in want]
        assert [tuple(w) for w in flatten_words(text_to_words(word_text, morph_text))] == \
               [tuple(w) for w in ref.dictionary.flatten_words(want)]
    for flatten in (False, True):
        mine = scan_features(PAIRS, SimpleTrigramEncoder(), predefined_features={(6, n): 1 for n in range(1, 9)}, flatten=flatten)
        theirs = ref.features.scan_features(PAIRS, ref.features.SimpleTrigramEncoder(),
                                            predefined_features={(6, n): 1 for n in range(1, 9)}, flatten=flatten)
        assert mine[0] == theirs[0] and mine[1] == theirs[1] and mine[2] == theirs[2]
    assert scan_dictionary(PAIRS) == ref.features.scan_dictionary(PAIRS)


def test_weight_format_round_trip():
    idx_to_feature, feature_to_idx, _ = scan_features(PAIRS, SimpleTrigramEncoder())
    coef = np.linspace(-1, 1, len(idx_to_feature))
    score = load_params({'idx_to_feature': idx_to_feature, 'coefficient': list(coef)})
    assert score.encoder.feature_dic == feature_to_idx
    assert isinstance(score, pkg.beam.SimpleTrigramFeatureScore) and score.num_features == len(coef)
    back = dump_params(score)
    assert back['idx_to_feature'] == idx_to_feature and np.allclose(back['coefficient'], coef)
    with pytest.raises(ValueError):
        load_params({'idx_to_feature': idx_to_feature, 'coefficient': [0.0]})


def test_feature_padding_with_vocabulary():
    """C5's 10 M-weight target exceeds what lattice chains of the sample yield: the rest comes from
    word n-grams over the dictionary (seeded, so every rank builds the same table)."""
    import numpy as np
    from lattice_based_tagger_b200 import synth
    vocab = [('가%d' % i, 'Noun' if i % 2 else 'Verb') for i in range(500)]
    fd1, c1 = synth.make_features([], lambda s: [], lambda s: [], 5000, ['Noun', 'Verb'], seed=3, vocab=vocab)
    fd2, c2 = synth.make_features([], lambda s: [], lambda s: [], 5000, ['Noun', 'Verb'], seed=3, vocab=vocab)
    assert len(fd1) == 5000 and list(fd1) == list(fd2) and np.array_equal(c1, c2)
    assert sorted(fd1.values()) == list(range(5000))
    assert {f[0] for f in fd1} >= {0, 2, 3, 4, 6, 7, 8}
    # without a vocabulary the dictionary stops where the lattice chains run dry
    fd3, _ = synth.make_features([], lambda s: [], lambda s: [], 5000, ['Noun', 'Verb'], seed=3)
    assert len(fd3) < 5000
