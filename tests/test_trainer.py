"""Trainer: the reference's `train()` flow with a working perceptron epoch
whose decoding step is the batched GPU tagger."""

import pytest

import lattice_based_tagger_b200 as pkg
from lattice_based_tagger_b200.dictionary.text import text_to_words
from lattice_based_tagger_b200.trainer import load_params, train
from lattice_based_tagger_b200.trainer.train import _same_path, _surface

PAIRS = [
    ('너무너무너무 는  아이오아이 의  노래  입니다',
     '너무너무너무/Noun 는/Josa  아이오아이/Noun 의/Josa  노래/Noun  이/Adjective+ㅂ니다/Eomi'),
    ('아이 는  노래 를  했다', '아이/Noun 는/Josa  노래/Noun 를/Josa  하/Verb+았다/Eomi'),
    ('아이오아이 는  공연 을  했다', '아이오아이/Noun 는/Josa  공연/Noun 을/Josa  하/Verb+았다/Eomi'),
    ('우와  노래 이  있다', '우와/Exclamation  노래/Noun 이/Josa  있/Adjective+다/Eomi'),
    ('아이 의  노래 는  공연 입니다', '아이/Noun 의/Josa  노래/Noun 는/Josa  공연/Noun 이/Adjective+ㅂ니다/Eomi'),
]


def test_annotation_helpers():
    assert _surface(PAIRS[0][0]) == '너무너무너무는 아이오아이의 노래 입니다'
    gold = text_to_words(*PAIRS[1])
    assert _same_path(gold, list(gold))
    other = list(gold)
    other[1] = other[1]._replace(tag0='Verb')
    assert not _same_path(other, gold)
    assert not _same_path(gold[:-1], gold)


@pytest.mark.gpu
def test_perceptron_fits_the_toy_corpus():
    dictionary = pkg.dictionary.DemoMorphemeDictionary()
    # a regulariser that prefers unknown words: with zero weights every sentence is tagged wrongly
    reg = pkg.beam.RegularizationScore(unknown_penalty=-0.1, known_preference=-0.3, syllable_penalty=0.0)
    before = pkg.Tagger(dictionary, score_funcs=pkg.beam.BeamScoreFunctions(reg))
    wrong = sum(not _same_path(seq.sequences, text_to_words(w, m))
                for (w, m), seq in zip(PAIRS, before.tag_batch([_surface(w) for w, _ in PAIRS])))
    assert wrong == len(PAIRS)

    params = train(PAIRS, dictionary, pkg.features.SimpleTrigramEncoder(), pkg.beam.SimpleTrigramFeatureScore(), reg,
                   max_epochs=20)
    assert set(params) == {'idx_to_feature', 'coefficient'}                       # trainer weight format, train.py:34-37
    assert len(params['idx_to_feature']) == len(params['coefficient']) > 0

    tagger = pkg.Tagger(dictionary, score_funcs=pkg.beam.BeamScoreFunctions(reg, load_params(params)))
    for (word_text, morph_text), seq in zip(PAIRS, tagger.tag_batch([_surface(w) for w, _ in PAIRS])):
        assert _same_path(seq.sequences, text_to_words(word_text, morph_text)), word_text
