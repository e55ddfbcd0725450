"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden fixtures."""

import pytest

import lattice_based_tagger_b200 as pkg
from oracle import lattice_oracle as lo
from tests import _cases, _golden

pytestmark = pytest.mark.gpu


def _lattice_key(edges):
    """reference order -> the order the device emits: stable by (end, begin)"""
    return sorted(edges, key=lambda w: (w[7], w[6]))


@pytest.mark.parametrize('name', _golden.names())
def test_golden_fixtures(name):
    payload = _golden.load(name)
    case = payload['case']
    dictionary, funcs = _cases.build_objects(case, pkg)
    tagger = pkg.Tagger(dictionary, score_funcs=funcs, k3_first=payload['k3_first'])
    sents = case['sentences']
    # lattice
    for sent, entry, (words, bindex) in zip(sents, payload['expected'], tagger.lattice_batch(sents)):
        want = _lattice_key([_golden.edge(w) for w in entry['lattice']])
        assert [tuple(w) for w in words[1:-1]] == want, sent
    # best path for every recorded beam size
    beams = sorted({int(k) for entry in payload['expected'] for k in entry['beams']})
    for k in beams:
        got = tagger.tag_batch(sents, beam_size=k, errors='none')
        for sent, entry, seq in zip(sents, payload['expected'], got):
            want = _golden.expected_survivors(entry, k)
            if want is None:
                assert seq is None, sent
                continue
            words, score, num_unk = want[0]
            assert [tuple(w) for w in seq.sequences] == words, (sent, k)
            assert seq.score == score, (sent, k)
            assert seq.num_unk == num_unk


@pytest.mark.parametrize('seed', range(2000, 2030))
def test_random_cases_against_oracle(seed):
    case = _cases.random_case(seed, n_sent=24, features=True, prefs=(seed % 2 == 0), max_sent_len=60)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)
    dictionary, funcs = _cases.build_objects(case, pkg)
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    oracle = lo.OracleTagger(dictionary, funcs)
    sents = case['sentences']
    lat_counters = lo.Counters()
    for sent, (words, bindex) in zip(sents, tagger.lattice_batch(sents)):
        assert [tuple(w) for w in words[1:-1]] == _lattice_key(oracle.lattice(sent, lat_counters)), sent
    for k in (1, 2, 5, 10, 33, 64):
        got = tagger.tag_batch(sents, beam_size=k, errors='none')
        counters = lo.Counters()
        for sent, seq in zip(sents, got):
            try:
                want = oracle.tag(sent, k, counters)
            except IndexError:
                assert seq is None, sent
                continue
            assert [tuple(w) for w in seq.sequences] == want.words, (sent, k)
            assert seq.score == want.score, (sent, k)
        dev = tagger.counters()
        # the device counts the work of the reference's control flow
        for name in ('T', 'F', 'Bk', 'W'):
            assert dev[name] == getattr(counters, name), (name, k)
        assert dev['E'] == lat_counters.E and dev['P'] == lat_counters.P


def test_error_behaviour():
    dictionary = pkg.dictionary.DemoMorphemeDictionary()
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore())
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    with pytest.raises(IndexError):
        tagger.tag('가나다라')                      # no dictionary hit at all
    empty = tagger.tag('')
    assert [w.word for w in empty.sequences] == ['BOS', 'EOS'] and empty.score == 0
    assert tagger.tag_batch([]) == []
    assert [len(s.sequences) for s in tagger.tag_batch(['', ' ', ''])] == [2, 2, 2]
    assert tagger.lattice_batch([]) == []
    with pytest.raises(ValueError):
        tagger.tag('노래\t입니다')
    with pytest.raises(ValueError):
        pkg.Tagger(dictionary, score_funcs=pkg.beam.BeamScoreFunctions(object()))


def test_demo_known_answer():
    dictionary = pkg.dictionary.DemoMorphemeDictionary()
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore(unknown_penalty=-.1, known_preference=0.5))
    best = pkg.Tagger(dictionary, score_funcs=funcs).tag('너무너무너무는 아이오아이의 노래 입니다')
    assert best.score == 15.5
    assert [(w.word, w.tag0, w.len) for w in best.sequences[1:-1]] == [
        ('너무너무너무', 'Noun', 7), ('는', 'Josa', 7), ('아이오아이', 'Noun', 6), ('의', 'Josa', 6),
        ('노래', 'Noun', 2), ('입니다', 'Adjective', 3)]


def test_buffers_grow_and_rerun(monkeypatch):
    """Tiny initial staging / edge capacities force the overflow flags, the host enlarges the
    buffers and reruns on the device; results must not change."""
    case = _cases.random_case(4242, n_sent=40, features=True, max_sent_len=60)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=1), 1)
    dictionary, funcs = _cases.build_objects(case, pkg)
    sents = [s for s in case['sentences']]
    oracle = lo.OracleTagger(dictionary, funcs)
    monkeypatch.setenv('LT_HIT_CAP', '8')
    monkeypatch.setenv('LT_EDGE_CAP', '16')
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    for sent, (words, bindex) in zip(sents, tagger.lattice_batch(sents)):
        assert [tuple(w) for w in words[1:-1]] == _lattice_key(oracle.lattice(sent)), sent
    got = tagger.tag_batch(sents, beam_size=5, errors='none')
    for sent, seq in zip(sents, got):
        try:
            want = oracle.tag(sent, 5)
        except IndexError:
            assert seq is None
            continue
        assert [tuple(w) for w in seq.sequences] == want.words
        assert seq.score == want.score


def test_trained_weight_format_end_to_end():
    """Tagged corpus -> scan_features -> trainer weight format -> load_params -> GPU tagger, against
    the oracle fed with the same objects."""
    from lattice_based_tagger_b200.features import scan_features
    from lattice_based_tagger_b200.trainer import load_params
    pairs = [('너무너무너무 는  아이오아이 의  노래  입니다',
              '너무너무너무/Noun 는/Josa  아이오아이/Noun 의/Josa  노래/Noun  이/Adjective+ㅂ니다/Eomi'),
             ('아이오아이 는  공연 을  했다', '아이오아이/Noun 는/Josa  공연/Noun 을/Josa  하/Verb+았다/Eomi')]
    idx_to_feature, _, counts = scan_features(pairs, pkg.features.SimpleTrigramEncoder(),
                                              predefined_features={(6, n): 1 for n in range(1, 9)})
    params = {'idx_to_feature': idx_to_feature, 'coefficient': [0.1 * c for c in counts]}
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore(), load_params(params))
    dictionary = pkg.dictionary.DemoMorphemeDictionary()
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    oracle = lo.OracleTagger(dictionary, funcs)
    for sent in ('너무너무너무는 아이오아이의 노래 입니다', '아이오아이는 공연을 했다', '노래를 했다 우와'):
        got, want = tagger.tag(sent), oracle.tag(sent, 5)
        assert [tuple(w) for w in got.sequences] == want.words
        assert got.score == want.score


def test_edge_cases_against_oracle():
    """Inputs the reference treats specially: blank sentences, repeated / leading / trailing spaces,
    a dictionary entry longer than the beam window, a dictionary without stand-alone tags
    (MorphemeLookup.max_len = 0 -> the sub-word scan is unbounded, lookup.py:229-230), single
    syllables, and a long sentence."""
    tag_to_morphs = {
        'Josa': {'는', '을', '가'}, 'Noun': {'가나다라마바사아자차', '가', '나다', '라마'},
        'Eomi': {'다', 'ㄴ다', '았다'}, 'Verb': {'하', '가'}, 'Adjective': {'나'},
    }
    rules = {'한': (('하', 'ㄴ'),), '갔': (('가', '았'),), '했다': (('하', '았다'),)}
    dictionary = pkg.dictionary.MorphemeDictionary(tag_to_morphs, rules)
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore(-0.3, 0.4, -0.1))
    sents = ['', ' ', '   ', '가', ' 가 ', '가  나다', '가나다라마바사아자차', '가나다라마바사아자차는 갔다', '한다 했다 갔다',
             '나다라마가는' * 40, '가 ' * 100, 'x가나다라마y']
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    oracle = lo.OracleTagger(dictionary, funcs)
    for k in (1, 5, 40):
        got = tagger.tag_batch(sents, beam_size=k, errors='none')
        for sent, seq in zip(sents, got):
            try:
                want = oracle.tag(sent, k)
            except IndexError:
                assert seq is None, repr(sent)
                continue
            assert [tuple(w) for w in seq.sequences] == want.words, (repr(sent), k)
            assert seq.score == want.score
    # no stand-alone tag, no predicate tag: max_len = 0
    only = pkg.dictionary.MorphemeDictionary({'Josa': {'는', '가'}, 'Eomi': {'다'}, 'Pronoun': {'나', '너는'}}, {})
    tagger2 = pkg.Tagger(only, score_funcs=funcs)
    oracle2 = lo.OracleTagger(only, funcs)
    assert tagger2.eojeol_lookup.max_len == 0
    for sent in ['나는 너는', '가나는다', '너는나']:
        words, _ = tagger2.lattice_batch([sent])[0]
        assert [tuple(w) for w in words[1:-1]] == _lattice_key(oracle2.lattice(sent))
        try:
            want = oracle2.tag(sent)
        except IndexError:
            with pytest.raises(IndexError):
                tagger2.tag(sent)
            continue
        assert [tuple(w) for w in tagger2.tag(sent).sequences] == want.words


def test_rejected_inputs():
    dictionary = pkg.dictionary.DemoMorphemeDictionary()
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore())
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    with pytest.raises(ValueError):
        tagger.tag('노래 \U0001F600')                       # outside the BMP
    with pytest.raises(ValueError):
        tagger.tag_batch(['노래'], beam_size=65)
    with pytest.raises(TypeError):
        pkg.Tagger(dictionary, score_funcs=None).tag('노래')
    import numpy as np
    enc = pkg.features.SimpleTrigramEncoder({(4, 1): 0})
    with pytest.raises(ValueError):
        pkg.Tagger(dictionary, score_funcs=pkg.beam.BeamScoreFunctions(
            pkg.beam.SimpleTrigramFeatureScore(enc, np.zeros(1, dtype=np.float32))))

    class Custom(pkg.beam.BeamScoreFunction):
        pass
    with pytest.raises(ValueError):
        pkg.Tagger(dictionary, score_funcs=pkg.beam.BeamScoreFunctions(Custom()))
    with pytest.raises(ValueError):
        pkg.Tagger(object(), score_funcs=funcs)              # not a morpheme dictionary


def _check_against_oracle(tagger, oracle, sents, beams):
    for sent, (words, bindex) in zip(sents, tagger.lattice_batch(sents)):
        assert [tuple(w) for w in words[1:-1]] == _lattice_key(oracle.lattice(sent)), sent
    for k in beams:
        got = tagger.tag_batch(sents, beam_size=k, errors='none')
        for sent, seq in zip(sents, got):
            try:
                want = oracle.tag(sent, k)
            except IndexError:
                assert seq is None, sent
                continue
            assert [tuple(w) for w in seq.sequences] == want.words, (sent, k)
            assert seq.score == want.score, (sent, k)


def _dense_case(seed):
    """One long eojeol whose every prefix and suffix is a dictionary word under many tags: the
    bucket of the eojeol's last syllable holds far more than the 32 edges the beam kernel caches,
    and most of them span more than the 8-syllable window."""
    import random
    rng = random.Random(seed)
    alphabet = ['가', '나', '다']
    word = ''.join(rng.choice(alphabet) for _ in range(14))
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


@pytest.mark.parametrize('seed', [7, 8])
def test_buckets_beyond_the_edge_cache(seed):
    case = _dense_case(seed)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)
    dictionary, funcs = _cases.build_objects(case, pkg)
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    oracle = lo.OracleTagger(dictionary, funcs)
    sents = case['sentences']
    words, bindex = tagger.lattice_batch(sents[:1])[0]
    last = [w for w in words[1:-1] if w.e == len(sents[0])]
    assert len(last) > 40 and max(w.e - w.b for w in last) > 8       # the case does what it is built for
    _check_against_oracle(tagger, oracle, sents, (1, 5, 10, 20, 32, 40))


def test_trail_in_hbm_and_generic_array_sizes(monkeypatch):
    """Long sentences (sentence arrays beyond the templated 64 / 128 elements) and the HBM
    back-pointer trail forced for short ones."""
    case = _cases.random_case(5151, n_sent=16, features=True, prefs=True, max_sent_len=180)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=5), 5)
    dictionary, funcs = _cases.build_objects(case, pkg)
    oracle = lo.OracleTagger(dictionary, funcs)
    sents = case['sentences']
    assert max(len(s) for s in sents) > 128
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    _check_against_oracle(tagger, oracle, sents, (5, 10, 12))
    short = [s for s in sents if len(s) <= 40]
    _check_against_oracle(tagger, oracle, short, (5, 10))              # 64-element arrays, trail in shared memory
    monkeypatch.setenv('LT_TRAIL_SMEM', '0')
    _check_against_oracle(tagger, oracle, short, (5, 10, 12))          # same kernels, trail in HBM


def test_dictionary_mutation_and_refresh():
    """`add` / `remove_words` on the host dictionary (reference
    `dictionary.py:244-262`) + `Tagger.refresh()` recompiles the device tables; results follow the
    oracle on the mutated dictionary, including the stale `verbs/adjectives/eomis` views the
    reference keeps after `remove_words`."""
    dictionary = pkg.dictionary.DemoMorphemeDictionary()
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore(unknown_penalty=-.1, known_preference=0.5))
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    sents = ['트와이스의 노래 입니다', '아이오아이는 노래를 했다', '노래 입니다']
    first = tagger.tag_batch(sents, errors='none')
    assert [w.tag0 for w in first[0].sequences[1:3]] == ['Unknown', 'Josa'] or first[0].sequences[1].tag0 == 'Unknown'
    dictionary.add({'트와이스'}, 'Noun')
    dictionary.add('짱', 'Suffix', force=True)
    dictionary.remove_words({'입니다', '이'}, 'Adjective')
    with pytest.raises(ValueError):
        dictionary.add('x', 'NoSuchTag')
    tagger.refresh()
    _check_against_oracle(tagger, lo.OracleTagger(dictionary, funcs), sents + ['트와이스짱'], (1, 5))
    assert tagger.tag(sents[0]).sequences[1].tag0 == 'Noun'


def test_retry_pass_and_adaptive_staging(monkeypatch):
    """A staging area far too small for most eojeols: the sentences go through the retry pass, the
    main pass's capacity adapts over successive batches (>= 64 sentences), results never change."""
    case = _cases.random_case(777, n_sent=90, features=True, max_sent_len=50)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=2), 2)
    dictionary, funcs = _cases.build_objects(case, pkg)
    sents = list(case['sentences'])
    oracle = lo.OracleTagger(dictionary, funcs)
    monkeypatch.setenv('LT_HIT_CAP', '8')
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    for _ in range(4):                       # capacities are sticky per tagger: 8 -> 16 -> 32 ...
        _check_against_oracle(tagger, oracle, sents, (5,))
