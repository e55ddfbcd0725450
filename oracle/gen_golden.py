"""Writes tests/golden/*.json.gz from the REFERENCE ITSELF (test infrastructure).

Run in the build container, where the read-only reference checkout exists:

    PYTHONHASHSEED=0 PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Every fixture stores a plain-data case (tests/_cases.py), the hash-order facts observed in the
generating process (tag order Q1, rule order Q2, 2- vs 3-syllable conjugation order Q3) and what
`lattice_tagger` returned: the lattice (`sentence_lookup`), every beam survivor
(`beam_search`) for several beam sizes, or the IndexError the reference raises.  Scores are
stored as `float.hex()` so they replay bit for bit.
"""

import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.append('/root/reference')
sys.dont_write_bytecode = True

import numpy  # noqa: E402
numpy.int = int   # beam/score_funcs.py:143

import lattice_tagger as ref  # noqa: E402
from oracle import lattice_oracle as lo  # noqa: E402
from tests import _cases  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
BEAMS = (1, 3, 5, 32)


def k3_first_flags(rules):
    flags = {}
    for key in rules:
        if len(key) == 3 and key[:2] in rules:
            k2, k3 = key[:2], key
            flags[k3] = next(iter({k2, k3})) == k3
    return flags


def run_reference(dictionary, funcs, sentences, beams=BEAMS, all_survivors_for=(5,)):
    lookup = ref.dictionary.MorphemeLookup(dictionary, flatten=False)
    expected = []
    for sent in sentences:
        words, bindex = ref.dictionary.sentence_lookup_as_begin_index(sent, lookup)
        entry = {'lattice': [list(w) for w in words[1:-1]], 'beams': {}}
        chars = sent.replace(' ', '')
        for k in beams:
            try:
                matures = ref.beam.beam_search(bindex, chars, funcs, beam_size=k)
            except IndexError:
                entry['beams'][str(k)] = 'IndexError'
                continue
            keep = matures if k in all_survivors_for else matures[:1]
            entry['beams'][str(k)] = [
                {'words': [list(w) for w in m.sequences], 'score': float(m.score).hex(),
                 'num_unk': m.num_unk} for m in keep]
        expected.append(entry)
    return expected


def dump(name, payload):
    path = os.path.join(GOLDEN, name + '.json.gz')
    raw = json.dumps(payload, ensure_ascii=False, separators=(',', ':')).encode('utf-8')
    with gzip.GzipFile(path, 'wb', mtime=0) as f:
        f.write(raw)
    print('%-28s %8d bytes' % (name, os.path.getsize(path)))


def case_from_dictionary(dictionary, sentences, funcs_spec):
    tags = list(dictionary.tag_to_morphs.keys())
    return {
        'seed': None, 'tags': tags,
        'tag_to_morphs': {t: sorted(dictionary.tag_to_morphs[t]) for t in tags},
        'rules': {k: [list(c) for c in v] for k, v in dictionary.rules.items()},
        'sentences': sentences, 'funcs': funcs_spec, 'feature_keys': [], 'coefficients': [],
    }


def features_for(case, seed):
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)


def main():
    os.makedirs(GOLDEN, exist_ok=True)

    # 1. randomised adversarial cases
    for seed in range(1000, 1024):
        case = _cases.random_case(seed, features=True, prefs=(seed % 3 == 0))
        features_for(case, seed)
        dictionary, funcs = _cases.build_objects(case, ref)
        payload = {'case': case, 'k3_first': k3_first_flags(dictionary.rules),
                   'expected': run_reference(dictionary, funcs, case['sentences'])}
        payload['case']['coefficients'] = [float(c).hex() for c in case['coefficients']]
        dump('random_%d' % seed, payload)

    # 2. the demo dictionary (reference resources/demo_morph) with the README-style sentences
    demo = ref.dictionary.DemoMorphemeDictionary()
    sentences = ['너무너무너무는 아이오아이의 노래 입니다', '아이오아이의 노래를 했다', '우와! 노래를했다',
                 '공연을했다', '춤을 춥니다', '야호 아이오아이는 공연을 합니다', '이 노래는 너무너무너무 입니다',
                 '', '가나다라']
    spec = [{'kind': 'reg', 'unknown_penalty': -0.1, 'known_preference': 0.5, 'syllable_penalty': -0.2},
            {'kind': 'mpref', 'table': {'Noun': {'아이오아이': 2.2}}},
            {'kind': 'wpref', 'table': {'Adjective': {'입니다': 3.3}}},
            {'kind': 'trigram'}]
    case = case_from_dictionary(demo, sentences, spec)
    features_for(case, 7)
    dictionary, funcs = _cases.build_objects(case, ref)
    payload = {'case': case, 'k3_first': k3_first_flags(dictionary.rules),
               'expected': run_reference(dictionary, funcs, sentences, all_survivors_for=BEAMS)}
    payload['case']['coefficients'] = [float(c).hex() for c in case['coefficients']]
    dump('demo_morph', payload)

    # 3. BASELINE config 1: the bundled `base` dictionary (Noun.txt is absent from the checkout)
    base = ref.dictionary.BaseMorphemeDictionary()
    sentences = ['오늘은 날씨가 매우 좋았습니다', '나는 어제 친구와 함께 영화를 보았다', '그는 빨리 달렸지만 늦었다',
                 '아주 예쁜 꽃이 피었습니다', '우리는 모두 함께 갔다', '너무 추워서 집에 있었다',
                 '그것은 정말 아름다웠습니다', '이것은 무엇입니까', '빨리빨리 갑시다', '하늘이 파랬다',
                 '그녀는 천천히 걸어갔습니다 그리고 웃었다', '아이고 깜짝이야 정말 놀랐잖아요',
                 '열 명이 함께 왔고 스무 명은 돌아갔다']
    spec = [{'kind': 'reg', 'unknown_penalty': -0.1, 'known_preference': 0.2, 'syllable_penalty': -0.2},
            {'kind': 'trigram'}]
    case = case_from_dictionary(base, sentences, spec)
    features_for(case, 11)
    dictionary, funcs = _cases.build_objects(case, ref)
    assert list(dictionary.tag_to_morphs) == case['tags']
    payload = {'case': case, 'k3_first': k3_first_flags(dictionary.rules),
               'expected': run_reference(dictionary, funcs, sentences, beams=(1, 5, 16))}
    payload['case']['coefficients'] = [float(c).hex() for c in case['coefficients']]
    dump('base_c1', payload)


if __name__ == '__main__':
    main()
