"""Seeded adversarial cases for differential tests (small alphabets, dense lattices, ties).

`random_case(seed)` returns a plain-data description — dictionary with an explicit tag order,
conjugation rules with 1/2/3-syllable keys, sentences with random spacing and out-of-dictionary
syllables, a feature dictionary + coefficients and a list of scorer specs — from which
`build_objects(case, module)` instantiates either the reference's classes (`lattice_tagger`) or
this repository's descriptors (`lattice_based_tagger_b200`).
"""

import random

TAGS = ['Noun', 'Pronoun', 'Number', 'Josa', 'Adjective', 'Verb', 'Eomi', 'Adverb', 'Determiner',
        'Exclamation']
JAMO = 'ㄱㄴㄹㅁㅂㅆ'


def _rand_string(rng, alphabet, lo, hi):
    return ''.join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


def random_case(seed, n_sent=12, features=True, prefs=False, max_sent_len=30):
    rng = random.Random(seed)
    n_alpha = rng.choice([3, 4, 6, 10])
    alphabet = [chr(0xAC00 + 28 * rng.randrange(0, 390)) for _ in range(n_alpha)]
    alphabet = sorted(set(alphabet))
    outsiders = [chr(0xB000 + rng.randrange(0, 2000)) for _ in range(3)] + ['B', 'O', 'S', '!']

    tags = list(TAGS)
    rng.shuffle(tags)
    tags = tags[:rng.randint(5, 10)]
    for must in ('Noun', 'Josa', 'Eomi', 'Verb', 'Adjective'):
        if must not in tags and rng.random() < 0.9:
            tags.insert(rng.randrange(len(tags) + 1), must)
    if rng.random() < 0.15:
        tags.insert(rng.randrange(len(tags) + 1), 'Suffix')          # a tag outside tagset.py
    tag_to_morphs = {}
    for tag in tags:
        count = rng.randint(1, 14)
        if tag == 'Eomi':
            morphs = set()
            for _ in range(count):
                m = _rand_string(rng, alphabet, 1, 3)
                if rng.random() < 0.3:
                    m = rng.choice(JAMO) + m[1:]
                morphs.add(m)
        else:
            hi = 10 if (tag == 'Noun' and rng.random() < 0.1) else 4
            morphs = {_rand_string(rng, alphabet, 1, hi) for _ in range(count)}
        tag_to_morphs[tag] = sorted(morphs)

    rules = {}
    for _ in range(rng.randint(0, 12)):
        key = _rand_string(rng, alphabet, 1, 3)
        canons = []
        for _ in range(rng.randint(1, 3)):
            stem = _rand_string(rng, alphabet, 1, 2)
            eomi = _rand_string(rng, alphabet, 1, 2)
            if rng.random() < 0.3:
                eomi = rng.choice(JAMO) + eomi[1:]
            if (stem, eomi) not in canons:
                canons.append((stem, eomi))
        rules[key] = canons
    # make some rules productive: stem in Verb/Adjective and eomi + nothing in Eomi
    for key, canons in list(rules.items()):
        for stem, eomi in canons:
            if rng.random() < 0.5:
                for tag in ('Verb', 'Adjective'):
                    if tag in tag_to_morphs and rng.random() < 0.6:
                        tag_to_morphs[tag] = sorted(set(tag_to_morphs[tag]) | {stem})
                if 'Eomi' in tag_to_morphs:
                    tail = _rand_string(rng, alphabet, 0, 1)
                    tag_to_morphs['Eomi'] = sorted(set(tag_to_morphs['Eomi']) | {eomi + tail})

    sentences = []
    for _ in range(n_sent):
        length = rng.randint(1, max_sent_len)
        pieces = []
        total = 0
        while total < length:
            r = rng.random()
            if r < 0.6:
                tag = rng.choice(tags)
                piece = rng.choice(tag_to_morphs[tag])
            elif r < 0.9:
                piece = _rand_string(rng, alphabet, 1, 3)
            else:
                piece = rng.choice(outsiders)
            if rules and rng.random() < 0.25:
                piece += rng.choice(list(rules.keys()))
            pieces.append(piece)
            total += len(piece)
            if rng.random() < 0.4:
                pieces.append(' ' * rng.choice([1, 1, 1, 2]))
        sent = ''.join(pieces)
        if rng.random() < 0.1:
            sent = ' ' + sent
        if rng.random() < 0.1:
            sent = sent + ' '
        sentences.append(sent)
    sentences.append('')
    sentences.append(rng.choice(outsiders) * 3)          # no dictionary hit at all -> IndexError

    case = {'seed': seed, 'tags': tags, 'tag_to_morphs': tag_to_morphs, 'rules': rules,
            'sentences': sentences, 'funcs': [], 'feature_keys': [], 'coefficients': []}

    reg = {'kind': 'reg', 'unknown_penalty': rng.choice([-0.1, -0.5, -1.0]),
           'known_preference': rng.choice([0.2, 0.5, 1]), 'syllable_penalty': rng.choice([-0.2, -0.7, 0.0])}
    funcs = [reg]
    if prefs:
        mp = {}
        wp = {}
        for tag in rng.sample(tags, min(3, len(tags))):
            mp[tag] = {m: round(rng.uniform(-1, 2), 1) for m in rng.sample(tag_to_morphs[tag], min(3, len(tag_to_morphs[tag])))}
            wp[tag] = {m: round(rng.uniform(-1, 2), 1) for m in rng.sample(tag_to_morphs[tag], min(2, len(tag_to_morphs[tag])))}
        if 'Eomi' in tag_to_morphs:
            mp.setdefault('Eomi', {})[rng.choice(tag_to_morphs['Eomi'])] = 2
        funcs.append({'kind': 'mpref', 'table': mp})
        funcs.append({'kind': 'wpref', 'table': wp})
    if features:
        funcs.append({'kind': 'trigram'})
    rng.shuffle(funcs)
    case['funcs'] = funcs
    return case


def add_features(case, observed, seed, keep=0.6, extra=20):
    """Fill `feature_keys` / `coefficients` from feature tuples observed on lattices."""
    rng = random.Random(seed ^ 0x5EED)
    observed = sorted(set(observed), key=repr)
    keys = [f for f in observed if rng.random() < keep]
    tags = case['tags'] + ['BOS', 'Unknown']
    for _ in range(extra):
        keys.append((3, rng.choice(tags), rng.choice(tags)))
        keys.append((4, rng.randint(1, 9)))
        keys.append((6, rng.randint(1, 8)))
    keys = sorted(set(keys), key=repr)
    rng.shuffle(keys)
    case['feature_keys'] = keys
    # a coarse grid makes exact ties between different paths common
    if rng.random() < 0.5:
        case['coefficients'] = [rng.choice([-1.0, -0.5, 0.0, 0.25, 0.5, 1.0]) for _ in keys]
    else:
        case['coefficients'] = [rng.gauss(0.0, 1.0) for _ in keys]
    return case


def build_objects(case, pkg):
    """(dictionary, score_funcs) built from `pkg` = `lattice_tagger` or `lattice_based_tagger_b200`."""
    import numpy as np
    tag_to_morphs = {tag: set(case['tag_to_morphs'][tag]) for tag in case['tags']}
    rules = {key: tuple(tuple(c) for c in canons) for key, canons in case['rules'].items()}
    dictionary = pkg.dictionary.MorphemeDictionary(tag_to_morphs, rules)
    funcs = []
    for spec in case['funcs']:
        if spec['kind'] == 'reg':
            funcs.append(pkg.beam.RegularizationScore(spec['unknown_penalty'], spec['known_preference'],
                                                      spec['syllable_penalty']))
        elif spec['kind'] == 'mpref':
            funcs.append(pkg.beam.MorphemePreferenceScore(spec['table']))
        elif spec['kind'] == 'wpref':
            funcs.append(pkg.beam.WordPreferenceScore(spec['table']))
        else:
            feature_dic = {tuple(k): i for i, k in enumerate(case['feature_keys'])}
            encoder = pkg.features.SimpleTrigramEncoder(feature_dic)
            coefficients = np.asarray(case['coefficients'], dtype=np.float64)
            funcs.append(pkg.beam.SimpleTrigramFeatureScore(encoder, coefficients))
    return dictionary, pkg.beam.BeamScoreFunctions(*funcs)


def observed_features(case, oracle_module, per_sentence=200, seed=0):
    """Feature tuples of random (i, j, k) chains through each sentence's lattice (plus unknowns)."""
    rng = random.Random(seed)

    class _D:
        pass
    d = _D()
    d.tag_to_morphs = {tag: set(case['tag_to_morphs'][tag]) for tag in case['tags']}
    d.rules = {key: tuple(tuple(c) for c in canons) for key, canons in case['rules'].items()}
    d.verbs = d.tag_to_morphs.get('Verb', {})
    d.adjectives = d.tag_to_morphs.get('Adjective', {})
    d.eomis = d.tag_to_morphs.get('Eomi', {})
    view = oracle_module.DictView(d)
    out = []
    bos = ('BOS', 'BOS', None, 'BOS', None, 0, 0, 0, False)
    for sent in case['sentences']:
        chars = sent.replace(' ', '')
        edges = oracle_module.sentence_edges(sent, view)
        by_b = {}
        for edge in edges:
            by_b.setdefault(edge[6], []).append(edge)

        def pick(b):
            cands = list(by_b.get(b, []))
            for e in range(b + 1, min(len(chars), b + 4) + 1):
                sub = chars[b:e]
                cands.append((sub, sub, None, 'Unknown', None, e - b, b, e, False))
            return rng.choice(cands) if cands else None

        for _ in range(per_sentence):
            b = rng.randrange(0, max(1, len(chars)))
            wi = None if rng.random() < 0.2 else bos
            wj = bos
            if b > 0 or rng.random() < 0.5:
                first = pick(b)
                if first is None:
                    continue
                wj = first
                nxt = pick(first[7]) if first[7] < len(chars) else None
                if nxt is None:
                    out += oracle_module.feature_tuples(None, bos, first)
                    continue
                wk = nxt
                if wi is bos and rng.random() < 0.5:
                    third = pick(wk[7]) if wk[7] < len(chars) else None
                    if third is not None:
                        out += oracle_module.feature_tuples(wj, wk, third)
                out += oracle_module.feature_tuples(wi, wj, wk)
            else:
                first = pick(0)
                if first is not None:
                    out += oracle_module.feature_tuples(None, bos, first)
    return out
