"""Seeded synthetic workloads for the benchmark configurations.

There is no network and the reference ships neither a corpus nor weights, so every input of the
benchmark is generated: a Hangul alphabet, a morpheme dictionary with `tagset.py` tags, conjugation
rules, sentences composed from the dictionary, and a trigram feature dictionary with random
weights.  All generators take explicit seeds (0 dictionary, 1 rules, 2 sentences, 3 features).

`make_features` needs lattices and best paths to draw realistic feature tuples from; it takes
them from a callable so that the benchmark feeds it with this package's GPU tagger and the CPU
tests with the oracle.
"""

import numpy as np

from .dictionary import MorphemeDictionary
from .features import trigram_encoder
from .dictionary import Word
from .tagset import (Adjective, Adverb, Determiner, Eomi, Exclamation, Josa, Noun, Number, Pronoun,
                     Verb, BOS, Unk)

TAG_MIX = ((Noun, 0.40), (Pronoun, 0.0225), (Number, 0.0225), (Josa, 0.10), (Adjective, 0.08), (Verb, 0.15),
           (Eomi, 0.12), (Adverb, 0.06), (Determiner, 0.0225), (Exclamation, 0.0225))
LENGTH_PMF = (0.05, 0.30, 0.30, 0.20, 0.10, 0.05)          # 1..6 syllables, mirrors `base`
JAMO = [chr(c) for c in range(0x3131, 0x314F)]

CONFIGS = {
    # name: sentences, mean syllables, dictionary entries, alphabet size, beam, features
    'c2': dict(n_sent=10_000, mean_len=20, n_dict=100_000, alphabet=1500, beam=5, n_feat=100_000, fixed_len=False),
    'c3': dict(n_sent=1_000_000, mean_len=40, n_dict=1_000_000, alphabet=1500, beam=10, n_feat=100_000, fixed_len=False),
    # (alphabet 225: 8.1 lattice edges per position over 16 k positions, BASELINE configs[3] asks for ~8; with 400
    # syllables — the workload of every C4 number up to profiles/r2z — it was 4.8)
    'c4': dict(n_sent=100_000, mean_len=256, n_dict=100_000, alphabet=225, beam=32, n_feat=100_000, fixed_len=True),
    'c5': dict(n_sent=1_000_000, mean_len=40, n_dict=1_000_000, alphabet=1500, beam=10, n_feat=10_000_000, fixed_len=False),
    'tiny': dict(n_sent=256, mean_len=20, n_dict=5_000, alphabet=300, beam=5, n_feat=5_000, fixed_len=False),
}


def make_alphabet(size, seed=0):
    rng = np.random.default_rng(seed)
    codes = rng.choice(np.arange(0xAC00, 0xD7A4), size=size, replace=False)
    return [chr(int(c)) for c in codes]


def make_dictionary(n_entries, alphabet, seed=0):
    """`{tag: set(morph)}` in tagset order with `n_entries` distinct (morph, tag) pairs."""
    rng = np.random.default_rng(seed)
    tags = [t for t, _ in TAG_MIX]
    probs = np.asarray([p for _, p in TAG_MIX])
    probs = probs / probs.sum()
    tag_to_morphs = {t: set() for t in tags}
    total = 0
    alpha = np.asarray(alphabet)
    while total < n_entries:
        want = int((n_entries - total) * 1.3) + 16
        tag_idx = rng.choice(len(tags), size=want, p=probs)
        lengths = rng.choice(np.arange(1, 7), size=want, p=LENGTH_PMF)
        sylls = rng.integers(0, len(alphabet), size=(want, 6))
        jamo = rng.random(want) < 0.1
        jamo_pick = rng.integers(0, len(JAMO), size=want)
        for i in range(want):
            tag = tags[tag_idx[i]]
            morph = ''.join(alpha[sylls[i, :lengths[i]]])
            if tag == Eomi and jamo[i]:
                morph = JAMO[jamo_pick[i]] + morph[1:]
            bucket = tag_to_morphs[tag]
            if morph not in bucket:
                bucket.add(morph)
                total += 1
                if total >= n_entries:
                    break
    return tag_to_morphs


def make_rules(tag_to_morphs, alphabet, n_keys=5000, seed=1):
    """`{surface: ((stem, eomi), ...)}`: key length pmf (1,2,3) = (10,87,3)%, one to three canonical
    forms per key.  Stems are final syllables of predicates and eomis are prefixes of Eomi entries,
    so that a part of the rules produces dictionary-checked analyses."""
    rng = np.random.default_rng(seed)
    predicates = sorted(tag_to_morphs[Verb] | tag_to_morphs[Adjective])
    eomis = sorted(tag_to_morphs[Eomi])
    rules = {}
    alpha = np.asarray(alphabet)
    while len(rules) < n_keys:
        klen = int(rng.choice([1, 2, 3], p=[0.10, 0.87, 0.03]))
        key = ''.join(alpha[rng.integers(0, len(alphabet), size=klen)])
        if key in rules:
            continue
        canons = []
        for _ in range(int(rng.integers(1, 4))):
            stem = predicates[int(rng.integers(0, len(predicates)))][-1]
            e = eomis[int(rng.integers(0, len(eomis)))]
            eomi = e[:int(rng.integers(1, 3))]
            if (stem, eomi) not in canons:
                canons.append((stem, eomi))
        rules[key] = tuple(canons)
    return rules


def conjugated_surfaces(tag_to_morphs, rules, count, seed):
    """Surfaces `prefix + key + rest` whose lemmatisation (prefix + stem, eomi + rest) is in the
    dictionary for at least some entries — they make the lemmatizer's edges appear in the lattice."""
    rng = np.random.default_rng(seed)
    by_last = {}
    for m in tag_to_morphs[Verb] | tag_to_morphs[Adjective]:
        by_last.setdefault(m[-1], []).append(m)
    for v in by_last.values():
        v.sort()
    by_prefix = {}
    for m in tag_to_morphs[Eomi]:
        for k in (1, 2):
            if len(m) >= k:
                by_prefix.setdefault(m[:k], []).append(m)
    for v in by_prefix.values():
        v.sort()
    keys = sorted(k for k in rules if len(k) <= 2)
    out = []
    guard = 0
    while len(out) < count and guard < count * 50 and keys:
        guard += 1
        key = keys[int(rng.integers(0, len(keys)))]
        stem, eomi = rules[key][int(rng.integers(0, len(rules[key])))]
        preds = by_last.get(stem)
        tails = by_prefix.get(eomi)
        if not preds or not tails:
            continue
        pred = preds[int(rng.integers(0, len(preds)))]
        tail = tails[int(rng.integers(0, len(tails)))]
        out.append(pred[:-1] + key + tail[len(eomi):])
    return out


def make_sentences(tag_to_morphs, rules, n_sent, mean_len, alphabet, seed=2, fixed_len=False,
                   p_space=0.5, p_noise=0.02, p_conj=0.1):
    """Sentences of dictionary morphemes: 70 % Zipf(1.1) over a fixed permutation, 30 % uniform;
    an eojeol closes after each morpheme with probability `p_space`; `p_noise` of the syllables are
    replaced by out-of-alphabet syllables (they force unknown words)."""
    rng = np.random.default_rng(seed)
    morphs = sorted({m for ms in tag_to_morphs.values() for m in ms})
    perm = rng.permutation(len(morphs))
    conj = conjugated_surfaces(tag_to_morphs, rules, max(64, len(morphs) // 50), seed + 101)
    in_alpha = set(alphabet)
    outsiders = [chr(c) for c in range(0xAC00, 0xD7A4) if chr(c) not in in_alpha][:997]
    sents = []
    for _ in range(n_sent):
        if fixed_len:
            target = int(mean_len)
        else:
            target = max(2, int(round(rng.normal(mean_len, mean_len / 4.0))))
        pieces = []
        total = 0
        # draw a block of random numbers per sentence (cheaper than scalar calls)
        n_draw = target + 4
        zipf = np.minimum(rng.zipf(1.1, size=n_draw), len(morphs)) - 1
        unif = rng.integers(0, len(morphs), size=n_draw)
        pick = rng.random(n_draw)
        space = rng.random(n_draw) < p_space
        k = 0
        while total < target:
            if conj and pick[k] < p_conj:
                m = conj[int(unif[k]) % len(conj)]
            elif pick[k] < p_conj + 0.7 * (1 - p_conj):
                m = morphs[perm[zipf[k]]]
            else:
                m = morphs[unif[k]]
            if total + len(m) > target:
                m = m[:target - total]
            pieces.append(m)
            total += len(m)
            if space[k] and total < target:
                pieces.append(' ')
            k += 1
        sent = ''.join(pieces)
        if p_noise > 0:
            chars = list(sent)
            hits = np.nonzero(rng.random(len(chars)) < p_noise)[0]
            for h in hits:
                if chars[h] != ' ':
                    chars[h] = outsiders[int(rng.integers(0, len(outsiders)))]
            sent = ''.join(chars)
        sents.append(sent)
    return sents


def make_features(sample_sents, tag_fn, lattice_fn, n_features, tags, seed=3, vocab=None):
    """`feature_dic` (tuple -> index, insertion order) and fp64 `coefficients`.

    `tag_fn(sents) -> [Sequence | None]` gives best paths (features that really fire);
    `lattice_fn(sents) -> [(words, bindex)]` gives lattices from which random (i, j, k) chains pad
    the dictionary up to `n_features`.  All (3, ti, tj), (4, 1..8) and (6, 1..8) are included.
    When the lattice chains run dry before the target (the 10 M-weight table of C5), word n-grams
    over `vocab` = [(word, tag)] fill the rest: they rarely fire, they make the table as large as
    the configuration says.
    """
    rng = np.random.default_rng(seed)
    feature_dic = {}

    def add(feature):
        if feature not in feature_dic:
            feature_dic[feature] = len(feature_dic)

    n_path = max(1, len(sample_sents) // 4)
    for seq in tag_fn(sample_sents[:n_path]):
        if seq is None:
            continue
        words = seq.sequences
        previous = [None] + list(words)
        for wi, wj, wk in zip(previous, words, words[1:-1]):
            for f in trigram_encoder(wi, wj, wk):
                add(f)
    all_tags = list(tags) + [BOS, Unk]
    for ti in all_tags:
        for tj in all_tags:
            add((3, ti, tj))
    for n in range(1, 9):
        add((4, n))
        add((6, n))
    bos = Word(BOS, BOS, None, BOS, None, 0, 0, 0, False)
    lattices = lattice_fn(sample_sents)
    guard = 0
    while len(feature_dic) < n_features and guard < 64:
        guard += 1
        before = len(feature_dic)
        for sent, (words, bindex) in zip(sample_sents, lattices):
            if not bindex:
                continue
            chars = sent.replace(' ', '')
            n = len(chars)

            def pick(b):
                cands = bindex[b] if b < n else []
                if cands and rng.random() < 0.85:
                    return cands[int(rng.integers(0, len(cands)))]
                e = min(n, b + int(rng.integers(1, 4)))
                if e <= b:
                    return None
                sub = chars[b:e]
                return Word(sub, sub, None, Unk, None, e - b, b, e, False)

            for _ in range(8):
                b = int(rng.integers(0, n))
                wi, wj = (None, bos) if b == 0 else (bos, None)
                if b == 0:
                    wk = pick(0)
                    if wk is None:
                        continue
                else:
                    wj = pick(b)
                    if wj is None or wj.e >= n:
                        continue
                    wk = pick(wj.e)
                    if wk is None:
                        continue
                    if rng.random() < 0.6 and wk.e < n:
                        nxt = pick(wk.e)
                        if nxt is not None:
                            wi, wj, wk = wj, wk, nxt
                for f in trigram_encoder(wi, wj, wk):
                    if len(feature_dic) < n_features:
                        add(f)
        if len(feature_dic) == before:
            break
    if vocab and len(feature_dic) < n_features:
        words = [w for w, _ in vocab]
        wtags = [t for _, t in vocab]
        nv = len(words)
        while len(feature_dic) < n_features:
            m = min(1 << 20, (n_features - len(feature_dic)) * 5 // 4 + 16)
            a = rng.integers(0, nv, m)
            b = rng.integers(0, nv, m)
            c = rng.integers(0, nv, m)
            tmpl = rng.integers(0, 4, m)
            for x, y, z, t in zip(a.tolist(), b.tolist(), c.tolist(), tmpl.tolist()):
                if t == 0:
                    f = (0, words[x], words[y], wtags[y])
                elif t == 1:
                    f = (7, words[x], words[y], words[z])
                elif t == 2:
                    f = (2, wtags[x], words[y], wtags[y])
                else:
                    f = (8, words[x], words[y])
                if f not in feature_dic:
                    feature_dic[f] = len(feature_dic)
                    if len(feature_dic) >= n_features:
                        break
    coefficients = rng.standard_normal(len(feature_dic)).astype(np.float64)
    return feature_dic, coefficients


def build_workload(name_or_cfg, rank=0, n_sent=None):
    """Dictionary, rules and sentences of a benchmark configuration (features come separately)."""
    cfg = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    if n_sent is not None:
        cfg['n_sent'] = n_sent
    alphabet = make_alphabet(cfg['alphabet'], seed=0)
    tag_to_morphs = make_dictionary(cfg['n_dict'], alphabet, seed=0)
    rules = make_rules(tag_to_morphs, alphabet, n_keys=min(5000, max(50, cfg['n_dict'] // 20)), seed=1)
    dictionary = MorphemeDictionary(tag_to_morphs, rules)
    sents = make_sentences(tag_to_morphs, rules, cfg['n_sent'], cfg['mean_len'], alphabet,
                           seed=2 + 1000 * rank, fixed_len=cfg['fixed_len'])
    return cfg, dictionary, sents
