"""Corpus scanners (host side, training time): tagged corpus -> `feature_to_idx` / dictionary.

Counterparts of the reference's `scan_features` and `scan_dictionary`
(`features/utils.py:9-73`).  `scan_features` defines the index order of the trainer's weight
format — template id, then descending count, then the first field (`features/utils.py:50-51`) —
so weights trained against the reference line up with `feature_dic` built here.
"""

from collections import defaultdict

from ..dictionary.text import flatten_words, text_to_words


def scan_features(word_morph_pairs, encoder, min_count=1, predefined_features=None, verbose=False,
                  debug=False, flatten=False):
    """-> (idx_to_feature, feature_to_idx, idx_to_count)"""
    counter = defaultdict(int, predefined_features or {})
    for word_text, morph_text in word_morph_pairs:
        try:
            words = text_to_words(word_text, morph_text)
            if flatten:
                words = flatten_words(words)
            for features in encoder.transform_sequence(words):
                for feature in features:
                    counter[feature] += 1
        except Exception as exc:        # malformed pairs are skipped, as in the reference
            if debug:
                print('%s\nword_text : %s\nmorph_text : %s\n' % (exc, word_text, morph_text))
    kept = {f: c for f, c in counter.items() if c >= min_count}
    if verbose:
        print('scanned %d features' % len(kept))
    idx_to_feature = [f for f, _ in sorted(kept.items(), key=lambda fc: (fc[0][0], -fc[1], fc[0][1]))]
    idx_to_count = [kept[f] for f in idx_to_feature]
    feature_to_idx = {f: i for i, f in enumerate(idx_to_feature)}
    return idx_to_feature, feature_to_idx, idx_to_count


def scan_dictionary(word_morph_pairs, min_count=1):
    """-> ({tag: set(morph)}, {(morph, tag): count}) from the morpheme column of a corpus."""
    counter = defaultdict(int)
    for _, morph_text in word_morph_pairs:
        for eojeol in morph_text.split():
            for morph in eojeol.split('+'):
                morph = morph.strip()
                if len(morph) >= 3:
                    counter[morph] += 1
    counts = {tuple(k.split('/', 1)): c for k, c in counter.items() if c >= min_count}
    tag_to_morphs = defaultdict(set)
    for morph, tag in counts:
        tag_to_morphs[tag].add(morph)
    return dict(tag_to_morphs), counts
