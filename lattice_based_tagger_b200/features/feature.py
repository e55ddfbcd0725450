"""Feature templates of the trigram scorer (host side).

`trigram_encoder` states the nine feature templates of the reference
(`features/feature.py:76-121`) as Python tuples.  The decode path does not call it: the beam
kernel (`csrc/beam.cuh`) forms the same keys as 128-bit hashes and gathers their weights from the
device feature table.  It exists for the things that stay on the host — building a `feature_dic`
from tagged sequences, and the synthetic-workload generator — and it is the definition the
table compiler (`compile.py:pack_features`) parses keys against.
"""

from ..tagset import Adjective, Adverb, Noun, Unk, Verb

_CONTEXTUAL = frozenset((Noun, Adverb, Adjective, Verb))


def trigram_encoder(word_i, word_j, word_k):
    """Feature tuples of extending a hypothesis that ends in (word_i, word_j) with word_k.

    template  key                                  present when
    0         (wj.word, wk.word, wk.tag0)          always
    1         (wj.word, wk.tag0)                   always
    2         (wj.tag0, wk.word, wk.tag0)          always
    3         (wj.tag0, wk.tag0)                   always
    4         (wk.len,)                            always
    5         (wk.word, wk.tag0, wk.is_l)          always
    6         (min(8, wj.len),)                    wj is an unknown word
    7         (wi.word, wj.word, wk.word)          wi exists
    8         (w?.morph0, wk.morph0)               wk contextual and (wj contextual -> ?=j,
                                                   else wi exists and contextual -> ?=i)
    """
    tk = word_k.tag0
    tj = word_j.tag0
    out = [
        (0, word_j.word, word_k.word, tk),
        (1, word_j.word, tk),
        (2, tj, word_k.word, tk),
        (3, tj, tk),
        (4, word_k.len),
        (5, word_k.word, tk, word_k.is_l),
    ]
    if tj == Unk:
        out.append((6, min(8, word_j.len)))
    if word_i is not None:
        out.append((7, word_i.word, word_j.word, word_k.word))
    if tk in _CONTEXTUAL:
        if tj in _CONTEXTUAL:
            out.append((8, word_j.morph0, word_k.morph0))
        elif word_i is not None and word_i.tag0 in _CONTEXTUAL:
            out.append((8, word_i.morph0, word_k.morph0))
    return out


class WordsEncoder:
    """Encoder protocol (reference `features/feature.py:4-29`): `feature_dic` maps a feature
    tuple to its index in the coefficient vector; unseen tuples are dropped."""

    def __init__(self, feature_dic=None):
        self.feature_dic = feature_dic

    def is_trained(self):
        return self.feature_dic is not None

    def set_feature_dic(self, feature_dic):
        self.feature_dic = feature_dic
        return self


class SimpleTrigramEncoder(WordsEncoder):
    """`feature_dic` holder for the nine trigram templates
    (reference `features/feature.py:31-74`)."""

    def transform_word(self, word_i, word_j, word_k):
        features = trigram_encoder(word_i, word_j, word_k)
        if self.is_trained():
            features = [f for f in features if f in self.feature_dic]
        return features

    def transform_sequence(self, words):
        # words = [BOS, w1, ..., wn, EOS]; one feature list per real word
        previous = [None] + list(words)
        return [self.transform_word(wi, wj, wk)
                for wi, wj, wk in zip(previous, words, words[1:-1])]

    def encode_word(self, word_i, word_j, word_k):
        return [self.feature_dic[f] for f in self.transform_word(word_i, word_j, word_k)]

    def encode_sequence(self, words):
        if not self.is_trained():
            raise ValueError('Insert feature_dic first')
        return [[self.feature_dic[f] for f in features]
                for features in self.transform_sequence(words)]
