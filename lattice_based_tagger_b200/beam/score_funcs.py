"""Score-function descriptors (host side).

Same classes, constructor arguments and attributes as the reference's additive beam scorers
(`beam/score_funcs.py:7-144`).  In the reference each object scores one (hypothesis, word) pair
in Python; here the objects only *describe* the scorer.  `compile.py:pack_score_program` turns a
`BeamScoreFunctions` into an ordered device program (the order fixes the fp64 association,
SURVEY App. A Q6) and `csrc/beam.cuh` evaluates it for every transition.  There is no host
scorer: a `BeamScoreFunction` subclass the device cannot express is rejected when the tagger is
built.
"""

import numpy as np


class BeamScoreFunction:
    """Base class of the additive scorers (reference `score_funcs.py:7-15`)."""

    def __call__(self, sequence, word_k):
        return self.score(sequence, word_k)

    def score(self, seq, word_k):
        raise NotImplementedError(
            'transitions are scored on the device (csrc/beam.cuh); there is no host scorer')

    def evaluate(self, seq):
        raise NotImplementedError(
            'transitions are scored on the device (csrc/beam.cuh); there is no host scorer')


class BeamScoreFunctions:
    """Ordered sum of scorers (reference `score_funcs.py:18-54`)."""

    def __init__(self, *functions):
        for func in functions:
            if not _is_score_function(func):
                raise ValueError('functions must be instance of BeamScoreFunction')
        self.funcs = list(functions)


def _is_score_function(func):
    if isinstance(func, BeamScoreFunction):
        return True
    # the reference's own scorer objects are accepted as descriptors as well
    return any(c.__name__ == 'BeamScoreFunction' for c in type(func).__mro__)


class RegularizationScore(BeamScoreFunction):
    """Length prior: `known_preference * len` for dictionary words,
    `unknown_penalty * (len + 0.1)` for unknown words, plus `syllable_penalty` for one-syllable
    nouns (reference `score_funcs.py:56-73`)."""

    def __init__(self, unknown_penalty=-0.1, known_preference=0.2, syllable_penalty=-0.2):
        self.unknown_penalty = unknown_penalty
        self.known_preference = known_preference
        self.syllable_penalty = syllable_penalty


class MorphemePreferenceScore(BeamScoreFunction):
    """User bonus per (tag, morpheme), applied to both morphemes of a word
    (reference `score_funcs.py:75-88`)."""

    def __init__(self, tag_to_morph=None):
        self.tag_to_morph = {} if tag_to_morph is None else tag_to_morph


class WordPreferenceScore(BeamScoreFunction):
    """User bonus per (tag, surface word) (reference `score_funcs.py:90-100`)."""

    def __init__(self, tag_to_word=None):
        self.tag_to_word = {} if tag_to_word is None else tag_to_word


class SimpleTrigramFeatureScore(BeamScoreFunction):
    """Sparse linear scorer over the trigram templates (reference `score_funcs.py:102-144`).

    `coefficients[i]` is the weight of the feature whose `encoder.feature_dic` value is `i`;
    together they are the trainer's weight format (`trainer/train.py:34-37`).
    """

    def __init__(self, encoder=None, coefficients=None):
        self.set_encoder(encoder, coefficients)

    def set_encoder(self, encoder, coefficients=None):
        if encoder is None:
            self.num_features = 0
            self.coefficients = None
            self.encoder = encoder
            return self
        if not encoder.is_trained():
            raise ValueError('Encoder must be trained first')
        self.num_features = len(encoder.feature_dic)
        if coefficients is None:
            coefficients = np.zeros(self.num_features)
        if len(coefficients) != self.num_features:
            raise ValueError('Encoder and coefficients have different size features')
        self.coefficients = coefficients
        self.encoder = encoder
        return self
