"""The trainer's weight format (reference `trainer/train.py:34-37`).

`train()` in the reference returns `{'idx_to_feature': [feature tuple, ...], 'coefficient':
[float, ...]}`; index i of both lists belongs together.  `load_params` turns such a dict into the
scorer the tagger consumes, `dump_params` is its inverse.  (The reference's `train_epoch` is a
stub, `train.py:61-65`; fitting weights is outside the decode path.)
"""

import numpy as np

from ..beam import SimpleTrigramFeatureScore
from ..features import SimpleTrigramEncoder


def load_params(params):
    """`{'idx_to_feature', 'coefficient'}` -> `SimpleTrigramFeatureScore` ready for `Tagger`."""
    idx_to_feature = params['idx_to_feature']
    coefficient = np.asarray(params['coefficient'], dtype=np.float64)
    if len(idx_to_feature) != len(coefficient):
        raise ValueError('idx_to_feature and coefficient have different lengths')
    feature_dic = {tuple(feature): idx for idx, feature in enumerate(idx_to_feature)}
    return SimpleTrigramFeatureScore(SimpleTrigramEncoder(feature_dic), coefficient)


def dump_params(score):
    """Inverse of `load_params`."""
    feature_dic = score.encoder.feature_dic
    idx_to_feature = [None] * len(feature_dic)
    for feature, idx in feature_dic.items():
        idx_to_feature[idx] = feature
    return {'idx_to_feature': idx_to_feature, 'coefficient': [float(c) for c in score.coefficients]}
