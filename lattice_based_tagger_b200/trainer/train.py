"""Trainer around the batched GPU decoder.

The reference's `train()` (`trainer/train.py:1-42`) scans the features of a tagged corpus, builds
`BeamScoreFunctions(regularity_func, score_func.set_encoder(encoder))`, creates a `Tagger` and
calls `fit_parameter` -> `train_epoch` — which is a stub there (`train.py:61-65` returns `coef`
unchanged, and `train()` itself stops on undefined names before it gets that far).  This module
keeps the reference's functions, arguments and result format (`{'idx_to_feature',
'coefficient'}`, `train.py:34-37`) and supplies the missing epoch: a structured perceptron whose
decoding step is `Tagger.tag_batch` — the whole corpus in one device pass per epoch.

    w <- w + phi(gold) - phi(predicted)          for every sentence whose prediction is not the gold path

with phi = counts of the trigram features of a path (`SimpleTrigramEncoder.encode_sequence`,
`features/feature.py:62-68`).  All sentences of an epoch are decoded with the same weights (the
updates of one epoch are summed), so an epoch is one `tag_batch` call plus host-side counting.
There is no CPU decoder: without the CUDA library `Tagger` raises.
"""

import numpy as np

from ..beam import BeamScoreFunctions
from ..dictionary.text import text_to_words
from ..features.utils import scan_features
from ..tagger import Tagger


def _surface(word_text):
    """'너무너무너무 는  아이오아이 의' -> '너무너무너무는 아이오아이의' (the string `Tagger.tag` receives)."""
    return ' '.join(eojeol.replace(' ', '') for eojeol in word_text.split('  '))


def _same_path(predicted, gold):
    """Word-by-word identity of two [BOS .. EOS] lists on what the annotation defines:
    span, morphemes and tags (`Word.len` and `Word.is_l` follow from those)."""
    if len(predicted) != len(gold):
        return False
    for p, g in zip(predicted, gold):
        if (p.b, p.e, p.morph0, p.morph1, p.tag0, p.tag1) != (g.b, g.e, g.morph0, g.morph1, g.tag0, g.tag1):
            return False
    return True


def train(word_morph_pairs, dictionary, encoder, score_func, regularity_func,
          max_epochs=100, min_feature_count=1, predefined_features=None,
          verbose=False, debug=False, beam_size=5, device=0):
    """Reference `train()` (`trainer/train.py:1-42`): -> {'idx_to_feature': [...], 'coefficient': [...]}"""
    word_morph_pairs = list(word_morph_pairs)
    if verbose:
        print('Scanning features ...')
    if predefined_features is None:
        # length-of-unknown-word features (template 6), lengths 1..8 (`train.py:9-12`)
        predefined_features = {(6, length): min_feature_count for length in range(1, 9)}

    idx_to_feature, feature_to_idx, idx_to_feature_count = scan_features(
        word_morph_pairs, encoder, min_count=min_feature_count, verbose=verbose, debug=debug,
        predefined_features=predefined_features)

    encoder.feature_dic = feature_to_idx
    funcs = BeamScoreFunctions(regularity_func, score_func.set_encoder(encoder))
    tagger = Tagger(dictionary, encoder=encoder, score_funcs=funcs, device=device)

    if verbose:
        print('Estimating parameter ...')
    try:
        coef = fit_parameter(word_morph_pairs, encoder, tagger, max_epochs, verbose=verbose, beam_size=beam_size)
    finally:
        tagger.close()

    params = {'idx_to_feature': idx_to_feature, 'coefficient': [float(c) for c in coef]}
    if debug:
        params['idx_to_feature_count'] = idx_to_feature_count
    return params


def fit_parameter(word_morph_pairs, encoder, tagger, max_epochs=100, verbose=False, beam_size=5, patience=5,
                  averaged=False):
    """Reference `fit_parameter()` (`train.py:44-59`).  Stops once an epoch makes no mistake, or when the
    number of mistakes has not improved for `patience` epochs (a corpus with an unreachable gold path — a word the
    dictionary cannot produce — never reaches zero); returns the best weights seen (`averaged=True`: the mean of
    the weights of all epochs, the averaged perceptron)."""
    coef = np.zeros(len(encoder.feature_dic), dtype=np.float64)
    best_loss, best_coef, stale = None, coef, 0
    total = np.zeros_like(coef)
    epochs = 0
    for epoch in range(1, max_epochs + 1):
        decoded_with = coef
        coef, loss = train_epoch(word_morph_pairs, encoder, tagger, coef, epoch, verbose, beam_size=beam_size)
        total += coef
        epochs += 1
        if best_loss is None or loss < best_loss:
            best_loss, best_coef, stale = loss, decoded_with, 0       # `loss` was measured with the weights the epoch started from
        else:
            stale += 1
        if loss == 0 or stale >= patience:
            break
    if averaged:
        return total / max(1, epochs)
    return best_coef


def _trigram_scorer(tagger):
    from ..beam import SimpleTrigramFeatureScore
    for func in tagger.score_funcs.funcs:
        if isinstance(func, SimpleTrigramFeatureScore):
            return func
    raise ValueError('the tagger has no SimpleTrigramFeatureScore to train')


def train_epoch(word_morph_pairs, encoder, tagger, coef, epoch, verbose, beam_size=5):
    """One perceptron pass (the reference's TODO, `train.py:61-65`): -> (coef, loss).

    loss = number of sentences whose best path under `coef` differs from the annotation.
    """
    coef = np.asarray(coef, dtype=np.float64).copy()
    scorer = _trigram_scorer(tagger)
    if scorer.encoder is encoder and scorer.coefficients is not None and len(scorer.coefficients) == len(coef) \
            and getattr(tagger, '_trained_features', None) is encoder.feature_dic:
        scorer.coefficients = coef
        tagger.update_weights()                        # same features as the device tables hold: weights in place
    else:
        scorer.set_encoder(encoder, coef)
        tagger.refresh()                               # first epoch: feature table compiled once
        tagger._trained_features = encoder.feature_dic

    golds, sents = [], []
    for word_text, morph_text in word_morph_pairs:
        try:
            gold = text_to_words(word_text, morph_text)
        except Exception:                              # malformed pairs are skipped, as scan_features does
            continue
        golds.append(gold)
        sents.append(_surface(word_text))

    predicted = tagger.tag_batch(sents, beam_size=beam_size, errors='none')
    update = np.zeros_like(coef)
    loss = 0
    for gold, seq in zip(golds, predicted):
        if seq is not None and _same_path(seq.sequences, gold):
            continue
        loss += 1
        for idxs in encoder.encode_sequence(gold):
            for idx in idxs:
                update[idx] += 1.0
        if seq is not None:
            for idxs in encoder.encode_sequence(seq.sequences):
                for idx in idxs:
                    update[idx] -= 1.0
    coef += update
    if verbose:
        print('epoch %d: %d of %d sentences differ from the annotation' % (epoch, loss, len(golds)))
    return coef, loss
