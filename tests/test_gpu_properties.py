"""Full-size checks of the benchmark workload (BASELINE config C2: 10k sentences, 100k-morpheme
dictionary, beam 5) through properties that do not need the oracle's beam search:

* every returned path tiles its sentence exactly (contiguous [b, e) spans from 0 to L);
* the returned score equals the score of that path re-evaluated word by word with the score
  program (the oracle's transition increment, i.e. the reference's arithmetic), bit for bit;
* the lattice's edges are dictionary facts: every edge's morphemes are in the dictionary sets;
* results do not depend on how the batch is cut (single call == two half batches == reversed order)
  and repeat exactly from run to run;
* the device work counters are consistent with the outputs (W = words returned, E = edges returned).
"""

import numpy as np
import pytest

import lattice_based_tagger_b200 as pkg
from lattice_based_tagger_b200 import synth
from oracle import lattice_oracle as lo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def workload():
    cfg, dictionary, sents = synth.build_workload('c2')
    reg = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore())
    reg_tagger = pkg.Tagger(dictionary, score_funcs=reg)
    status = reg_tagger.tag_batch_packed(sents, cfg['beam'])[3]
    good = [i for i in range(len(sents)) if status[i] == 0]
    for i in range(len(sents)):
        if status[i] != 0:
            sents[i] = sents[good[i % len(good)]]
    sample = sents[:1000]
    feature_dic, coef = synth.make_features(
        sample, lambda s: reg_tagger.tag_batch(s, cfg['beam'], errors='none'), reg_tagger.lattice_batch,
        cfg['n_feat'], list(dictionary.tag_to_morphs), seed=3)
    reg_tagger.close()
    funcs = pkg.beam.BeamScoreFunctions(
        pkg.beam.RegularizationScore(),
        pkg.beam.SimpleTrigramFeatureScore(pkg.features.SimpleTrigramEncoder(feature_dic), coef))
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    return cfg, dictionary, funcs, sents, tagger


def test_paths_tile_sentences_and_counters_agree(workload):
    cfg, dictionary, funcs, sents, tagger = workload
    path_off, edges, scores, status = tagger.tag_batch_packed(sents, cfg['beam'])
    assert (status == 0).all()
    counters = tagger.counters()
    assert counters['sentences'] == len(sents)
    assert counters['W'] == len(edges) == path_off[-1]
    lengths = np.array([len(s.replace(' ', '')) for s in sents])
    assert counters['L'] == lengths.sum()
    b, e = edges['b'].astype(np.int64), edges['e'].astype(np.int64)
    first = path_off[:-1]
    last = path_off[1:] - 1
    assert (b[first] == 0).all()
    assert (e[last] == lengths).all()
    inner = np.ones(len(edges), dtype=bool)
    inner[first] = False
    assert (b[inner] == e[np.nonzero(inner)[0] - 1]).all()          # contiguous
    assert ((e - b) >= 1).all() and ((e - b) <= 8).all()             # beam window (beam.py:30)
    assert np.isfinite(scores).all()


def test_scores_equal_reevaluated_paths(workload):
    cfg, dictionary, funcs, sents, tagger = workload
    program = lo.ScoreProgram(funcs)
    sample = sents[::20]                                              # 500 sentences
    for sent, seq in zip(sample, tagger.tag_batch(sample, cfg['beam'])):
        words = [tuple(w) for w in seq.sequences]
        score = 0
        for i in range(1, len(words) - 1):
            word_i = None if i == 1 else words[i - 2]
            score = score + program.increment(word_i, words[i - 1], words[i])
        assert score == seq.score, sent


def test_lattice_edges_are_dictionary_facts(workload):
    cfg, dictionary, funcs, sents, tagger = workload
    sample = sents[:300]
    for sent, (words, bindex) in zip(sample, tagger.lattice_batch(sample)):
        chars = sent.replace(' ', '')
        for w in words[1:-1]:
            assert w.word == chars[w.b:w.e]
            if w.morph1 is None:
                assert w.morph0 in dictionary.tag_to_morphs[w.tag0]
            else:
                assert w.tag1 == 'Eomi' and w.morph1 in dictionary.eomis
                assert w.morph0 in (dictionary.adjectives if w.tag0 == 'Adjective' else dictionary.verbs)


def test_batch_composition_and_repeatability(workload):
    cfg, dictionary, funcs, sents, tagger = workload
    k = cfg['beam']
    whole = tagger.tag_batch_packed(sents, k)
    again = tagger.tag_batch_packed(sents, k)
    for a, b in zip(whole, again):
        assert (a == b).all()
    half = len(sents) // 2
    lo_part = tagger.tag_batch_packed(sents[:half], k)
    hi_part = tagger.tag_batch_packed(sents[half:], k)
    assert (np.concatenate([lo_part[2], hi_part[2]]) == whole[2]).all()
    assert (np.concatenate([lo_part[1], hi_part[1]]) == whole[1]).all()
    rev = tagger.tag_batch_packed(sents[::-1], k)
    assert (rev[2][::-1] == whole[2]).all()


def test_other_beam_sizes_keep_invariants(workload):
    cfg, dictionary, funcs, sents, tagger = workload
    sample = sents[:2000]
    lengths = np.array([len(s.replace(' ', '')) for s in sample])
    for k in (1, 10, 32, 64):
        path_off, edges, scores, status = tagger.tag_batch_packed(sample, k)
        assert (status == 0).all()
        assert (edges['e'][path_off[1:] - 1] == lengths).all()
        assert (edges['b'][path_off[:-1]] == 0).all()


def test_small_div_exhaustive(tmp_path):
    """The lattice kernel splits flat item indices with a float multiply by the hardware's approximate reciprocal
    (lattice.cuh: small_div / small_rcp) and decodes triangular indices through an approximate square root: both
    are checked against integer arithmetic over their whole range on the device they run on."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('no nvcc on this box')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'check_small_div')
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17',
                    '-I', os.path.join(root, 'lattice_based_tagger_b200', 'csrc'), '-o', exe,
                    os.path.join(root, 'tests', 'cuda', 'check_small_div.cu')], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and 'mismatches 0' in out.stdout, out.stdout + out.stderr
