import sys, time, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
import lattice_based_tagger_b200 as pkg
from lattice_based_tagger_b200 import synth, _native
from lattice_based_tagger_b200.tagger.tagger import pack_sentences
import bench
cfg, dictionary, sents = synth.build_workload('c2')
reg = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore())
rt = pkg.Tagger(dictionary, score_funcs=reg)
st = rt.tag_batch_packed(sents, 5)[3]
good = [i for i in range(len(sents)) if st[i] == 0]
for i in range(len(sents)):
    if st[i] != 0: sents[i] = sents[good[i % len(good)]]
fd, coef = synth.make_features(bench.feature_sample(sents), lambda s: rt.tag_batch(s, 5, errors='none'), rt.lattice_batch, cfg['n_feat'], list(dictionary.tag_to_morphs), seed=3)
funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore(), pkg.beam.SimpleTrigramFeatureScore(pkg.features.SimpleTrigramEncoder(fd), coef))
t = pkg.Tagger(dictionary, score_funcs=funcs)
text, off = pack_sentences(sents)
n = len(sents); nu = int(off[-1])
def pinned(nb): return torch.empty(nb, dtype=torch.uint8).pin_memory()
h_text = pinned(text.nbytes); h_text.numpy()[:] = text.view(np.uint8)
h_off = pinned(off.nbytes); h_off.numpy()[:] = off.view(np.uint8)
h_poff = pinned(4*(n+1)); h_edges = pinned(16*nu); h_scores = pinned(8*n); h_status = pinned(4*n)
lib, b = t._lib, t._batch
def step():
    _native.check(lib.lt_tag_batch_host(b, ctypes.c_void_p(h_text.data_ptr()), ctypes.c_void_p(h_off.data_ptr()), n, 5,
        ctypes.c_void_p(h_poff.data_ptr()), ctypes.c_void_p(h_edges.data_ptr()), nu, ctypes.c_void_p(h_scores.data_ptr()), ctypes.c_void_p(h_status.data_ptr())))
t.timings()
for _ in range(5): step()
acc = {}
N = 30
t0 = time.perf_counter()
for _ in range(N):
    step()
    for k, v in t.timings().items(): acc[k] = acc.get(k, 0) + v
wall = (time.perf_counter() - t0) / N * 1e3
print('wall ms/step (incl timings call)', wall, {k: round(v / N, 4) for k, v in acc.items()})
t0 = time.perf_counter()
for _ in range(N): step()
print('wall ms/step', (time.perf_counter() - t0) / N * 1e3)
