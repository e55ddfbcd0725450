import sys
sys.path.insert(0, '/root/repo')
import lattice_based_tagger_b200 as pkg
from oracle import lattice_oracle as lo
from tests import _cases
for seed in (3001, 3002):
    case = _cases.random_case(seed, n_sent=10, features=True, prefs=True, max_sent_len=40)
    _cases.add_features(case, _cases.observed_features(case, lo, seed=seed), seed)
    d, f = _cases.build_objects(case, pkg)
    t = pkg.Tagger(d, score_funcs=f)
    o = lo.OracleTagger(d, f)
    for k in (1, 5, 12, 40):
        got = t.tag_batch(case['sentences'], k, errors='none')
        for s, g in zip(case['sentences'], got):
            try:
                w = o.tag(s, k)
            except IndexError:
                assert g is None; continue
            assert [tuple(x) for x in g.sequences] == w.words and g.score == w.score
    t.lattice_batch(case['sentences'])
print('sanitizer workload ok')
