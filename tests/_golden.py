"""Loader for tests/golden/*.json.gz (written by oracle/gen_golden.py from the reference)."""

import glob
import gzip
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def names():
    return sorted(os.path.basename(p)[:-len('.json.gz')] for p in glob.glob(os.path.join(GOLDEN, '*.json.gz')))


def _feature_key(key):
    return tuple(key)


def load(name):
    with gzip.open(os.path.join(GOLDEN, name + '.json.gz'), 'rb') as f:
        payload = json.loads(f.read().decode('utf-8'))
    case = payload['case']
    case['feature_keys'] = [_feature_key(k) for k in case['feature_keys']]
    case['coefficients'] = [float.fromhex(c) for c in case['coefficients']]
    case['rules'] = {k: [tuple(c) for c in v] for k, v in case['rules'].items()}
    return payload


def edge(fields):
    return tuple(fields)


def expected_survivors(entry, k):
    got = entry['beams'][str(k)]
    if got == 'IndexError':
        return None
    return [([edge(w) for w in m['words']], float.fromhex(m['score']), m['num_unk']) for m in got]
