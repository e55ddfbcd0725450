"""Differential fuzzing of the kernel SOURCES against the oracle on the CPU (test infrastructure; not
collected by pytest): random dictionaries / rules / features / sentences per seed, every sentence's
lattice, best path, fp64 score and work counters compared bit for bit through the SIMT emulator, with
the beam sizes, sentence lengths, k-best survivors and lookup modes rotating over the seeds.

    python tests/fuzz_emulated.py FIRST_SEED LAST_SEED        # prints `ok SEED` / `FAIL SEED` + traceback
"""

import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import lattice_based_tagger_b200 as pkg                      # noqa: E402
from oracle import lattice_oracle as lo                      # noqa: E402
from tests import _cases, _checks, _emu                      # noqa: E402

BEAMS = [(1, 5), (2, 10), (3, 33), (7, 64), (16, 40)]
LENGTHS = (8, 25, 60, 110, 200)


def run_seed(seed):
    case = _checks.make_case(seed, n_sent=6 + seed % 13, max_sent_len=LENGTHS[seed % 5])
    dictionary, funcs = _cases.build_objects(case, pkg)
    tagger = pkg.Tagger(dictionary, score_funcs=funcs)
    oracle = lo.OracleTagger(dictionary, funcs)
    beams = BEAMS[(seed // 5) % 5]
    _checks.check_against_oracle(tagger, oracle, case['sentences'], beams, counters=True)
    if seed % 3 == 0:
        _checks.check_kbest(tagger, oracle, case['sentences'], (beams[0], 5))
    if seed % 7 == 0:
        _checks.check_lookup_modes(case)


def main(first, last):
    fails = 0
    with _emu.emulated():
        for seed in range(first, last):
            try:
                run_seed(seed)
                print('ok', seed, flush=True)
            except Exception:
                fails += 1
                print('FAIL', seed, flush=True)
                traceback.print_exc()
    print('done: %d seeds, %d failures' % (last - first, fails))
    return 1 if fails else 0


if __name__ == '__main__':
    sys.exit(main(int(sys.argv[1]), int(sys.argv[2])))
