import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_ROOT = '/root/reference'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'lattice_tagger'))


def import_reference():
    """Import the read-only reference package (only present in the build container)."""
    if not have_reference():
        pytest.skip('reference checkout not present on this machine')
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    import numpy
    if not hasattr(numpy, 'int'):
        numpy.int = int            # beam/score_funcs.py:143 uses the alias numpy removed in 1.24
    import lattice_tagger
    return lattice_tagger
