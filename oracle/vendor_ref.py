"""Vendors the UNMODIFIED reference package into oracle/_ref/ (test / baseline infrastructure).

    python oracle/vendor_ref.py

The reference (`/root/reference`, read-only, pure Python) exists only in the build container;
`oracle/_ref/` is git-ignored but travels to the GPU box with the repository snapshot, so that the
CPU baseline of `bench.py` (`cpu_baseline.kind = "reference"`, `--impl reference`) and the C1
configuration (the bundled `base` dictionary) can run the real thing there.  Nothing is edited: the
files are copied byte for byte, and nothing under `lattice_based_tagger_b200/` imports them.
`__graft_entry__.build()` runs this when the reference checkout is present.
"""

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/lattice_tagger'
DST = os.path.join(ROOT, 'oracle', '_ref', 'lattice_tagger')


def vendor(force=False):
    """-> path of the vendored package, or None when there is no reference checkout and no copy."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST) and not force:
        newest_src = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(SRC) for f in fs)
        newest_dst = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(DST) for f in fs)
        if newest_dst >= newest_src:
            return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    return DST


def import_reference():
    """Import the vendored reference (with the `numpy.int` alias its score function needs,
    beam/score_funcs.py:143); returns the module or None."""
    path = vendor()
    if path is None:
        return None
    parent = os.path.dirname(path)
    sys.dont_write_bytecode = True
    if parent not in sys.path:
        sys.path.append(parent)
    import numpy
    if not hasattr(numpy, 'int'):
        numpy.int = int
    import lattice_tagger
    return lattice_tagger


if __name__ == '__main__':
    print(vendor(force='--force' in sys.argv))
