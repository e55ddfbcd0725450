"""Builds tests/simt/liblt_emu.so: the CUDA sources of the package compiled with g++ against the
SIMT emulator (simt.h) — TEST INFRASTRUCTURE.  The product only ever loads liblt_b200.so."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'lattice_based_tagger_b200', 'csrc')
LIB = os.path.join(HERE, 'liblt_emu.so')


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh'))]
    out += [os.path.join(HERE, 'simt.h'), os.path.join(HERE, 'simt.cpp'), os.path.join(ROOT, 'include', 'lt_b200.h')]
    return out


def build(force=False, opt='-O1'):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in sources()):
        return LIB
    cxx = os.environ.get('CXX', 'g++')
    common = [cxx, '-std=c++17', opt, '-g', '-fPIC', '-DLT_SIMT_EMU', '-ffp-contract=off', '-fno-strict-aliasing',
              '-Wno-unknown-pragmas', '-Wno-attributes', '-I', HERE, '-I', CSRC]
    obj_k = os.path.join(HERE, 'lt_emu.o')
    obj_s = os.path.join(HERE, 'simt.o')
    subprocess.run(common + ['-x', 'c++', '-c', os.path.join(CSRC, 'lt_b200.cu'), '-o', obj_k], check=True)
    subprocess.run(common + ['-c', os.path.join(HERE, 'simt.cpp'), '-o', obj_s], check=True)
    subprocess.run([cxx, '-shared', '-o', LIB, obj_k, obj_s], check=True)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv))
