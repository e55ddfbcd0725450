"""Helper of the CPU test suite: runs the package's CUDA kernel SOURCES on the host through the
SIMT emulator (tests/simt) — test infrastructure only.

`emulated()` is a context manager that makes `lattice_based_tagger_b200._native.load()` hand out
`tests/simt/liblt_emu.so` (the same `csrc/*.cu*` compiled with g++ against `simt.h`) while the
block runs, and restores the real state afterwards.  The product has no hook for this: the swap is
done by patching the module attribute from the outside.
"""

import contextlib
import ctypes
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def _builder():
    spec = importlib.util.spec_from_file_location('_simt_build', os.path.join(HERE, 'simt', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_emu_lib = None


def emu_lib():
    global _emu_lib
    if _emu_lib is None:
        from lattice_based_tagger_b200 import _native
        lib = ctypes.CDLL(_builder().build())
        for name, (restype, argtypes) in _native.SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _emu_lib = lib
    return _emu_lib


@contextlib.contextmanager
def emulated():
    from lattice_based_tagger_b200 import _native
    saved = _native._lib
    _native._lib = emu_lib()
    try:
        yield
    finally:
        _native._lib = saved
