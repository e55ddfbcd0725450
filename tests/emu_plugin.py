"""pytest plugin (test infrastructure): runs the `-m gpu` parity tests where no GPU exists, through the
SIMT emulator build of the kernel SOURCES (tests/_emu.py) instead of `liblt_b200.so`:

    python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_trainer.py \
        -m gpu -p tests.emu_plugin -n 8 -k "not small_div"

(`test_small_div_exhaustive` checks the hardware's approximate reciprocal / square root and needs the
device; `test_config_samples` takes ~3 min here.  With `LT_SIMT_MEMCHECK=1` in the environment every device
buffer and every launch's shared memory end at a guard page, with `LT_SIMT_FILL=0x00` / `0xFF` fresh memory
holds another pattern: tests/simt/simt.h.)  Only loaded when named with `-p`: the default runs —
`-m "not gpu"` here, `-m gpu` on the GPU box — never see it, and a pass through it is a statement about
the kernels' logic, not about the device build.
"""

import pytest


@pytest.fixture(autouse=True, scope='session')
def _emulated_session():
    from tests import _emu
    with _emu.emulated():
        yield
