import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def emu():
    """The CPU *emulation* of the CUDA kernels (tests/cuemu) -- test infrastructure,
    lets the GPU-less suite check kernel arithmetic against the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "cuemu"))
    import build_emu
    from rustpde_b200 import _ffi

    return _ffi.Lib(build_emu.build())


@pytest.fixture(scope="session")
def gpu():
    """The product CUDA library on cuda:0 (fails loudly if missing)."""
    from rustpde_b200 import _ffi

    return _ffi.product_lib()
