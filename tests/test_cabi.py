"""The C ABI library loads and exports every symbol include/rustpde_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rustpde_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def product_so():
    from rustpde_b200 import build
    return build.build()


def test_header_matches_binding():
    from rustpde_b200 import _ffi
    assert declared_functions() == _ffi.EXPORTED_SYMBOLS


def test_product_library_exports_every_declared_symbol(product_so):
    out = subprocess.run(["nm", "-D", "--defined-only", product_so], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rp_[a-z0-9_]+)", out))
    missing = [f for f in declared_functions() if f not in exported]
    assert not missing, missing
    stray = [s for s in exported if s not in declared_functions()]
    assert not stray, stray


def test_product_library_loads_and_is_cuda_build(product_so):
    lib = ctypes.CDLL(product_so)
    assert lib.rp_version() >= 100
    assert lib.rp_is_emulated() == 0
    sass = subprocess.run(["cuobjdump", "-lelf", product_so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_product_path_fails_loudly_without_gpu(product_so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rustpde_b200 import _ffi
    with pytest.raises(Exception):
        _ffi.Lib(product_so)  # rp_init must fail: there is no CPU fallback


def test_package_never_references_oracle_or_emulator():
    pkg = os.path.join(ROOT, "rustpde_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn
                assert "librustpde_b200_emu" not in txt, fn
