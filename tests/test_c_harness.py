"""A gcc-compiled C program (tests/c_harness/harness.c) that includes include/rustpde_b200.h and calls the C ABI
directly -- no ctypes in between -- checked against the oracle.

* not gpu: the harness compiles and links against the product library (every symbol it uses resolves), and runs to
  completion against the CPU emulation build of the same sources.
* gpu: it runs against librustpde_b200.so on the B200 and its numbers match the oracle (config 1, 5 steps:
  Nu / Nuvol / Re / Ekin <= 1e-9 relative; Hholtz::new2 solve against the analytic solution <= 1e-3, the reference's
  own tolerance for examples/hholtz_2d.rs)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_harness", "harness.c")


def build(tmp_path, libdir, libname):
    exe = str(tmp_path / ("harness_" + libname))
    cmd = ["gcc", "-O2", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run_and_check(exe, tmp_path):
    from rustpde_b200 import _ffi
    import oracle as O
    eig_file = str(tmp_path / "eig.bin")
    r = subprocess.run([exe, _ffi.find_lapack() or "", eig_file], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "HARNESS_OK" in out
    assert "shape_error_code 2" in out  # RP_ERR_SHAPE
    vals = dict(zip(*[iter(re.search(r"navier (.*)", out).group(1).split())] * 2))
    raw = np.fromfile(eig_file)
    m = 62
    eig = (raw[:m], raw[m:m + m * m].reshape(m, m), raw[m + m * m:].reshape(m, m))
    o = O.Navier2D.new(64, 64, 1e5, 1.0, 0.02, 1.0, True, banded=True, eig_data=eig)
    o.set_velocity(0.2, 1.0, 1.0)
    o.set_temperature(0.2, 1.0, 1.0)
    for _ in range(5):
        o.update()
    ref = {"nu": o.eval_nu(), "nuvol": o.eval_nuvol(), "re": o.eval_re(), "ekin": o.eval_ekin(), "time": o.time}
    for k, b in ref.items():
        a = float(vals[k])
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (k, a, b)
    tv = [float(x) for x in re.search(r"temp_vhat 62 62 (.*)", out).group(1).split()]
    assert np.abs(np.array(tv).reshape(4, 4) - o.temp.vhat[:4, :4]).max() <= 1e-9 * np.abs(o.temp.vhat).max()
    err = float(re.search(r"analytic_max_abs_err (\S+)", out).group(1))
    assert err < 1e-3, err
    return True


def test_c_harness_links_against_product_library(tmp_path):
    build(tmp_path, os.path.join(ROOT, "rustpde_b200"), "rustpde_b200")


def test_c_harness_runs_on_emulation(tmp_path, emu):
    libdir = os.path.dirname(emu.path)
    assert run_and_check(build(tmp_path, libdir, "rustpde_b200_emu"), tmp_path)


@pytest.mark.gpu
def test_c_harness_on_gpu(tmp_path, gpu):
    exe = build(tmp_path, os.path.join(ROOT, "rustpde_b200"), "rustpde_b200")
    assert run_and_check(exe, tmp_path)
    # the process really loaded the CUDA library
    r = subprocess.run(["ldd", exe], capture_output=True, text=True)
    assert "librustpde_b200.so" in r.stdout
