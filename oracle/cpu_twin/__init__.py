"""ctypes wrapper of the C++ lane-parallel CPU restatement (oracle/cpu_twin/twin.cpp).

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py): built by `build()` (g++ -O3 -march=x86-64-v3 -fopenmp) into
oracle/_build/liboracle_cpu.so, timed by bench.py's `cpu_baseline` / `--impl reference` legs, checked against the
numpy oracle in tests/test_cpu_twin.py.  Never imported by rustpde_b200/.
"""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "twin.cpp")
LIB = os.path.join(os.path.dirname(HERE), "_build", "liboracle_cpu.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-fopenmp", "-fcx-limited-range", "-fno-math-errno", "-fPIC", "-shared", SRC, "-o", LIB, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr)
    return LIB


def _find_blas():
    for p in sys.path:
        for pat in ("scipy.libs/libscipy_openblas*.so", "numpy.libs/libscipy_openblas*.so"):
            hits = sorted(glob.glob(os.path.join(p, pat)))
            if hits:
                return hits[0]
    return None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError("cpu twin not built (oracle.cpu_twin.build())")
        lib = C.CDLL(LIB)
        lib.tw_create.restype = C.c_void_p
        lib.tw_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.tw_destroy.argtypes = [C.c_void_p]
        lib.tw_set_ics.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
        lib.tw_update.argtypes = [C.c_void_p, C.c_int]
        lib.tw_time.argtypes = [C.c_void_p]
        lib.tw_time.restype = C.c_double
        lib.tw_get_vhat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
        lib.tw_get_vhat.restype = C.c_longlong
        lib.tw_set_blas.argtypes = [C.c_char_p]
        lib.tw_set_threads.argtypes = [C.c_int]
        lib.tw_set_threads(int(os.environ.get("RUSTPDE_TWIN_THREADS", "0")))
        blas = _find_blas()
        if blas:
            lib.tw_set_blas(blas.encode())  # dgemm on one thread: the caller exports OPENBLAS_NUM_THREADS=1 (README.md:10-16)
        _lib = lib
    return _lib


def available():
    try:
        build()
        _load()
        return True
    except Exception:
        return False


def threads():
    return _load().tw_set_threads(0)


def set_threads(n):
    """Lane-parallel threads of the twin (rayon's default in the reference = all host CPUs).  torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would otherwise leave the twin on one thread."""
    return _load().tw_set_threads(int(n))


class TwinNavier:
    """Navier2D::new / new_periodic + set_velocity(0.2,1,1) + set_temperature(0.2,1,1) style runs on the CPU twin."""

    def __init__(self, nx, ny, ra, pr, dt, aspect, adiabatic, periodic, eig=None):
        lib = _load()
        self._lib = lib
        self.nx, self.ny, self.periodic = nx, ny, periodic
        args = [None, None, None]
        if not periodic:
            if eig is None:
                raise ValueError("confined twin needs eig=(lam, Q, P) (unshifted eigenvalues)")
            self._eig = [np.ascontiguousarray(e, dtype=np.float64) for e in eig]
            args = [e.ctypes.data for e in self._eig]
        self._h = lib.tw_create(nx, ny, ra, pr, dt, aspect, int(adiabatic), int(periodic), *args)
        if not self._h:
            raise RuntimeError("cpu twin: construction failed")

    def __del__(self):
        try:
            self._lib.tw_destroy(self._h)
        except Exception:
            pass

    def set_ics(self, amp_v=0.2, amp_t=0.2, m=1.0, n=1.0):
        self._lib.tw_set_ics(self._h, amp_v, amp_t, m, n)

    def update(self, nsteps=1):
        self._lib.tw_update(self._h, int(nsteps))

    @property
    def time(self):
        return self._lib.tw_time(self._h)

    def vhat(self, which):
        """0 temp, 1 ux, 2 uy, 3 pres, 4 pseudo pressure."""
        n = self._lib.tw_get_vhat(self._h, which, None, 0)
        buf = np.zeros(n)
        self._lib.tw_get_vhat(self._h, which, buf.ctypes.data, n)
        mx = (self.nx // 2 + 1) if self.periodic else (self.nx if which == 3 else self.nx - 2)
        a = buf.view(np.complex128) if self.periodic else buf
        return a.reshape(mx, -1)


def make_navier(wl, eig=None):
    """bench.py workload tuple -> twin with the benchmark's initial conditions; eig: (lam, Q, P) for confined runs
    (computed with scipy like the numpy oracle when not given)."""
    periodic, nx, ny, ra, pr, dt, aspect, adiabatic, _ = wl
    if not periodic and eig is None:
        import oracle as O
        o = O.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, banded=True)
        ts = o.solver[3].solver
        lam = ts.lam[0].copy()
        if abs(lam[0] + 1e-10) < 1e-10:
            lam = lam + 1e-10
        eig = (lam, ts.bwd[0], ts.fwd[0])
    t = TwinNavier(nx, ny, ra, pr, dt, aspect, adiabatic, periodic, eig)
    t.set_ics()
    return t
