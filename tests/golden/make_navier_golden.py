"""Generates the 1000-step Navier2D golden fixtures of tests/golden/ from the numpy oracle (oracle/, the CPU
restatement of navier.rs:737-765).  Run here (no GPU needed):

    python tests/golden/make_navier_golden.py [confined128] [periodic512]

* navier_confined128_1000.npz -- `Navier2D::new(128, 129, Ra=1e5, Pr=1, dt=0.01, aspect=1, adiabatic)`,
  ICs set_velocity(0.2,1,1) + set_temperature(0.2,1,1), 1000 steps.  The pressure-Poisson eigen set-up data
  (lam, Q, P) are part of the fixture: they come from the library's own even/odd block diagonalisation
  (exactly checkerboard, so the device runs the parity-split GEMMs), computed once here through the CPU
  emulation build of the library, and are consumed by BOTH the oracle (now) and the device (in the test) --
  end-to-end parity must not depend on which LAPACK build produced the eigenvectors.
* navier_periodic512_1000.npz -- BASELINE config 3: `new_periodic(512, 513, Ra=1e7, Pr=1, dt=2e-3, aspect=1)`,
  same ICs, 1000 steps (no eigen data on this path).

Stored: observables [Nu, Nuvol, Re, |div|, Ekin] (functions.rs:12-101, average.rs:25-57) at steps 250, 500,
750, 1000, and the low-mode 24x24 blocks plus a strided sample of the final temp / ux / uy / pres coefficients.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cuemu")):
    sys.path.insert(0, p)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import oracle as O  # noqa: E402

CHECKPOINTS = (250, 500, 750, 1000)
CASES = {
    "confined128": dict(periodic=False, nx=128, ny=129, ra=1e5, pr=1.0, dt=0.01, aspect=1.0, adiabatic=True),
    "periodic512": dict(periodic=True, nx=512, ny=513, ra=1e7, pr=1.0, dt=2e-3, aspect=1.0, adiabatic=True),
}


def sample(a):
    """Low-mode block and a strided sample of a coefficient array (what the test compares)."""
    return a[:24, :24].copy(), a[::7, ::5].copy()


def run(name):
    c = CASES[name]
    out = {}
    if c["periodic"]:
        o = O.Navier2D.new_periodic(c["nx"], c["ny"], c["ra"], c["pr"], c["dt"], c["aspect"], banded=True)
    else:
        import build_emu
        import rustpde_b200 as R
        from rustpde_b200 import _ffi
        emu = _ffi.Lib(build_emu.build())
        n = R.Navier2D.new(c["nx"], c["ny"], c["ra"], c["pr"], c["dt"], c["aspect"], c["adiabatic"], lib=emu)
        lam, q, p = n.export_eig()  # unshifted eigenvalues, exactly checkerboard Q and P
        del n
        out.update(eig_lam=lam, eig_q=q, eig_p=p)
        o = O.Navier2D.new(c["nx"], c["ny"], c["ra"], c["pr"], c["dt"], c["aspect"], c["adiabatic"], banded=True, eig_data=(lam, q, p))
    o.set_velocity(0.2, 1.0, 1.0)
    o.set_temperature(0.2, 1.0, 1.0)
    obs = []
    t0 = time.time()
    for step in range(1, CHECKPOINTS[-1] + 1):
        o.update()
        if step in CHECKPOINTS:
            obs.append([o.eval_nu(), o.eval_nuvol(), o.eval_re(), o.div_norm(), o.eval_ekin()])
            print(name, step, obs[-1], "%.0fs" % (time.time() - t0), flush=True)
    out["checkpoints"] = np.array(CHECKPOINTS)
    out["observables"] = np.array(obs)
    out["time"] = np.array(o.time)
    for fname, f in (("temp", o.temp), ("ux", o.ux), ("uy", o.uy), ("pres", o.pres[0])):
        lo, st = sample(f.vhat)
        out[fname + "_low"] = lo
        out[fname + "_strided"] = st
    out["params"] = np.array([c["nx"], c["ny"], c["ra"], c["pr"], c["dt"], c["aspect"], float(c["adiabatic"])])
    np.savez_compressed(os.path.join(HERE, "navier_%s_1000.npz" % name), **out)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        run(nm)
