"""The C++ lane-parallel CPU restatement (oracle/cpu_twin, the CPU baseline bench.py times) against the numpy oracle:
same initial conditions, same eigen set-up data -> same fields to rounding."""
import numpy as np
import pytest

import oracle as O
from oracle import cpu_twin as T


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("periodic,nx,ny,steps", [(False, 24, 33, 6), (True, 32, 33, 6), (False, 64, 64, 20), (True, 64, 65, 20),
                                                  (False, 33, 40, 4), (True, 24, 21, 4)])
def test_twin_matches_numpy_oracle(periodic, nx, ny, steps):
    T.build()
    if periodic:
        o = O.Navier2D.new_periodic(nx, ny, 1e5, 1.0, 0.01, 1.0, banded=True)
        eig = None
    else:
        o = O.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, banded=True)
        ts = o.solver[3].solver
        lam = ts.lam[0].copy()
        if abs(lam[0] + 1e-10) < 1e-10:
            lam = lam + 1e-10
        eig = (lam, ts.bwd[0], ts.fwd[0])
    o.set_velocity(0.2, 1.0, 1.0)
    o.set_temperature(0.2, 1.0, 1.0)
    t = T.TwinNavier(nx, ny, 1e5, 1.0, 0.01, 1.0, True, periodic, eig)
    t.set_ics()
    for i, f in enumerate((o.temp, o.ux, o.uy)):
        assert rel(t.vhat(i), f.vhat) <= 1e-13
    for _ in range(steps):
        o.update()
    t.update(steps)
    assert abs(t.time - o.time) < 1e-12
    for i, f in enumerate((o.temp, o.ux, o.uy, o.pres[0])):
        assert rel(t.vhat(i), f.vhat) <= 1e-10, (i, rel(t.vhat(i), f.vhat))


def test_twin_reports_threads():
    T.build()
    assert T.threads() >= 1 and T.available()
