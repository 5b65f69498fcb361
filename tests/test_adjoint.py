"""Navier2DAdjoint (navier_adjoint.rs:128-1068) on the device versus the oracle restatement; the four
fast-diagonalisation solvers (three Hholtz smoothers + pressure Poisson, plus the inner Navier2D's Poisson) use the
library's own eigen set-up, exported to the oracle.  CPU emulation here, the B200 in test_gpu_parity.py."""
import numpy as np
import pytest

import oracle as O
import parity_cases as pc
import rustpde_b200 as R


def check_adjoint(lib, periodic, nx, ny, steps=4, tol=1e-9):
    if periodic:
        a = R.Navier2DAdjoint.new_periodic(nx, ny, 1e4, 1.0, 0.01, 1.0, lib=lib)
        o = O.Navier2DAdjoint.new_periodic(nx, ny, 1e4, 1.0, 0.01, 1.0)
    else:
        a = R.Navier2DAdjoint.new(nx, ny, 1e4, 1.0, 0.01, 1.0, True, lib=lib)
        o = O.Navier2DAdjoint.new(nx, ny, 1e4, 1.0, 0.01, 1.0, True, eig=a.export_eig())
    for m in (a, o):
        m.set_velocity(0.2, 1.0, 1.0)
        m.set_temperature(0.2, 1.0, 1.0)
    for k in range(steps):
        a.update(1)
        o.update()
    assert abs(a.time - o.time) < 1e-12
    errs = {}
    for name, fa, fo in (("temp", a.temp, o.temp), ("ux", a.ux, o.ux), ("uy", a.uy, o.uy)):
        errs[name] = pc.rel(fa[0].vhat, fo[0].vhat)
        errs[name + "_res"] = pc.rel(fa[1].vhat, fo[1].vhat)
    errs["pres"] = pc.rel(a.pres[0].vhat, o.pres[0].vhat)
    assert max(errs.values()) <= tol, errs
    sm, un = a.residuals()
    osm, oun = o.residuals()
    for x, y in zip(sm + un, osm + oun):
        assert abs(x - y) <= 1e-9 * max(1.0, abs(y)), (sm, un, osm, oun)
    got = a.eval()
    ref = [o.eval_nu(), o.eval_nuvol(), o.eval_re(), o.div_norm()]
    for x, y in zip(got, ref):
        assert abs(x - y) <= 1e-9 * max(1.0, abs(y)), (got, ref)
    assert a.exit() == o.exit()
    return errs


@pytest.mark.parametrize("periodic,nx,ny", [(False, 24, 33), (True, 32, 33), (False, 20, 20)])
def test_adjoint(emu, periodic, nx, ny):
    check_adjoint(emu, periodic, nx, ny)


def test_adjoint_integrate(emu, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # (callback() may write to ./data)
    a = R.Navier2DAdjoint.new(24, 33, 1e4, 1.0, 0.01, 1.0, True, lib=emu)
    a.set_velocity(0.2, 1.0, 1.0)
    a.set_temperature(0.2, 1.0, 1.0)
    steps = R.integrate(a, 0.03, 0.02)
    assert steps == 3 and len(a.diagnostics["Nu"]) == 1
