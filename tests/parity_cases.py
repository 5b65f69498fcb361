"""Parity cases shared by the CPU-emulation tests (test_emu_parity.py) and the
GPU tests (test_gpu_parity.py): the device path behind the C ABI versus the
numpy oracle, on the same seeded inputs.

Tolerance (north_star): <= 1e-10 relative per transform / solve, measured as
max|a-b| / max|b|.  The fast-diagonalisation solves are compared with the
eigen set-up data (lam, Q, P) SHARED between oracle and device;
with independently computed set-up data the comparison is made against the
reference algorithm's own noise floor instead.
"""
import numpy as np

import oracle as O
import rustpde_b200 as R

TOL = 1e-10

RB = {"chebyshev": R.chebyshev, "cheb_dirichlet": R.cheb_dirichlet, "cheb_neumann": R.cheb_neumann, "fourier_r2c": R.fourier_r2c}
OB = {"chebyshev": O.chebyshev, "cheb_dirichlet": O.cheb_dirichlet, "cheb_neumann": O.cheb_neumann, "fourier_r2c": O.fourier_r2c}


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def smooth_field(x, y):
    """SURVEY 8d per-kernel input: sin(3x~+0.3) cos(2y~) + 0.1 x~ y~^2 on normalised coords."""
    xs = (x - x[0]) / (x[-1] - x[0])
    ys = (y - y[0]) / (y[-1] - y[0])
    return np.sin(3.0 * xs + 0.3)[:, None] * np.cos(2.0 * ys)[None, :] + 0.1 * xs[:, None] * (ys ** 2)[None, :]


def make_fields(lib, kx, nx, ky, ny):
    f = R.Field2(R.Space2(RB[kx](nx), RB[ky](ny)), lib=lib)
    of = O.Field2(O.Space2(OB[kx](nx), OB[ky](ny)))
    return f, of


def check_field_ops(lib, kx, nx, ky, ny, seed=1234, tol=TOL, grads=((1, 0), (0, 1), (2, 0), (0, 2), (1, 1))):
    """forward / backward / to_ortho / from_ortho / gradient of Field2 (src/field.rs:103-129)."""
    rng = np.random.default_rng(seed)
    f, of = make_fields(lib, kx, nx, ky, ny)
    out = {}
    for name, v in (("white", rng.uniform(-1, 1, (nx, ny))), ("smooth", smooth_field(of.x[0][:nx], of.x[1]))):
        f.v = v
        of.v = v.copy()
        f.forward()
        of.forward()
        out["forward_" + name] = rel(f.vhat, of.vhat)
        f.backward()
        of.backward()
        out["backward_" + name] = rel(f.v, of.v)
    out["to_ortho"] = rel(f.to_ortho(), of.to_ortho())
    g = rng.uniform(-1, 1, f.shape_ortho)
    if f.is_complex:
        g = g + 1j * rng.uniform(-1, 1, f.shape_ortho)
    f.from_ortho(g)
    of.from_ortho(g)
    out["from_ortho"] = rel(f.vhat, of.vhat)
    for d in grads:
        for sc in (None, [1.5, 0.7]):
            out["gradient_%d%d_%s" % (d[0], d[1], "s" if sc else "n")] = rel(f.gradient(list(d), sc), of.gradient(list(d), sc))
    bad = {k: v for k, v in out.items() if not v <= tol}
    assert not bad, bad
    return out


def check_adi(lib, nx, ny, c=(0.3, 0.7), seed=7, tol=TOL):
    rng = np.random.default_rng(seed)
    f, of = make_fields(lib, "cheb_dirichlet", nx, "cheb_dirichlet", ny)
    s, os_ = R.HholtzAdi(f, c), O.HholtzAdi(of, c, banded=True)
    b = rng.uniform(-1, 1, (nx, ny))
    e1 = rel(s.solve(b), os_.solve(b))
    bc = b + 1j * rng.uniform(-1, 1, (nx, ny))
    e2 = rel(s.solve(bc), os_.solve(bc))
    assert e1 <= tol and e2 <= tol, (e1, e2)
    return e1, e2


def check_tensor_shared_eig(lib, which, kx, ky, nx, ny, c=(1.0, 0.8), alpha=2.0, seed=11, tol=TOL):
    """Hholtz / Poisson with Chebyshev x: strict mode -- oracle and device consume the same (lam, Q, P)."""
    rng = np.random.default_rng(seed)
    f, of = make_fields(lib, kx, nx, ky, ny)
    if which == "poisson":
        osol = O.Poisson(of, c, banded=True)
        lam = osol.solver.lam[0].copy()
        # hand the device the pre-shift eigenvalues; it applies poisson.rs:80-83 itself
        if abs(lam[0] + 1e-10) < 1e-10:
            lam = lam + 1e-10
        sol = R.Poisson(f, c, eig=(lam, osol.solver.bwd[0], osol.solver.fwd[0]))
    else:
        osol = O.Hholtz(of, c, alpha=alpha, banded=True)
        sol = R.Hholtz(f, c, alpha=alpha, eig=(osol.solver.lam[0], osol.solver.bwd[0], osol.solver.fwd[0]))
    b = rng.uniform(-1, 1, (nx, ny))
    ref = osol.solve(b)
    e1 = rel(sol.solve(b), ref)
    e2 = rel(sol.solve(b * (1 + 1j)), ref * (1 + 1j))
    return e1, e2, sol, osol


def check_tensor_own_eig(lib, which, kx, ky, nx, ny, c=(1.0, 0.8), alpha=2.0, seed=13):
    """Hholtz / Poisson with the library's own set-up (even/odd block diagonalisation ->
    exactly checkerboard Q, P -> parity-split GEMMs); the oracle consumes the exported (lam, Q, P)."""
    rng = np.random.default_rng(seed)
    f, of = make_fields(lib, kx, nx, ky, ny)
    if which == "poisson":
        sol, osol = R.Poisson(f, c), O.Poisson(of, c, banded=True)
    else:
        sol, osol = R.Hholtz(f, c, alpha=alpha), O.Hholtz(of, c, alpha=alpha, banded=True)
    lam, q, p = sol.export_eig()
    m = nx - 2
    # the exported decomposition is exactly checkerboard and reproduces inv(Cx) Ax
    for k in range(m):
        assert not (np.any(q[0::2, k] != 0) and np.any(q[1::2, k] != 0)), "Q not checkerboard"
        assert not (np.any(p[k, 0::2] != 0) and np.any(p[k, 1::2] != 0)), "P not checkerboard"
    assert np.all(np.diff(lam) <= 0), "eigenvalues not sorted descending (utils.rs:80-94)"
    lam_ref = np.sort(osol.solver.lam[0])[::-1]
    assert np.abs(lam - lam_ref).max() <= 1e-6 * max(1.0, np.abs(lam_ref).max()), "eigenvalues differ from dgeev on the full matrix"
    osol.solver.lam[0], osol.solver.bwd[0], osol.solver.fwd[0] = lam, q, p
    b = rng.uniform(-1, 1, (nx, ny))
    ref = osol.solve(b)
    e1 = rel(sol.solve(b), ref)
    e2 = rel(sol.solve(b * (1 + 1j)), ref * (1 + 1j))
    return e1, e2


def check_tensor_fourier(lib, nx, ny, seed=5, tol=TOL):
    rng = np.random.default_rng(seed)
    f, of = make_fields(lib, "fourier_r2c", nx, "cheb_dirichlet", ny)
    b = rng.uniform(-1, 1, (nx // 2 + 1, ny)) + 1j * rng.uniform(-1, 1, (nx // 2 + 1, ny))
    e1 = rel(R.Hholtz(f, [0.1, 0.2]).solve(b), O.Hholtz(of, [0.1, 0.2], banded=True).solve(b))
    e2 = rel(R.Poisson(f, [1.0, 1.0]).solve(b), O.Poisson(of, [1.0, 1.0], banded=True).solve(b))
    assert e1 <= tol and e2 <= tol, (e1, e2)
    return e1, e2


def make_navier_pair(lib, periodic, nx, ny, ra, pr, dt, aspect=1.0, adiabatic=True, ics=True, own_eig=False):
    """Device Navier2D and oracle Navier2D with identical set-up data and deterministic ICs.
    own_eig: the device library does the pressure-Poisson eigen set-up itself (parity-split
    mode) and the oracle consumes the exported (lam, Q, P); else the oracle's go to the device."""
    if periodic:
        o = O.Navier2D.new_periodic(nx, ny, ra, pr, dt, aspect, banded=True)
        n = R.Navier2D.new_periodic(nx, ny, ra, pr, dt, aspect, lib=lib)
    elif own_eig:
        n = R.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, lib=lib)
        lam, q, p = n.export_eig()  # lam already carries the poisson.rs:80-83 shift
        o = O.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, banded=True, eig_data=(lam, q, p))
        o.solver[3].solver.lam[0] = lam.copy()
    else:
        o = O.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, banded=True)
        ts = o.solver[3].solver
        lam = ts.lam[0].copy()
        if abs(lam[0] + 1e-10) < 1e-10:
            lam = lam + 1e-10
        n = R.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, eig=(lam, ts.bwd[0], ts.fwd[0]), lib=lib)
    if ics:
        for x in (n, o):
            x.set_velocity(0.2, 1.0, 1.0)
            x.set_temperature(0.2, 1.0, 1.0)
    return n, o


def navier_field_errors(n, o):
    return {
        "temp": rel(n.temp.vhat, o.temp.vhat),
        "ux": rel(n.ux.vhat, o.ux.vhat),
        "uy": rel(n.uy.vhat, o.uy.vhat),
        "pres": rel(n.pres[0].vhat, o.pres[0].vhat) if np.abs(o.pres[0].vhat).max() > 0 else 0.0,
    }


def oracle_diag(o):
    return [o.eval_nu(), o.eval_nuvol(), o.eval_re(), o.div_norm(), o.eval_ekin()]


def check_navier_steps(lib, periodic, nx, ny, nsteps, ra=1e5, pr=1.0, dt=0.01, adiabatic=True, tol=1e-9, batch=1, own_eig=False):
    n, o = make_navier_pair(lib, periodic, nx, ny, ra, pr, dt, adiabatic=adiabatic, own_eig=own_eig)
    e0 = navier_field_errors(n, o)
    assert max(e0.values()) <= TOL, e0
    done = 0
    while done < nsteps:
        k = min(batch, nsteps - done)
        n.update(k)
        for _ in range(k):
            o.update()
        done += k
    n.sync()
    err = navier_field_errors(n, o)
    assert max(err.values()) <= tol, err
    dn, do = n.eval(), oracle_diag(o)
    derr = [abs(a - b) / max(abs(b), 1e-300) for a, b in zip(dn, do)]
    assert abs(n.time - o.time) < 1e-12
    return err, derr, dn, do


def check_staged_upload(lib, periodic, nx, ny, steps=3):
    """stage_state + commit_staged must be the same as assigning the four vhat arrays (bit for bit), also when the
    next upload is queued while the current step is still running."""
    import rustpde_b200 as R

    def make():
        n = (R.Navier2D.new_periodic(nx, ny, 1e5, 1.0, 0.01, 1.0, lib=lib) if periodic
             else R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, lib=lib))
        n.set_velocity(0.2, 1.0, 1.0)
        n.set_temperature(0.2, 1.0, 1.0)
        return n

    src = make()
    states = []
    for _ in range(steps):
        src.update(2)
        states.append([np.array(f.vhat) for f in (src.temp, src.ux, src.uy, src.pres[0])])
    a, b = make(), make()
    outs_a, outs_b = [], []
    for st in states:  # plain assignment
        a.temp.vhat, a.ux.vhat, a.uy.vhat = st[0], st[1], st[2]
        a.pres[0].vhat = st[3]
        a.update(1)
        outs_a.append([np.array(f.vhat) for f in (a.temp, a.ux, a.uy, a.pres[0])])
    b.stage_state(*states[0])
    for k in range(len(states)):  # double-buffered: upload k+1 is queued behind update k
        b.commit_staged()
        b.update(1)
        if k + 1 < len(states):
            b.stage_state(*states[k + 1])
        outs_b.append([np.array(f.vhat) for f in (b.temp, b.ux, b.uy, b.pres[0])])
    for oa, ob in zip(outs_a, outs_b):
        for x, y in zip(oa, ob):
            assert np.array_equal(x, y)
    return True
