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
BAND_TOL = 1e-6  # banded metric (band_rel)

RB = {"chebyshev": R.chebyshev, "cheb_dirichlet": R.cheb_dirichlet, "cheb_neumann": R.cheb_neumann, "fourier_r2c": R.fourier_r2c}
OB = {"chebyshev": O.chebyshev, "cheb_dirichlet": O.cheb_dirichlet, "cheb_neumann": O.cheb_neumann, "fourier_r2c": O.fourier_r2c}


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def band_rel(a, b, nb=4, floor=1e-8):
    """Banded error metric: the array is cut into nb x nb blocks of mode bands (spectral coefficients span 15 decades,
    so a global max-norm cannot see the high modes); per band max|a-b| / max(max|b| of the band, floor) with
    floor = `floor` (default 1e-8) * global max|b| (rounding noise of a transform is relative to the LARGEST coefficient that went
    through it, so bands far below the floor are compared against the floor, not against themselves).  Returns the
    worst band; tolerance 1e-6 then means: every band down to 1e-8 of the peak is right to 6 digits, and nothing
    anywhere is off by more than 1e-14 of the peak."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    gmax = max(np.abs(b).max(), 1e-300)
    worst = 0.0
    r_edges = np.linspace(0, a.shape[0], nb + 1).astype(int)
    c_edges = np.linspace(0, a.shape[1], nb + 1).astype(int) if a.ndim > 1 else np.array([0, 1])
    for i in range(nb):
        for j in range(len(c_edges) - 1):
            sl = (slice(r_edges[i], r_edges[i + 1]), slice(c_edges[j], c_edges[j + 1])) if a.ndim > 1 else (slice(r_edges[i], r_edges[i + 1]),)
            bb = b[sl]
            if bb.size == 0:
                continue
            ref = max(np.abs(bb).max(), floor * gmax)
            worst = max(worst, float(np.abs(a[sl] - bb).max() / ref))
    return worst


def oracle_lam_unshifted(ts):
    """Eigenvalues of the oracle's FdmaTensor before the Poisson singularity shift (poisson.rs:80-83): the C ABI takes
    and returns unshifted eigenvalues (rustpde_b200.h, rp_solver_export_eig)."""
    lam = ts.lam[0].copy()
    if abs(lam[0] + 1e-10) < 1e-10:
        lam = lam + 1e-10
    return lam


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
        # the device takes the pre-shift eigenvalues and applies poisson.rs:80-83 itself
        sol = R.Poisson(f, c, eig=(oracle_lam_unshifted(osol.solver), osol.solver.bwd[0], osol.solver.fwd[0]))
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
    lam_used = lam - 1e-10 if (which == "poisson" and abs(lam[0]) < 1e-10) else lam  # poisson.rs:80-83
    osol.solver.lam[0], osol.solver.bwd[0], osol.solver.fwd[0] = lam_used, q, p
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
        # exported eigenvalues are unshifted; the oracle's Poisson applies poisson.rs:80-83 itself, like the device
        o = O.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, banded=True, eig_data=n.export_eig())
    else:
        o = O.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, banded=True)
        ts = o.solver[3].solver
        n = R.Navier2D.new(nx, ny, ra, pr, dt, aspect, adiabatic, eig=(oracle_lam_unshifted(ts), ts.bwd[0], ts.fwd[0]), lib=lib)
    if ics:
        for x in (n, o):
            x.set_velocity(0.2, 1.0, 1.0)
            x.set_temperature(0.2, 1.0, 1.0)
    return n, o


def navier_field_errors(n, o, metric=None):
    metric = metric or rel
    return {
        "temp": metric(n.temp.vhat, o.temp.vhat),
        "ux": metric(n.ux.vhat, o.ux.vhat),
        "uy": metric(n.uy.vhat, o.uy.vhat),
        "pres": metric(n.pres[0].vhat, o.pres[0].vhat) if np.abs(o.pres[0].vhat).max() > 0 else 0.0,
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
    berr = navier_field_errors(n, o, band_rel)
    assert max(berr.values()) <= max(BAND_TOL, 1e3 * tol), berr
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


def check_host_api_additions(lib, periodic, nx, ny):
    """Round-2 ABI additions against their blocking counterparts: profile() after an outside pressure write ==
    update(), eig export/import round trip, fetch_state == vhat, div_async/poll == div_norm,
    vhat row slabs, device-side average_axis versus the oracle."""
    import rustpde_b200 as R

    def make(eig=None):
        n = (R.Navier2D.new_periodic(nx, ny, 1e5, 1.0, 0.01, 1.0, lib=lib) if periodic
             else R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, eig=eig, lib=lib))
        n.set_velocity(0.2, 1.0, 1.0)
        n.set_temperature(0.2, 1.0, 1.0)
        return n

    a = make()
    a.update(3)
    state = [np.array(f.vhat) for f in (a.temp, a.ux, a.uy, a.pres[0])]
    # 1. profile(k) must advance exactly like update(k), also right after the pressure was written from outside
    b, c = make(), make()
    for n in (b, c):
        n.temp.vhat, n.ux.vhat, n.uy.vhat = state[0], state[1], state[2]
        n.pres[0].vhat = state[3]
    b.update(2)
    c.profile(2)
    for fb, fc in ((b.temp, c.temp), (b.ux, c.ux), (b.uy, c.uy), (b.pres[0], c.pres[0])):
        assert np.array_equal(fb.vhat, fc.vhat)
    assert abs(b.time - c.time) < 1e-14
    # 2. create_with_eig(export_eig()) reproduces the solver bit for bit (no double shift)
    if not periodic:
        d = make(eig=a.export_eig())
        d.update(3)
        for fa, fd in ((a.temp, d.temp), (a.ux, d.ux), (a.uy, d.uy), (a.pres[0], d.pres[0])):
            assert np.array_equal(fa.vhat, fd.vhat)
        lam = a.export_eig()[0]
        assert np.array_equal(lam, d.export_eig()[0])
    # 3. fetch_state (async, snapshot semantics): queued before further updates, must hold the state at the call
    bufs = [np.zeros_like(s) for s in state]
    a.fetch_state(*bufs)
    a.update(2)
    a.fetch_wait()
    for s, g in zip(state, bufs):
        assert np.array_equal(s, g)
    # 4. div_async / div_poll
    dn = b.div_norm()
    b.div_async()
    got = b.div_poll(wait=True)
    assert got == dn, (got, dn)
    assert not b.exit_async()
    # 5. row slabs of vhat
    full = np.array(b.ux.vhat)
    r0, nr = 1, min(5, full.shape[0] - 1)
    assert np.array_equal(b.ux.vhat_rows(r0, nr), full[r0:r0 + nr])
    b.ux.set_vhat_rows(r0, 2.0 * full[r0:r0 + nr])
    full[r0:r0 + nr] *= 2.0
    assert np.array_equal(np.array(b.ux.vhat), full)
    # 6. device-side average_axis(0) and average() versus the oracle (average.rs:25-57)
    f, of = make_fields(lib, "fourier_r2c" if periodic else "cheb_dirichlet", nx, "cheb_dirichlet", ny)
    v = smooth_field(of.x[0][:nx], of.x[1])
    f.v, of.v = v, v.copy()
    assert rel(f.average_axis(0), of.average_axis(0)) <= 1e-13
    assert abs(f.average() - of.average()) <= 1e-13 * max(1.0, abs(of.average()))
    return True


def check_navier_golden(lib, name, tol_obs=1e-8, tol_field=1e-8):
    """1000 steps against the committed oracle fixture tests/golden/navier_<name>_1000.npz (generated by
    tests/golden/make_navier_golden.py): Nu / Nuvol / Re / Ekin at steps 250, 500, 750, 1000 within tol_obs
    (north_star: 1e-8 after 1000 steps), final coefficients (low-mode block and strided sample of temp, ux, uy, pres)
    within tol_field of the field maximum.  Confined: the fixture's eigen set-up data are fed to the device."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "navier_%s_1000.npz" % name))
    nx, ny, ra, pr, dt, aspect, adiabatic = g["params"]
    nx, ny = int(nx), int(ny)
    if "eig_lam" in g.files:
        n = R.Navier2D.new(nx, ny, ra, pr, dt, aspect, bool(adiabatic), eig=(g["eig_lam"], g["eig_q"], g["eig_p"]), lib=lib)
    else:
        n = R.Navier2D.new_periodic(nx, ny, ra, pr, dt, aspect, lib=lib)
    n.set_velocity(0.2, 1.0, 1.0)
    n.set_temperature(0.2, 1.0, 1.0)
    done, worst = 0, {}
    for cp, ref in zip(g["checkpoints"], g["observables"]):
        n.update(int(cp) - done)
        done = int(cp)
        got = n.eval()
        for nm, a, b in zip(("Nu", "Nuvol", "Re", "div", "ekin"), got, ref):
            if nm == "div":
                continue  # |div| ~ 1e-5 is itself a difference of O(1) terms; it is covered by the field comparison
            e = abs(a - b) / max(1.0, abs(b))
            worst[nm] = max(worst.get(nm, 0.0), e)
            assert e <= tol_obs, (name, int(cp), nm, a, b)
    assert abs(n.time - float(g["time"])) < 1e-9
    for fname, f in (("temp", n.temp), ("ux", n.ux), ("uy", n.uy), ("pres", n.pres[0])):
        a = f.vhat
        scale = np.abs(a).max()
        for key, got in ((fname + "_low", a[:24, :24]), (fname + "_strided", a[::7, ::5])):
            e = float(np.abs(got - g[key]).max() / scale)
            worst[key] = e
            assert e <= tol_field, (name, key, e)
    return n, worst
