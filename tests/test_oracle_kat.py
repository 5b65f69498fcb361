"""Pin the CPU oracle against every known-answer vector the reference's own
unit tests / doc-tests hold for the Navier2D hot path.

Tolerance: the reference asserts at absolute 1e-3 (funspace/src/utils.rs:82-94);
where the reference prints more digits we assert to the printed precision.
"""
import math

import numpy as np
import pytest

import oracle as O
from oracle import solver as S
from oracle.funspace import StencilChebyshev, StencilChebyshevBoundary, diffmat_chebyshev, cheb_pinv
from oracle.navier import conv_term


def close(a, b, tol=1e-3):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol, np.max(np.abs(a - b))


# ---- transforms: conventions of ndrustfft (third-party, un-vendored) ------
def test_chebyshev_forward_backward_doc():  # ortho.rs:210-220, 260-270
    ch = O.chebyshev(4)
    close(ch.forward(np.array([1.0, 2, 3, 4]), 0), [2.5, 1.33333333, 0.0, 0.16666667], 1e-8)
    close(ch.backward(np.array([1.0, 2, 3, 4]), 0), [-2.0, 2.5, -3.5, 10.0], 1e-12)


def test_dirichlet_forward_backward_doc():  # composite.rs:351-362, 398-409
    cd = O.cheb_dirichlet(5)
    close(cd.forward(np.array([1.0, 2, 3, 4, 5]), 0), [2.0, 0.70710678, 1.0], 1e-8)
    close(cd.backward(np.array([1.0, 2, 3]), 0), [0.0, 1.1716, -4.0, 6.8284, 0.0], 1e-4)


def test_r2c_doc():  # r2c.rs:130-147, 177-194
    fo = O.fourier_r2c(4)
    close(fo.forward(np.array([1.0, 2, 3, 4]), 0), [10.0, -2 + 2j, -2.0], 1e-12)
    close(fo.backward(np.array([10.0, -2 + 2j, -2.0]), 0), [1.0, 2, 3, 4], 1e-12)


def test_r2c_second_derivative_doc():  # funspace/src/lib.rs:46-72
    fo = O.fourier_r2c(8)
    x = fo.coords()
    vhat = fo.forward(np.sin(2.0 * x), 0)
    dv = fo.backward(fo.differentiate(vhat, 2, 0), 0)
    close(dv, -4.0 * np.sin(2.0 * x), 1e-5)


def test_r2c_differentiate_lane_doc():  # r2c.rs:74-87
    fo = O.fourier_r2c(5)
    close(fo.differentiate(fo.k.copy(), 1, 0), fo.k ** 2, 1e-12)


def test_cheb_differentiate_lane_doc():  # ortho.rs:90-105
    ch = O.chebyshev(4)
    close(ch.differentiate(np.array([1.0, 2, 3, 4]), 2, 0), [12.0, 96.0, 0.0, 0.0], 1e-12)
    close(ch.differentiate(np.array([1.0, 2, 3, 4]), 0, 0), [1.0, 2, 3, 4], 0)


def test_cheby_differentiate_2d():  # ortho.rs:628-664
    nx, ny = 6, 4
    data = np.arange(nx * ny, dtype=float).reshape(nx, ny)
    exp0 = np.array(
        [[140, 149, 158, 167], [160, 172, 184, 196], [272, 288, 304, 320], [128, 136, 144, 152], [200, 210, 220, 230], [0, 0, 0, 0]],
        dtype=float,
    )
    close(O.chebyshev(nx).differentiate(data, 1, 0), exp0, 1e-10)
    exp1 = np.array([[10, 8, 18, 0], [26, 24, 42, 0], [42, 40, 66, 0], [58, 56, 90, 0], [74, 72, 114, 0], [90, 88, 138, 0]], dtype=float)
    close(O.chebyshev(ny).differentiate(data, 1, 1), exp1, 1e-10)


def test_chebdirichlet_to_ortho():  # composite.rs:613-646
    nx, ny = 5, 4
    c = np.arange((nx - 2) * ny, dtype=float).reshape(nx - 2, ny)
    exp = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [8, 8, 8, 8], [-4, -5, -6, -7], [-8, -9, -10, -11]], dtype=float)
    close(O.cheb_dirichlet(nx).to_ortho(c, 0), exp, 1e-12)
    c = np.arange(nx * (ny - 2), dtype=float).reshape(nx, ny - 2)
    exp = np.array([[0, 1, 0, -1], [2, 3, -2, -3], [4, 5, -4, -5], [6, 7, -6, -7], [8, 9, -8, -9]], dtype=float)
    close(O.cheb_dirichlet(ny).to_ortho(c, 1), exp, 1e-12)


def test_chebdirichlet_differentiate():  # composite.rs:648-688 + doc 512-522
    nx, ny = 6, 4
    data = np.arange(nx * ny, dtype=float).reshape(nx, ny)
    exp0 = np.array(
        [
            [-1440.0, -1548.0, -1656.0, -1764.0],
            [-5568.0, -5904.0, -6240.0, -6576.0],
            [-2688.0, -2880.0, -3072.0, -3264.0],
            [-4960.0, -5240.0, -5520.0, -5800.0],
            [-1920.0, -2040.0, -2160.0, -2280.0],
            [-3360.0, -3528.0, -3696.0, -3864.0],
            [0.0, 0.0, 0.0, 0.0],
            [0.0, 0.0, 0.0, 0.0],
        ]
    )
    close(O.cheb_dirichlet(nx + 2).differentiate(data, 2, 0), exp0, 1e-9)
    exp1 = np.array(
        [
            [-56.0, -312.0, -96.0, -240.0, 0.0, 0.0],
            [-184.0, -792.0, -288.0, -560.0, 0.0, 0.0],
            [-312.0, -1272.0, -480.0, -880.0, 0.0, 0.0],
            [-440.0, -1752.0, -672.0, -1200.0, 0.0, 0.0],
            [-568.0, -2232.0, -864.0, -1520.0, 0.0, 0.0],
            [-696.0, -2712.0, -1056.0, -1840.0, 0.0, 0.0],
        ]
    )
    close(O.cheb_dirichlet(ny + 2).differentiate(data, 2, 1), exp1, 1e-9)
    close(O.cheb_dirichlet(5).differentiate(np.array([1.0, 2, 3]), 2, 0), [-88.0, -48.0, -144.0, 0.0, 0.0], 1e-10)


def test_stencils():  # composite_stencil.rs:416-454
    st = StencilChebyshev(5, "dirichlet")
    close(st.solve(np.array([2.0, 0.7071, -1.0, -0.7071, -1.0])), [2.0, 0.70710678, 1.0])
    close(st.multiply(np.array([2.0, 0.70710678, 1.0])), [2.0, 0.7071, -1.0, -0.7071, -1.0])
    z = np.array([2.0, 0.7071, -1.0, -0.7071, -1.0]) * (1 + 1j)
    close(st.solve(z), np.array([2.0, 0.70710678, 1.0]) * (1 + 1j))
    sb = StencilChebyshevBoundary(4, "dirichlet")
    close(sb.solve(np.array([1.0, 2, 3, 4])), [-1.0, 3.0], 1e-12)
    close(sb.multiply(np.array([1.0, 2.0])), [1.5, 0.5, 0.0, 0.0], 1e-12)


def test_pinv_is_pseudo_inverse_doc():  # ortho.rs:497-506
    ch = O.chebyshev(5)
    peye = ch.laplace_inv() @ ch.laplace()
    close(peye[2:, :], ch.laplace_inv_eye(), 1e-12)
    assert diffmat_chebyshev(6, 1).shape == (6, 6) and cheb_pinv(6, 1)[1, 0] == 1.0


def test_space2_to_ortho_matches_traits_doc():  # funspace/src/lib.rs composite example
    cd = O.cheb_dirichlet(8)
    comp = cd.forward(np.arange(1.0, 9.0), 0)
    ortho = cd.to_ortho(comp, 0)
    assert comp.shape == (6,) and ortho.shape == (8,)
    close(cd.from_ortho(ortho, 0), comp, 1e-12)


# ---- averages: src/field/average.rs:11-24, 37-50 ---------------------------
def test_average_doc():
    f = O.Field2(O.Space2(O.chebyshev(6), O.chebyshev(5)))
    f.v[:, :] = np.arange(5.0)[None, :]
    close(f.average_axis(0), [0.0, 1.0, 2.0, 3.0, 4.0], 1e-12)
    assert abs(f.average() - 2.0) < 1e-12


# ---- banded solvers ---------------------------------------------------------
def _test_matrix(nx):  # fdma.rs doc / fdma_tensor.rs:272-287
    m = np.zeros((nx, nx))
    for i in range(nx):
        j = float(i + 1)
        m[i, i] = 0.5 * j
        if i > 1:
            m[i, i - 2] = 10.0 * j
        if i < nx - 2:
            m[i, i + 2] = 1.5 * j
        if i < nx - 4:
            m[i, i + 4] = 2.5 * j
    return m


@pytest.mark.parametrize("nx", [6, 9, 16])
def test_fdma_residual(nx):  # fdma.rs:134-160, 262-321
    m = _test_matrix(nx)
    b = np.arange(nx, dtype=float)
    x = S.Fdma.from_matrix(m).solve(b, 0)
    close(m @ x, b, 1e-9)
    b2 = np.arange(nx * 3, dtype=float).reshape(nx, 3)
    close(m @ S.Fdma.from_matrix(m).solve(b2, 0), b2, 1e-8)
    bc = b * (1 + 2j)
    close(m @ S.Fdma.from_matrix(m).solve(bc, 0), bc, 1e-9)


def test_fdma_unswept_asserts():  # fdma.rs:167-170
    with pytest.raises(AssertionError):
        S.Fdma.from_matrix_raw(_test_matrix(6)).solve(np.arange(6.0), 0)


def test_matvec_fdma():  # matvec.rs:269-359
    nx = 8
    m = _test_matrix(nx)
    b = np.arange(nx, dtype=float)
    close(S.MatVecFdma(m).solve(b, 0), m @ b, 1e-12)
    b2 = np.arange(nx * 5, dtype=float).reshape(nx, 5)
    close(S.MatVecFdma(m).solve(b2, 0), m @ b2, 1e-12)
    b3 = np.arange(nx * 5, dtype=float).reshape(5, nx)
    close(S.MatVecFdma(m).solve(b3, 1), b3 @ m.T, 1e-12)


def test_eig_reconstruct():  # utils.rs:182-203
    t = np.tile(np.arange(1.0, 6.0), (5, 1))
    e, q, qi = S.eig(t)
    close(q @ np.diag(e) @ qi, t, 1e-9)
    assert np.all(np.diff(e) <= 1e-12)  # sorted largest -> smallest


def test_fdma_tensor_2d():  # fdma_tensor.rs:322-397 (residual of the Kronecker system)
    nx, ny = 6, 8
    a0, c0, a1, c1 = _test_matrix(nx), _test_matrix(nx)[::-1, ::-1].copy() + np.eye(nx), _test_matrix(ny), np.eye(ny)
    alpha = 0.3
    t = S.FdmaTensor([a0, a1], [c0, c1], [False, False], alpha)
    b = np.arange(nx * ny, dtype=float).reshape(nx, ny)
    g = t.solve(b)
    # [(A0 x C1) + (C0 x A1) + alpha (C0 x C1)] g = f
    lhs = a0 @ g @ c1.T + c0 @ g @ a1.T + alpha * c0 @ g @ c1.T
    close(lhs, b, 1e-7)


def test_hholtz_adi_1d_kat():  # hholtz_adi.rs:154-173 (pypde vector)
    # 1-D ADI == matvec then Fdma along the only axis
    f = O.Field2(O.Space2(O.cheb_dirichlet(7), O.cheb_dirichlet(7)))
    h = O.HholtzAdi(f, [1.0, 1.0])
    b = np.arange(1.0, 8.0)
    x = h.solver[0].solve(h.matvec[0].solve(b, 0), 0)
    close(x, [-0.08214845, -0.10466761, -0.06042153, 0.04809052, 0.04082296], 1e-8)


def test_hholtz_adi_2d_kat():  # hholtz_adi.rs:176-207
    f = O.Field2(O.Space2(O.cheb_dirichlet(7), O.cheb_dirichlet(7)))
    b = np.tile(np.arange(1.0, 8.0), (7, 1))
    y = np.array(
        [
            [-7.083e-03, -9.025e-03, -5.210e-03, 4.146e-03, 3.520e-03],
            [5.809e-04, 7.402e-04, 4.273e-04, -3.401e-04, -2.887e-04],
            [1.699e-04, 2.165e-04, 1.250e-04, -9.951e-05, -8.447e-05],
            [-1.007e-03, -1.283e-03, -7.406e-04, 5.895e-04, 5.004e-04],
            [-6.775e-04, -8.632e-04, -4.983e-04, 3.966e-04, 3.366e-04],
        ]
    )
    for banded in (False, True):
        close(O.HholtzAdi(f, [1.0, 1.0], banded=banded).solve(b), y, 1e-6)  # printed to 4 digits (truncated)


def _analytic(field, fx, fy):
    x, y = field.x
    return fx(x)[:, None] * fy(y)[None, :]


@pytest.mark.parametrize("banded", [False, True])
def test_hholtz_adi_analytic(banded):  # hholtz_adi.rs:210-269
    n = math.pi / 2.0
    alpha = 1e-5
    f = O.Field2(O.Space2(O.cheb_dirichlet(16), O.cheb_dirichlet(7)))
    f.v = _analytic(f, lambda x: np.cos(n * x), lambda y: np.cos(n * y))
    exp = f.v / (1.0 + alpha * n * n * 2.0)
    f.forward()
    f.vhat = O.HholtzAdi(f, [alpha, alpha], banded=banded).solve(f.to_ortho())
    f.backward()
    close(f.v, exp)
    f = O.Field2(O.Space2(O.fourier_r2c(16), O.cheb_dirichlet(7)))
    f.v = _analytic(f, np.cos, lambda y: np.cos(n * y))
    exp = f.v / (1.0 + alpha * n * n + alpha)
    f.forward()
    f.vhat = O.HholtzAdi(f, [alpha, alpha], banded=banded).solve(f.to_ortho())
    f.backward()
    close(f.v, exp)


@pytest.mark.parametrize("banded", [False, True])
def test_hholtz_analytic(banded):  # hholtz.rs:220-279
    n = math.pi / 2.0
    f = O.Field2(O.Space2(O.cheb_dirichlet(64), O.cheb_dirichlet(64)))
    f.v = _analytic(f, lambda x: np.cos(n * x), lambda y: np.cos(n * y))
    exp = f.v / (1.0 + n * n * 2.0)
    f.forward()
    f.vhat = O.Hholtz(f, [1.0, 1.0], banded=banded).solve(f.to_ortho())
    f.backward()
    close(f.v, exp)
    alpha = 1e-5
    f = O.Field2(O.Space2(O.fourier_r2c(16), O.cheb_dirichlet(7)))
    f.v = _analytic(f, np.cos, lambda y: np.cos(n * y))
    exp = f.v / (1.0 + alpha * n * n + alpha)
    f.forward()
    f.vhat = O.Hholtz(f, [alpha, alpha], banded=banded).solve(f.to_ortho())
    f.backward()
    close(f.v, exp)


def test_hholtz_new2_example():  # examples/hholtz_2d.rs:7-31
    n = math.pi / 2.0
    alpha = 1e-1
    f = O.Field2(O.Space2(O.cheb_dirichlet(64), O.cheb_dirichlet(64)))
    f.v = _analytic(f, lambda x: np.cos(n * x), lambda y: np.cos(n * y))
    exp = alpha / (1.0 + alpha * n * n * 2.0) * f.v
    f.forward()
    f.vhat = O.Hholtz.new2(f, [1.0, 1.0], 1.0 / alpha).solve(f.to_ortho())
    f.backward()
    close(f.v, exp)


def test_poisson_1d_kat():  # poisson.rs:188-205
    f = O.Field2(O.Space2(O.cheb_dirichlet(8), O.cheb_dirichlet(8)))
    p = O.Poisson(f, [1.0])  # N = 1 path: matvec + pre-swept Fdma of c*A
    x = p.solve(np.arange(1.0, 9.0))
    close(x, [0.1042, 0.0809, 0.0625, 0.0393, -0.0417, -0.0357], 1e-4)


POISSON_2D = np.array(
    [
        [0.01869736, 0.0244178, 0.01403203, -0.0202917, -0.0196697],
        [-0.0027890, -0.004035, -0.0059870, -0.0023490, -0.0046850],
        [-0.0023900, -0.007947, -0.0085570, -0.0189310, -0.0223680],
        [-0.0038940, -0.006622, -0.0096270, -0.0079020, -0.0120490],
        [0.00025400, -0.006752, -0.0082940, -0.0316230, -0.0361640],
        [-0.0001120, -0.004374, -0.0066430, -0.0216410, -0.0262570],
    ]
)


@pytest.mark.parametrize("banded", [False, True])
def test_poisson_2d_kat(banded):  # poisson.rs:208-274 (real and complex)
    f = O.Field2(O.Space2(O.cheb_dirichlet(8), O.cheb_dirichlet(7)))
    p = O.Poisson(f, [1.0, 1.0], banded=banded)
    b = np.tile(np.arange(1.0, 8.0), (8, 1))
    close(p.solve(b), POISSON_2D, 2e-6)
    close(p.solve(b * (1 + 1j)), POISSON_2D * (1 + 1j), 2e-6)


@pytest.mark.parametrize("banded", [False, True])
def test_poisson_analytic(banded):  # poisson.rs:277-339
    n = math.pi / 2.0
    f = O.Field2(O.Space2(O.cheb_dirichlet(8), O.cheb_dirichlet(7)))
    f.v = _analytic(f, lambda x: np.cos(n * x), lambda y: np.cos(n * y))
    exp = -1.0 / (n * n * 2.0) * f.v
    f.forward()
    f.vhat = O.Poisson(f, [1.0, 1.0], banded=banded).solve(f.to_ortho())
    f.backward()
    close(f.v, exp)
    f = O.Field2(O.Space2(O.fourier_r2c(16), O.cheb_dirichlet(7)))
    f.v = _analytic(f, lambda x: np.cos(2.0 * x), lambda y: np.cos(n * y))
    exp = -1.0 / (4.0 + n * n) * f.v
    f.forward()
    f.vhat = O.Poisson(f, [1.0, 1.0], banded=banded).solve(f.to_ortho())
    f.backward()
    close(f.v, exp)


def test_conv_term():  # conv_term.rs:65-119
    nx = ny = 12
    temp = O.Field2(O.Space2(O.cheb_dirichlet(nx), O.cheb_neumann(nx)))
    ux = O.Field2(O.Space2(O.cheb_dirichlet(nx), O.cheb_dirichlet(nx)))
    field = O.Field2(O.Space2(O.chebyshev(nx), O.chebyshev(nx)))
    x, y = field.x
    ax = ay = math.pi
    temp.v = np.sin(ax * x)[:, None] * np.cos(ay * y)[None, :]
    ux.v = np.sin(ax * x)[:, None] * np.sin(ay * y)[None, :]
    temp.forward()
    close(conv_term(temp, field, ux.v, [1, 0], None), ax * np.cos(ax * x)[:, None] * np.cos(ay * y)[None, :] * ux.v)
    close(conv_term(temp, field, ux.v, [0, 1], None), -ay * np.sin(ax * x)[:, None] * np.sin(ay * y)[None, :] * ux.v)


# ---- closed-form banded operators == literal dense construction (SURVEY 8a'') ----
@pytest.mark.parametrize("kind", ["chebyshev", "dirichlet", "neumann"])
@pytest.mark.parametrize("n", [8, 17])
def test_closed_form_operators(kind, n):
    base = O.chebyshev(n) if kind == "chebyshev" else O.CompositeChebyshev(n, kind)
    f = O.Field2(O.Space2(base, O.chebyshev(6)))
    mat_a, mat_b, precond = f.ingredients_for_hholtz(0)
    A, C = S.closed_form_a_c(base)
    m = n - 2
    close(S._dense_from_diags(C, m), mat_a, 2e-17)
    close(S._dense_from_diags(A, m), mat_b, 2e-17)
    mv_dense, mv_cf = S.MatVecFdma(precond), S.closed_form_precond(n)
    for k in ("low", "dia", "up1", "up2"):
        close(getattr(mv_dense, k), getattr(mv_cf, k), 2e-17)
    # the dense products really are banded (nothing outside -2,0,2,4)
    band = S._dense_from_diags({k: S.diag(mat_a, o) for k, o in (("low", -2), ("dia", 0), ("up1", 2), ("up2", 4))}, m)
    close(band, mat_a, 1e-18)


# ---- Navier2D: smoke (navier.rs:139-152) + physics sanity -----------------
def test_navier_smoke_and_sanity():
    nav = O.Navier2D.new(33, 33, 1e5, 1.0, 0.01, 1.0, True)
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    steps = O.integrate(nav, 0.2, None)
    assert steps == 20 and abs(nav.time - 0.2) < 1e-12
    assert nav.div_norm() < 1e-2 and not nav.exit()
    assert 0.5 < nav.eval_nu() < 5.0
    # banded setup path gives the same trajectory
    nb = O.Navier2D.new(33, 33, 1e5, 1.0, 0.01, 1.0, True, banded=True)
    nb.set_velocity(0.2, 1.0, 1.0)
    nb.set_temperature(0.2, 1.0, 1.0)
    O.integrate(nb, 0.2, None)
    close(nb.temp.vhat, nav.temp.vhat, 1e-12)
    close(nb.ux.vhat, nav.ux.vhat, 1e-12)


def test_navier_periodic_smoke():
    nav = O.Navier2D.new_periodic(32, 33, 1e5, 1.0, 0.01, 1.0)
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    O.integrate(nav, 0.1, None)
    assert nav.div_norm() < 1e-1  # IC is not exactly periodic (navier.rs:1044)
    assert nav.temp.vhat.dtype == np.complex128 and nav.temp.vhat.shape == (17, 31)


def test_bc_rbc_profile():  # navier.rs:314-332: T = +0.5 at y=-1, -0.5 at y=+1
    bc = O.Navier2D.bc_rbc(9, 9)
    close(bc.v[:, 0], 0.5 * np.ones(9), 1e-13)
    close(bc.v[:, -1], -0.5 * np.ones(9), 1e-13)
    close(bc.v, np.tile(-0.5 * bc.x[1], (9, 1)), 1e-13)
