"""Parity tests proper: the CUDA path behind the C ABI versus the oracle on a
real B200 (pytest -m gpu).  Tolerances are written in each test."""
import numpy as np
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded(gpu):
    assert not gpu.emulated and gpu.path.endswith("librustpde_b200.so")
    import subprocess, os
    maps = open("/proc/%d/maps" % os.getpid()).read()
    assert "librustpde_b200.so" in maps


@pytest.mark.parametrize(
    "kx,nx,ky,ny",
    [
        ("chebyshev", 17, "chebyshev", 33),
        ("cheb_dirichlet", 16, "cheb_dirichlet", 17),
        ("cheb_neumann", 33, "cheb_dirichlet", 20),
        ("cheb_dirichlet", 64, "cheb_neumann", 64),      # config 1 (non-pow2 periods, Bluestein)
        ("chebyshev", 7, "cheb_dirichlet", 9),
        ("fourier_r2c", 24, "chebyshev", 21),
        ("fourier_r2c", 512, "cheb_dirichlet", 513),      # config 3
        ("cheb_dirichlet", 1024, "cheb_dirichlet", 1025),  # config 2
        ("cheb_neumann", 300, "cheb_neumann", 257),
    ],
)
def test_field_ops(gpu, kx, nx, ky, ny):
    # <= 1e-10 relative per transform (north_star)
    pc.check_field_ops(gpu, kx, nx, ky, ny, tol=1e-10)


@pytest.mark.parametrize("kx,nx,ky,ny", [("cheb_dirichlet", 2048, "cheb_dirichlet", 2049), ("fourier_r2c", 2048, "cheb_dirichlet", 2049)])
def test_field_ops_full_size(gpu, kx, nx, ky, ny):
    # config-4 grid: every transform / gradient against the oracle, <= 1e-10 relative
    pc.check_field_ops(gpu, kx, nx, ky, ny, tol=1e-10, grads=((1, 0), (0, 1)))


def test_long_lanes_8193(gpu):
    # config-5 lane lengths (x: r2c 8192, y: DCT-I 8193) on a thin slab
    import oracle as O, rustpde_b200 as R
    rng = np.random.default_rng(3)
    f, of = pc.make_fields(gpu, "fourier_r2c", 8192, "cheb_dirichlet", 65)
    v = rng.uniform(-1, 1, (8192, 65))
    f.v, of.v = v, v.copy()
    f.forward(); of.forward()
    assert pc.rel(f.vhat, of.vhat) <= 1e-10
    f.backward(); of.backward()
    assert pc.rel(f.v, of.v) <= 1e-10
    f, of = pc.make_fields(gpu, "fourier_r2c", 64, "cheb_dirichlet", 8193)
    v = rng.uniform(-1, 1, (64, 8193))
    f.v, of.v = v, v.copy()
    f.forward(); of.forward()
    assert pc.rel(f.vhat, of.vhat) <= 1e-10
    f.backward(); of.backward()
    assert pc.rel(f.v, of.v) <= 1e-10
    assert pc.rel(f.gradient([1, 1], None), of.gradient([1, 1], None)) <= 1e-10


@pytest.mark.parametrize("nx,ny", [(7, 7), (33, 40), (129, 66), (512, 513), (2048, 2049)])
def test_hholtz_adi(gpu, nx, ny):
    pc.check_adi(gpu, nx, ny, tol=1e-10)


@pytest.mark.parametrize("which,kx,ky,nx,ny", [
    ("poisson", "cheb_neumann", "cheb_neumann", 16, 19),
    ("hholtz", "cheb_dirichlet", "cheb_dirichlet", 64, 64),
    ("poisson", "cheb_neumann", "cheb_neumann", 140, 34),
    ("poisson", "cheb_neumann", "cheb_neumann", 256, 257),
    ("hholtz", "cheb_dirichlet", "cheb_dirichlet", 512, 513),
])
def test_fast_diag_shared_eig(gpu, which, kx, ky, nx, ny):
    """Strict mode: same (lam, Q, P) on both sides.  Tolerance max(1e-10, floor(n)) where floor(n) is the
    difference between two summation orders of the *oracle's own* GEMMs (P is ill-conditioned: SURVEY section 7)."""
    e1, e2, sol, osol = pc.check_tensor_shared_eig(gpu, which, kx, ky, nx, ny)
    rng = np.random.default_rng(11)
    b = rng.uniform(-1, 1, (nx, ny))
    ref = osol.solve(b)
    # oracle-vs-oracle floor: same algorithm, GEMM accumulated in a different (pairwise, float64) order
    P, Q = osol.solver.fwd[0], osol.solver.bwd[0]
    osol.solver.fwd[0] = P[:, ::-1].copy()
    rhs_flip = None
    import oracle.solver as S
    class Flip:
        def __init__(self, mv): self.mv = mv
    alt = osol.solver.fwd[0]
    osol.solver.fwd[0] = P
    # emulate a different summation order by splitting the contraction in halves
    def solve_split():
        rhs = osol.matvec[0].solve(b, 0)
        rhs = osol.matvec[1].solve(rhs, 1)
        h = P.shape[1] // 2
        out = P[:, h:] @ rhs[h:] + P[:, :h] @ rhs[:h]
        l = (osol.solver.lam[0] + osol.solver.alpha)[:, None]
        f0, f1 = osol.solver.fdma
        S.fdma_solve_multi(f0.low[None] + f1.low[None] * l, f0.dia[None] + f1.dia[None] * l,
                           f0.up1[None] + f1.up1[None] * l, f0.up2[None] + f1.up2[None] * l, out)
        return Q[:, h:] @ out[h:] + Q[:, :h] @ out[:h]
    floor = pc.rel(solve_split(), ref)
    tol = max(1e-10, 20.0 * floor)
    assert e1 <= tol and e2 <= tol, (e1, e2, floor)


@pytest.mark.parametrize("nx,ny", [(16, 7), (32, 33), (512, 513)])
def test_fast_diag_fourier(gpu, nx, ny):
    pc.check_tensor_fourier(gpu, nx, ny, tol=1e-10)


def test_hholtz_example_1024(gpu):
    """Config 2: examples/hholtz_2d.rs scaled to 1024x1025, analytic solution (reference tolerance 1e-3)."""
    import math, rustpde_b200 as R
    nx, ny = 1024, 1025
    f = R.Field2(R.Space2(R.cheb_dirichlet(nx), R.cheb_dirichlet(ny)), lib=gpu)
    x, y = f.x
    n = math.pi / 2.0
    alpha = 1e-1
    v = np.cos(n * x)[:, None] * np.cos(n * y)[None, :]
    f.v = v
    f.forward()
    h = R.Hholtz.new2(f, [1.0, 1.0], 1.0 / alpha)
    f.vhat = h.solve(f.to_ortho())
    f.backward()
    assert np.abs(f.v - alpha / (1.0 + alpha * n * n * 2.0) * v).max() < 1e-3


@pytest.mark.parametrize("nx,ny,adiabatic,steps", [(16, 17, True, 6), (33, 33, False, 6), (64, 64, True, 100), (128, 129, True, 20)])
def test_navier_confined(gpu, nx, ny, adiabatic, steps):
    # fields <= 1e-9 relative after `steps` steps, diagnostics <= 1e-9 (set-up data shared)
    err, derr, dn, do = pc.check_navier_steps(gpu, False, nx, ny, steps, adiabatic=adiabatic, tol=1e-9, batch=5)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("nx,ny,adiabatic,steps,own_eig", [
    (32, 33, True, 6, False),
    (64, 65, False, 50, True),
    (1024, 1025, True, 3, True),     # config-2 grid
    (2048, 2049, True, 2, True),     # config-4 grid: the benchmark configuration itself
])
def test_navier_confined_specialised_kernels(gpu, nx, ny, adiabatic, steps, own_eig):
    """The hand-specialised x/y pass kernels (Bluestein DCT along x, pow2 DCT along y) and, with own_eig,
    the parity-split GEMMs on the library's own eigen set-up (exported to the oracle).
    Fields <= 1e-9 relative, Nu / Nuvol / Re / |div| / Ekin <= 1e-9 relative."""
    import rustpde_b200 as R
    assert R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, adiabatic, lib=gpu).kernel_path() == (True, True)
    ra, dt = (1e5, 0.01) if nx < 1000 else (1e9, 1e-4)
    err, derr, dn, do = pc.check_navier_steps(gpu, False, nx, ny, steps, ra=ra, dt=dt, adiabatic=adiabatic, tol=1e-9, batch=2, own_eig=own_eig)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("which,kx,ky,nx,ny", [
    ("poisson", "cheb_neumann", "cheb_neumann", 16, 19),
    ("poisson", "cheb_dirichlet", "cheb_dirichlet", 141, 34),
    ("hholtz", "cheb_dirichlet", "cheb_dirichlet", 512, 513),
    ("poisson", "cheb_neumann", "cheb_neumann", 1024, 1025),
])
def test_fast_diag_parity_split(gpu, which, kx, ky, nx, ny):
    """Parity-split mode: the library's own set-up (even/odd block diagonalisation, exactly checkerboard
    Q and P) and two half-size DMMA GEMM pairs; the oracle consumes the exported (lam, Q, P).
    Tolerance as in test_fast_diag_shared_eig: max(1e-10, 20 x summation-order floor of the oracle)."""
    import oracle.solver as S
    e1, e2 = pc.check_tensor_own_eig(gpu, which, kx, ky, nx, ny)
    tol = 1e-10 if nx <= 141 else 1e-8
    assert e1 <= tol and e2 <= tol, (e1, e2)


@pytest.mark.parametrize("nx,ny,steps", [(32, 33, 6), (64, 65, 50), (512, 513, 20), (2048, 2049, 2),
                                         (128, 129, 10),     # 2-lane tiles on both axes (small instantiation)
                                         (8192, 65, 3),      # config-5 lane length along x (r2c 8192, 2-lane tile)
                                         (64, 8193, 3)])     # config-5 lane length along y (DCT-I 8193, 2-lane tile)
def test_navier_periodic_specialised_kernels(gpu, nx, ny, steps):
    """Specialised periodic kernels (fast_p.cu): config-3 grid 512x513 and the 2048x2049 grid.
    Fields <= 1e-9 relative, diagnostics <= 1e-9 relative."""
    import rustpde_b200 as R
    assert R.Navier2D.new_periodic(nx, ny, 1e6, 1.0, 2e-3, 1.0, lib=gpu).kernel_path()[0]
    ra, dt = (1e6, 2e-3) if nx < 1000 else (1e9, 1e-4)
    err, derr, dn, do = pc.check_navier_steps(gpu, True, nx, ny, steps, ra=ra, dt=dt, tol=1e-9, batch=5)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("pw", ["0", "1"])
@pytest.mark.parametrize("nx,ny,steps", [(32, 33, 6), (64, 65, 20), (512, 513, 10), (2048, 2049, 2), (64, 8193, 3), (8192, 65, 3)])
def test_navier_periodic_row_sweeps_and_tile_kernels(gpu, monkeypatch, nx, ny, steps, pw):
    """The per-mode Helmholtz / Poisson passes and the projection + pressure update forced onto the warp-serial row sweeps
    (fast_pw.cu, RUSTPDE_B200_PW=1) and onto the tile kernels (=0) at every size; by default the row count picks one
    (rows >= 448 / 1280 / 1536)."""
    monkeypatch.setenv("RUSTPDE_B200_PW", pw)
    ra, dt = (1e6, 2e-3) if nx < 1000 else (1e9, 1e-4)
    err, derr, dn, do = pc.check_navier_steps(gpu, True, nx, ny, steps, ra=ra, dt=dt, tol=1e-9, batch=5)
    assert max(derr) < 1e-9, (pw, derr, dn, do)


@pytest.mark.parametrize("nx,ny,steps", [(16, 17, 6), (24, 20, 6), (128, 129, 50), (512, 513, 5)])
def test_navier_periodic(gpu, nx, ny, steps):
    err, derr, dn, do = pc.check_navier_steps(gpu, True, nx, ny, steps, ra=1e6, dt=2e-3, tol=1e-9, batch=5)
    assert max(derr) < 1e-9, (derr, dn, do)


def test_navier_1000_steps_nu_ke(gpu):
    """north_star end-to-end: Nusselt number and kinetic energy within 1e-8 after 1000 steps
    (config 1 grid 64x64, Ra=1e5, dt=0.02, smooth ICs, set-up data shared)."""
    n, o = pc.make_navier_pair(gpu, False, 64, 64, 1e5, 1.0, 0.02)
    n.update(1000)
    for _ in range(1000):
        o.update()
    dn, do = n.eval(), pc.oracle_diag(o)
    for name, a, b in zip(("Nu", "Nuvol", "Re", "div", "ekin"), dn, do):
        if name == "div":
            continue
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (name, a, b)
    assert abs(n.time - o.time) < 1e-9


@pytest.mark.parametrize("name,specialised", [("confined128", (True, True)), ("periodic512", (True, False))])
def test_navier_1000_steps_golden_specialised_kernels(gpu, name, specialised):
    """north_star end-to-end on the kernels the benchmark times: 1000 steps on grids served by the specialised x/y
    kernels (confined 128x129 incl. the parity-split DMMA GEMMs; periodic 512x513 = BASELINE config 3, Ra=1e7,
    dt=2e-3) against the committed oracle fixtures: Nu, Nuvol, Re and kinetic energy within 1e-8 at steps 250 ... 1000,
    final spectral coefficients within 1e-8 of the field maximum."""
    n, worst = pc.check_navier_golden(gpu, name, tol_obs=1e-8, tol_field=1e-8)
    assert n.kernel_path() == specialised, n.kernel_path()
    print(name, worst)


def test_navier_confined_2048_20_steps(gpu):
    """The benchmark configuration itself (2048x2049, Ra=1e9, dt=1e-4), 20 steps against the oracle fed with the
    library's exported (lam, Q, P): fields <= 1e-9 (max-norm) and <= 1e-6 banded, Nu / Nuvol / Re / |div| / Ekin <= 1e-9."""
    err, derr, dn, do = pc.check_navier_steps(gpu, False, 2048, 2049, 20, ra=1e9, dt=1e-4, tol=1e-9, batch=20, own_eig=True)
    assert max(derr) < 1e-9, (derr, dn, do)


def test_navier_no_dealias_specialised_kernels(gpu):
    """pub dealias = false (navier.rs:188): the 2/3 cuts fused into yk_conv / xk_forward / pk_r2c are switched off."""
    import oracle as O, rustpde_b200 as R
    for periodic, nx, ny in ((False, 64, 65), (True, 64, 65)):
        n, o = pc.make_navier_pair(gpu, periodic, nx, ny, 1e5, 1.0, 0.01, ics=False, own_eig=not periodic)
        n.dealias = False
        o.dealias = False
        for x in (n, o):
            x.set_velocity(0.2, 1.0, 1.0)
            x.set_temperature(0.2, 1.0, 1.0)
        assert n.kernel_path()[0]
        n.update(10)
        for _ in range(10):
            o.update()
        err = pc.navier_field_errors(n, o)
        assert max(err.values()) <= 1e-9, err
        # and the cut really matters at this resolution
        n2, _ = pc.make_navier_pair(gpu, periodic, nx, ny, 1e5, 1.0, 0.01, own_eig=not periodic)
        n2.update(10)
        assert pc.rel(n2.temp.vhat, n.temp.vhat) > 1e-12


@pytest.mark.parametrize("nx,ny", [(1024, 1025), (2048, 2049)])
def test_standalone_solvers_specialised_kernels(gpu, nx, ny):
    """Config 2 and the config-4 grid: stand-alone HholtzAdi / Hholtz / Poisson::solve on the specialised kernels
    (xk_adi + yk_adi; b2x + parity-split DMMA GEMM + yk_mode + GEMM), real data.  HholtzAdi <= 1e-10; fast
    diagonalisation with the oracle consuming the exported (lam, Q, P): <= max(1e-10, 20 x GEMM summation-order floor)."""
    import rustpde_b200 as R, oracle as O
    f = R.Field2(R.Space2(R.cheb_dirichlet(nx), R.cheb_dirichlet(ny)), lib=gpu)
    assert R.HholtzAdi(f, [0.3, 0.7]).path_info()["specialised"]
    pc.check_adi(gpu, nx, ny, tol=1e-10)
    if nx > 1024:
        return  # the oracle's dense 2046^2 set-up is exercised through Navier2D at this size
    e1, e2 = pc.check_tensor_own_eig(gpu, "hholtz", "cheb_dirichlet", "cheb_dirichlet", nx, ny)
    assert e1 <= 1e-8 and e2 <= 1e-8, (e1, e2)


@pytest.mark.parametrize("periodic,nx,ny", [(False, 128, 129), (True, 128, 129)])
def test_snapshot_roundtrip_gpu(gpu, tmp_path, periodic, nx, ny):
    """write() / read() (navier.rs:956-1014) on the device: dataset layout, bit-exact restart, broadcast on another grid."""
    from test_snapshot import check_snapshot_roundtrip
    assert check_snapshot_roundtrip(gpu, periodic, nx, ny, tmp_path)


def test_host_api_additions_gpu(gpu):
    """fetch_state / div_async / vhat row slabs / device averages / profile() == update() on the real streams."""
    assert pc.check_host_api_additions(gpu, False, 128, 129)
    assert pc.check_host_api_additions(gpu, True, 128, 129)


@pytest.mark.parametrize("periodic,nx,ny", [(False, 128, 129), (True, 128, 129)])
def test_solid_masks_gpu(gpu, periodic, nx, ny):
    """Volume penalisation (navier.rs:552-608) fused into the product kernel: fields <= 1e-9 against the oracle."""
    from test_solid_stats import check_solid
    check_solid(gpu, periodic, nx, ny, steps=10)


def test_statistics_gpu(gpu, tmp_path):
    from test_solid_stats import check_statistics
    assert check_statistics(gpu, False, 128, 129, tmp_path)


def test_navier_confined_tile_only_schedule(gpu, monkeypatch):
    """RUSTPDE_B200_XW=0: the round-1 schedule with the ADI-x sweeps inside the tile kernel (13 launches); the default
    (warp-serial ADI-x sweeps of fast_xw.cu, 14 launches) is what every other confined test runs; RUSTPDE_B200_XW=2 runs
    the divergence and the projection as warp-serial sweeps too (xw_div, xw_project: UL order of the from_ortho solve)."""
    import rustpde_b200 as R
    for xw, launches in (("0", 13), ("1", 14), ("2", 14)):
        monkeypatch.setenv("RUSTPDE_B200_XW", xw)
        for nx, ny, steps in ((64, 65, 10), (530, 129, 3)):
            n = R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, lib=gpu)
            n.set_velocity(0.2, 1.0, 1.0)
            n.set_temperature(0.2, 1.0, 1.0)
            n.update(1)
            assert n.launches_per_step() == launches
            err, derr, dn, do = pc.check_navier_steps(gpu, False, nx, ny, steps, tol=1e-9, batch=2, own_eig=True)
            assert max(derr) < 1e-9, (xw, derr, dn, do)


def test_navier_confined_column_scan_kernels(gpu, monkeypatch):
    """RUSTPDE_B200_XS=1: the x-direction sweeps as streaming column scans (fast_xs.cu) -- same results as the tile
    kernels to rounding, fields <= 1e-9 against the oracle."""
    import rustpde_b200 as R
    monkeypatch.setenv("RUSTPDE_B200_XS", "1")
    for nx, ny, steps in ((64, 65, 20), (530, 129, 3), (2048, 2049, 2)):
        ra, dt = (1e5, 0.01) if nx < 1000 else (1e9, 1e-4)
        n = R.Navier2D.new(nx, ny, ra, 1.0, dt, 1.0, True, lib=gpu)
        n.set_velocity(0.2, 1.0, 1.0)
        n.set_temperature(0.2, 1.0, 1.0)
        n.update(1)
        assert n.launches_per_step() == 15  # 13 + the split forward DCT + d/dx pres
        err, derr, dn, do = pc.check_navier_steps(gpu, False, nx, ny, steps, ra=ra, dt=dt, tol=1e-9, batch=2, own_eig=True)
        assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("periodic,nx,ny,steps", [(False, 128, 129, 5), (True, 128, 129, 5), (False, 512, 513, 2)])
def test_adjoint_gpu(gpu, periodic, nx, ny, steps):
    """Navier2DAdjoint (navier_adjoint.rs:128-1068): adjoint and residual fields <= 1e-9 relative against the oracle fed with
    the library's exported eigen set-up data; three Hholtz smoothers = six parity-split DMMA GEMMs per confined step."""
    from test_adjoint import check_adjoint
    print(check_adjoint(gpu, periodic, nx, ny, steps=steps, tol=1e-9))


def test_graph_and_eager_agree(gpu):
    import rustpde_b200 as R
    outs = []
    for graph in (True, False):
        n = R.Navier2D.new_periodic(64, 65, 1e5, 1.0, 0.01, 1.0, lib=gpu)
        n.set_graph(graph)
        n.set_velocity(0.2, 1.0, 1.0)
        n.set_temperature(0.2, 1.0, 1.0)
        n.update(7)
        outs.append(n.temp.vhat)
    assert np.array_equal(outs[0], outs[1])


def test_full_size_properties(gpu):
    """Config-4 grid (2048x2049), size-independent properties: linearity of the transforms,
    forward(backward(c)) == c on composite coefficients, and to_ortho/from_ortho round trip."""
    import rustpde_b200 as R
    rng = np.random.default_rng(5)
    f = R.Field2(R.Space2(R.cheb_dirichlet(2048), R.cheb_dirichlet(2049)), lib=gpu)
    c1 = rng.uniform(-1, 1, f.shape_spectral)
    c2 = rng.uniform(-1, 1, f.shape_spectral)
    f.vhat = c1; f.backward(); v1 = f.v
    f.vhat = c2; f.backward(); v2 = f.v
    f.vhat = 2.0 * c1 - 3.0 * c2; f.backward(); v3 = f.v
    assert pc.rel(v3, 2.0 * v1 - 3.0 * v2) < 1e-11
    f.forward()
    assert pc.rel(f.vhat, 2.0 * c1 - 3.0 * c2) < 1e-9
    o = f.to_ortho()
    f.from_ortho(o)
    assert pc.rel(f.vhat, 2.0 * c1 - 3.0 * c2) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("periodic,nx,ny", [(False, 64, 65), (True, 64, 65), (False, 1024, 1025)])
def test_staged_state_upload(gpu, periodic, nx, ny):
    """rp_navier_stage_state / commit_staged (copy stream, overlaps update()) == plain vhat uploads, bit for bit."""
    assert pc.check_staged_upload(gpu, periodic, nx, ny)


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny,steps", [(33, 33, 6), (129, 129, 6), (1025, 129, 3), (2049, 65, 2)])
def test_navier_confined_pow2_period_x(gpu, nx, ny, steps):
    """nx = 2^k + 1: power-of-two DCT period along x inside the x kernels (fast_x.cu dct_x), up to the largest tile."""
    import rustpde_b200 as R
    assert R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, lib=gpu).kernel_path() == (True, True)
    err, derr, dn, do = pc.check_navier_steps(gpu, False, nx, ny, steps, tol=1e-9, batch=2, own_eig=True)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny", [(530, 129), (1030, 33), (300, 257), (200, 65)])
def test_navier_confined_partial_lanes(gpu, nx, ny):
    """x lanes much shorter than the instantiated Bluestein length (2048 / 4096 rows): chunk-major coefficient tables
    with fewer rows of slots than the kernels' chunk bound, partly empty staged strips.  Diagnostics <= 1e-9 relative."""
    import rustpde_b200 as R
    assert R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, lib=gpu).kernel_path()[0]
    err, derr, dn, do = pc.check_navier_steps(gpu, False, nx, ny, 2, tol=1e-9, batch=2)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.gpu
def test_navier_periodic_256(gpu):
    """r2c length 256 / y lanes of 257 points (instantiations added last): specialised kernels, diagnostics <= 1e-9."""
    import rustpde_b200 as R
    assert R.Navier2D.new_periodic(256, 257, 1e5, 1.0, 0.01, 1.0, lib=gpu).kernel_path()[0]
    err, derr, dn, do = pc.check_navier_steps(gpu, True, 256, 257, 2, tol=1e-9, batch=2)
    assert max(derr) < 1e-9, (derr, dn, do)
