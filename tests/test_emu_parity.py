"""CPU-side check of the CUDA kernels' arithmetic: the same kernel sources,
compiled against the cuemu emulator, versus the numpy oracle.  (The parity
tests proper are test_gpu_parity.py; these keep index arithmetic, layouts and
numerics honest on a box without a GPU.)"""
import numpy as np
import pytest

import parity_cases as pc


@pytest.mark.parametrize(
    "kx,nx,ky,ny",
    [
        ("chebyshev", 17, "chebyshev", 33),          # pow2 DCT lengths (N = 16, 32)
        ("cheb_dirichlet", 16, "cheb_dirichlet", 17),  # Bluestein x (N = 15), pow2 y
        ("cheb_neumann", 33, "cheb_dirichlet", 20),    # pow2 x, Bluestein y (N = 19)
        ("cheb_dirichlet", 64, "cheb_neumann", 64),    # config-1 shape: N = 63 both
        ("fourier_r2c", 16, "cheb_dirichlet", 17),
        ("fourier_r2c", 24, "chebyshev", 21),          # Bluestein r2c (n = 24)
        ("fourier_r2c", 64, "cheb_neumann", 65),
        ("chebyshev", 7, "cheb_dirichlet", 9),         # tiny, odd/ragged pairs
    ],
)
def test_field_ops(emu, kx, nx, ky, ny):
    pc.check_field_ops(emu, kx, nx, ky, ny)


@pytest.mark.parametrize("nx,ny", [(7, 7), (16, 7), (33, 40), (129, 66)])
def test_hholtz_adi(emu, nx, ny):
    pc.check_adi(emu, nx, ny)


@pytest.mark.parametrize("nx,ny", [(24, 33), (40, 65), (18, 33), (130, 33)])
def test_standalone_solvers_on_specialised_kernels(emu, nx, ny):
    """Stand-alone HholtzAdi / Hholtz / Poisson::solve (real data) run on the specialised kernels when the shape
    qualifies (xk_adi + yk_adi; b2x + DMMA GEMM + yk_mode + DMMA GEMM) and agree with the lane programs."""
    import rustpde_b200 as R
    f = R.Field2(R.Space2(R.cheb_dirichlet(nx), R.cheb_dirichlet(ny)), lib=emu)
    assert R.HholtzAdi(f, [0.3, 0.7]).path_info()["specialised"]
    assert R.Hholtz(f, [1.0, 0.8], 2.0).path_info() == {"specialised": True, "split_gemm": True, "launches": 4}
    pc.check_adi(emu, nx, ny)
    for which, kx in (("hholtz", "cheb_dirichlet"), ("poisson", "cheb_neumann")):
        e1, e2 = pc.check_tensor_own_eig(emu, which, kx, kx, nx, ny)
        assert e1 <= pc.TOL and e2 <= pc.TOL, (which, e1, e2)
        e1, e2, sol, _ = pc.check_tensor_shared_eig(emu, which, kx, kx, nx, ny)
        assert e1 <= pc.TOL and e2 <= pc.TOL, (which, e1, e2)
    # resident repeats leave the result unchanged
    rng = np.random.default_rng(3)
    h = R.Hholtz(f, [1.0, 1.0], 10.0)
    b = rng.uniform(-1, 1, (nx, ny))
    x1 = h.solve(b)
    h.solve_resident(3)
    h.sync()
    assert np.array_equal(x1, h.solve(b))


def test_hholtz_adi_kat(emu):  # hholtz_adi.rs:176-207 through the device path
    import rustpde_b200 as R
    f = R.Field2(R.Space2(R.cheb_dirichlet(7), R.cheb_dirichlet(7)), lib=emu)
    x = R.HholtzAdi(f, [1.0, 1.0]).solve(np.tile(np.arange(1.0, 8.0), (7, 1)))
    y = np.array(
        [
            [-7.083e-03, -9.025e-03, -5.210e-03, 4.146e-03, 3.520e-03],
            [5.809e-04, 7.402e-04, 4.273e-04, -3.401e-04, -2.887e-04],
            [1.699e-04, 2.165e-04, 1.250e-04, -9.951e-05, -8.447e-05],
            [-1.007e-03, -1.283e-03, -7.406e-04, 5.895e-04, 5.004e-04],
            [-6.775e-04, -8.632e-04, -4.983e-04, 3.966e-04, 3.366e-04],
        ]
    )
    assert np.abs(x - y).max() < 1e-6


@pytest.mark.parametrize("which,kx,ky,nx,ny", [
    ("poisson", "cheb_neumann", "cheb_neumann", 16, 19),
    ("poisson", "cheb_dirichlet", "cheb_dirichlet", 8, 7),
    ("hholtz", "cheb_dirichlet", "cheb_dirichlet", 40, 33),
    ("poisson", "cheb_neumann", "cheb_neumann", 140, 34),   # more than one 128-row GEMM tile
])
def test_fast_diag_shared_eig(emu, which, kx, ky, nx, ny):
    e1, e2, _, _ = pc.check_tensor_shared_eig(emu, which, kx, ky, nx, ny)
    assert e1 <= pc.TOL and e2 <= pc.TOL, (e1, e2)


@pytest.mark.parametrize("which,kx,ky,nx,ny", [
    ("poisson", "cheb_neumann", "cheb_neumann", 16, 19),
    ("poisson", "cheb_dirichlet", "cheb_dirichlet", 9, 7),     # odd number of modes: 4 even + 3 odd
    ("hholtz", "cheb_dirichlet", "cheb_dirichlet", 40, 33),
    ("poisson", "cheb_neumann", "cheb_neumann", 141, 34),
])
def test_fast_diag_parity_split(emu, which, kx, ky, nx, ny):
    """Library's own set-up (even/odd block diagonalisation, exactly checkerboard Q and P) ->
    two half-size GEMM pairs; the oracle consumes the exported set-up data."""
    e1, e2 = pc.check_tensor_own_eig(emu, which, kx, ky, nx, ny)
    assert e1 <= pc.TOL and e2 <= pc.TOL, (e1, e2)


def test_poisson_kat_own_lapack(emu):  # poisson.rs:208-243, set-up through the library's own dgeev
    import rustpde_b200 as R
    from test_oracle_kat import POISSON_2D
    f = R.Field2(R.Space2(R.cheb_dirichlet(8), R.cheb_dirichlet(7)), lib=emu)
    b = np.tile(np.arange(1.0, 8.0), (8, 1))
    x = R.Poisson(f, [1.0, 1.0]).solve(b)
    assert np.abs(x - POISSON_2D).max() < 2e-6
    xc = R.Poisson(f, [1.0, 1.0]).solve(b * (1 + 1j))
    assert np.abs(xc - POISSON_2D * (1 + 1j)).max() < 2e-6


@pytest.mark.parametrize("nx,ny", [(16, 7), (32, 33)])
def test_fast_diag_fourier(emu, nx, ny):
    pc.check_tensor_fourier(emu, nx, ny)


@pytest.mark.parametrize("nx,ny,adiabatic", [(16, 17, True), (24, 20, False), (33, 33, True)])
def test_navier_confined(emu, nx, ny, adiabatic):
    err, derr, dn, do = pc.check_navier_steps(emu, False, nx, ny, 4, adiabatic=adiabatic, batch=2)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("nx,ny,adiabatic,own_eig", [(32, 33, True, False), (24, 33, False, True), (64, 65, True, True), (30, 65, True, False),
                                                     (32, 129, True, True)])  # 2-lane y tiles
def test_navier_confined_specialised_kernels(emu, nx, ny, adiabatic, own_eig):
    """Sizes served by the specialised x/y kernels (fast_x.cu / fast_y.cu): Bluestein DCT along x,
    power-of-two DCT along y; own_eig also exercises the parity-split GEMMs."""
    import rustpde_b200 as R
    assert R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, adiabatic, lib=emu).kernel_path() == (True, True)
    err, derr, dn, do = pc.check_navier_steps(emu, False, nx, ny, 4, adiabatic=adiabatic, batch=2, own_eig=own_eig)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("nx,ny,own_eig", [(33, 33, False), (65, 65, True), (129, 33, True), (257, 33, False)])
def test_navier_confined_pow2_period_x(emu, nx, ny, own_eig):
    """nx = 2^k + 1 (the usual Gauss-Lobatto choice): the DCT period along x is a power of two, the x kernels run the y
    kernels' dct_pow2 inside the Bluestein tile (fast_x.cu dct_x) -- these grids were on the lane programs before."""
    import rustpde_b200 as R
    assert R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, lib=emu).kernel_path() == (True, True)
    err, derr, dn, do = pc.check_navier_steps(emu, False, nx, ny, 4, tol=1e-9, batch=2, own_eig=own_eig)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("nx,ny", [(530, 129), (1030, 33), (300, 257), (200, 65)])
def test_navier_confined_specialised_kernels_partial_lanes(emu, nx, ny):
    """x lanes much shorter than the instantiated Bluestein length (2048 / 4096): the chunk-major coefficient tables
    have fewer rows of slots than the kernels' compile-time chunk bound, the staged strips are partly empty."""
    import rustpde_b200 as R
    assert R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, lib=emu).kernel_path()[0]
    err, derr, dn, do = pc.check_navier_steps(emu, False, nx, ny, 2, tol=1e-9, batch=2)
    assert max(derr) < 1e-9, (derr, dn, do)


def test_lane_program_and_specialised_kernels_agree(emu, monkeypatch):
    """The generic lane programs and the specialised kernels implement the same step."""
    import rustpde_b200 as R
    outs = []
    for no_fast in ("1", "0"):
        monkeypatch.setenv("RUSTPDE_B200_NO_FAST", no_fast)
        n = R.Navier2D.new(32, 33, 1e5, 1.0, 0.01, 1.0, True, lib=emu)
        n.set_velocity(0.2, 1.0, 1.0)
        n.set_temperature(0.2, 1.0, 1.0)
        n.update(3)
        outs.append((n.temp.vhat, n.ux.vhat, n.uy.vhat, n.pres[0].vhat))
    for a, b in zip(*outs):
        assert pc.rel(a, b) < 1e-11


@pytest.mark.parametrize("nx,ny", [(16, 17), (24, 20)])
def test_navier_periodic(emu, nx, ny):
    err, derr, dn, do = pc.check_navier_steps(emu, True, nx, ny, 4, batch=2)
    assert max(derr) < 1e-9, (derr, dn, do)


@pytest.mark.parametrize("nx,ny", [(32, 33), (64, 65), (32, 65), (128, 129), (32, 129), (128, 33)])  # 129 / 128: 2-lane tiles
def test_navier_periodic_specialised_kernels(emu, nx, ny):
    """Sizes served by the specialised periodic kernels (fast_p.cu): pow2 r2c/c2r along x, pow2 DCT along y."""
    import rustpde_b200 as R
    assert R.Navier2D.new_periodic(nx, ny, 1e5, 1.0, 0.01, 1.0, lib=emu).kernel_path()[0]
    err, derr, dn, do = pc.check_navier_steps(emu, True, nx, ny, 4, batch=2)
    assert max(derr) < 1e-9, (derr, dn, do)


def test_navier_periodic_specialised_aspect_no_dealias(emu):
    import oracle as O
    import rustpde_b200 as R
    o = O.Navier2D.new_periodic(32, 33, 1e4, 0.7, 0.02, 1.5, banded=True)
    n = R.Navier2D.new_periodic(32, 33, 1e4, 0.7, 0.02, 1.5, lib=emu)
    o.dealias = False
    n.dealias = False
    assert n.kernel_path()[0]
    for x in (n, o):
        x.set_velocity(0.1, 2.0, 1.0)
        x.set_temperature(0.3, 1.0, 2.0)
    n.update(3)
    for _ in range(3):
        o.update()
    err = pc.navier_field_errors(n, o)
    assert max(err.values()) < 1e-9, err


def test_navier_aspect_and_no_dealias(emu):
    import oracle as O
    import rustpde_b200 as R
    o = O.Navier2D.new_periodic(16, 17, 1e4, 0.7, 0.02, 1.5, banded=True)
    n = R.Navier2D.new_periodic(16, 17, 1e4, 0.7, 0.02, 1.5, lib=emu)
    o.dealias = False
    n.dealias = False
    for x in (n, o):
        x.set_velocity(0.1, 2.0, 1.0)
        x.set_temperature(0.3, 1.0, 2.0)
    n.update(3)
    for _ in range(3):
        o.update()
    err = pc.navier_field_errors(n, o)
    assert max(err.values()) < 1e-9, err


def test_shape_errors(emu):
    import rustpde_b200 as R
    f = R.Field2(R.Space2(R.cheb_dirichlet(8), R.cheb_dirichlet(9)), lib=emu)
    with pytest.raises(R.RustpdeError):
        f.v = np.zeros((8, 8))
    with pytest.raises(R.RustpdeError):
        R.HholtzAdi(f, [1.0, 1.0]).solve(np.zeros((9, 9)))
    with pytest.raises(R.RustpdeError):
        f.from_ortho(np.zeros((6, 7)))


@pytest.mark.parametrize("periodic,nx,ny", [(False, 32, 33), (True, 32, 33)])
def test_staged_state_upload(emu, periodic, nx, ny):
    assert pc.check_staged_upload(emu, periodic, nx, ny)


def test_commit_without_stage_fails(emu):
    import rustpde_b200 as R
    n = R.Navier2D.new(16, 17, 1e4, 1.0, 0.01, 1.0, True, lib=emu)
    with pytest.raises(R.RustpdeError):
        n.commit_staged()


@pytest.mark.parametrize("periodic,nx,ny", [(False, 24, 33), (True, 32, 33), (False, 20, 20)])
def test_host_api_additions(emu, periodic, nx, ny):
    """profile() == update() after an outside pressure write, eig round trip, fetch_state, div_async, row slabs,
    device-side averages (deterministic two-stage reductions)."""
    assert pc.check_host_api_additions(emu, periodic, nx, ny)


def test_band_rel_metric():
    """The banded metric must see an error confined to the small high modes that the global max-norm misses."""
    rng = np.random.default_rng(0)
    b = rng.uniform(-1, 1, (64, 64)) * np.logspace(0, -12, 64)[None, :]
    a = b.copy()
    a[:, 32:] *= 1.5  # coefficients of relative size 1e-6 and below: invisible to the global max-norm at 1e-6
    assert pc.rel(a, b) < 1e-6
    assert pc.band_rel(a, b) > 0.1
    assert pc.band_rel(b, b) == 0.0


def test_column_scan_kernels_agree_with_tile_kernels(emu, monkeypatch):
    """The schedules of the x-direction banded sweeps give the same step to rounding: tile kernels only
    (RUSTPDE_B200_XW=0, 13 launches), the default (warp-serial ADI-x sweeps of fast_xw.cu, 14 launches), divergence and
    projection as warp-serial sweeps too (RUSTPDE_B200_XW=2, 14 launches) and the streaming column scans
    (RUSTPDE_B200_XS=1, fast_xs.cu, 15 launches); the last one also against the oracle."""
    import rustpde_b200 as R
    outs = []
    for xw, xs in (("0", "0"), ("1", "0"), ("2", "0"), ("1", "1")):
        monkeypatch.setenv("RUSTPDE_B200_XW", xw)
        monkeypatch.setenv("RUSTPDE_B200_XS", xs)
        n = R.Navier2D.new(40, 33, 1e5, 1.0, 0.01, 1.0, True, lib=emu)
        n.set_velocity(0.2, 1.0, 1.0)
        n.set_temperature(0.2, 1.0, 1.0)
        n.update(4)
        outs.append([np.array(f.vhat) for f in (n.temp, n.ux, n.uy, n.pres[0])] + [n.launches_per_step()])
    assert [o[-1] for o in outs] == [13, 14, 14, 15]
    for other in outs[1:]:
        for a, b in zip(outs[0][:-1], other[:-1]):
            assert pc.rel(a, b) <= 1e-12
    monkeypatch.setenv("RUSTPDE_B200_XS", "1")
    err, derr, dn, do = pc.check_navier_steps(emu, False, 40, 33, 4, tol=1e-9, batch=2, own_eig=True)
    assert max(derr) < 1e-9


def test_periodic_row_sweeps_agree_with_tile_kernels(emu, monkeypatch):
    """The per-mode Helmholtz / Poisson passes and the projection + pressure update of the periodic step as warp-serial
    row sweeps (fast_pw.cu, chosen for large row counts; forced here with RUSTPDE_B200_PW=1) versus the tile kernels
    pk_hholtz / pk_divpois / pk_project (RUSTPDE_B200_PW=0): same step to rounding, ragged row counts included."""
    import rustpde_b200 as R
    for nx, ny in ((32, 33), (64, 65), (128, 33)):
        outs = []
        for pw in ("0", "1"):
            monkeypatch.setenv("RUSTPDE_B200_PW", pw)
            n = R.Navier2D.new_periodic(nx, ny, 1e5, 1.0, 0.01, 1.0, lib=emu)
            assert n.kernel_path()[0]
            n.set_velocity(0.2, 1.0, 1.0)
            n.set_temperature(0.2, 1.0, 1.0)
            n.update(4)
            outs.append([np.array(f.vhat) for f in (n.temp, n.ux, n.uy, n.pres[0])])
        for a, b in zip(*outs):
            assert 0.0 < pc.rel(a, b) <= 1e-12  # (not bitwise equal: the two paths really are different kernels)
