"""Checkpoint / restart (navier.rs:956-1014, field/write.rs:82-115, field/read.rs:56-122) through the C ABI, on the
CPU emulation build; the GPU variant is test_gpu_parity.py::test_snapshot_roundtrip_gpu."""
import numpy as np
import pytest

import oracle as O
import rustpde_b200 as R
from rustpde_b200 import snapshot as S


def make(lib, periodic, nx, ny):
    n = (R.Navier2D.new_periodic(nx, ny, 1e5, 1.0, 0.01, 1.0, lib=lib) if periodic
         else R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True, lib=lib))
    n.set_velocity(0.2, 1.0, 1.0)
    n.set_temperature(0.2, 1.0, 1.0)
    return n


def check_snapshot_roundtrip(lib, periodic, nx, ny, tmp_path):
    a = make(lib, periodic, nx, ny)
    a.update(3)
    path = str(tmp_path / ("snap_%d.rpsnap" % periodic))
    a.write(path)
    d = S.read_snapshot(path)
    # dataset names and shapes of the reference layout
    want = {"x", "dx", "y", "dy", "time", "ra", "pr", "nu", "kappa"}
    for g in ("temp", "ux", "uy", "pres"):
        want |= {g + "/v"} | ({g + "/vhat_re", g + "/vhat_im"} if periodic else {g + "/vhat"})
    assert set(d) == want
    assert d["time"].shape == () and abs(float(d["time"]) - a.time) < 1e-15 and float(d["nu"]) == a.nu and float(d["kappa"]) == a.ka
    assert d["ux/v"].shape == (nx, ny) and d["x"].shape[0] >= nx and d["y"].shape == (ny,)
    vh = d["ux/vhat_re"] + 1j * d["ux/vhat_im"] if periodic else d["ux/vhat"]
    assert np.array_equal(vh, a.ux.vhat)
    # temp/v includes the boundary field (navier.rs:988-991): T = +0.5 at y = -1 ... -0.5 at y = +1 plus the perturbation
    a.temp.backward()
    y = d["y"]
    assert np.abs(d["temp/v"] - (a.temp.v + (-0.5 * y)[None, :])).max() < 1e-12
    # restart: same state, same continuation (bit for bit: vhat is stored exactly)
    b = make(lib, periodic, nx, ny)
    b.read(path)
    assert abs(b.time - a.time) < 1e-15
    a.update(2)
    b.update(2)
    for fa, fb in ((a.temp, b.temp), (a.ux, b.ux), (a.uy, b.uy), (a.pres[0], b.pres[0])):
        assert np.array_equal(fa.vhat, fb.vhat)
    # python writer / reader agree with the C++ one
    p2 = str(tmp_path / "copy.rpsnap")
    S.write_snapshot(p2, d)
    d2 = S.read_snapshot(p2)
    assert set(d2) == set(d) and all(np.array_equal(d[k], d2[k]) for k in d)
    # restart on another grid: the block both shapes share is copied, the rest keeps its values (read.rs:113-122)
    c = make(lib, periodic, nx + 8, ny + 8)
    before = np.array(c.ux.vhat)
    c.read(path)
    after = np.array(c.ux.vhat)
    r, k = a.ux.shape_spectral if periodic else (nx - 2, ny - 2)
    r = min(r, after.shape[0])
    assert np.array_equal(after[:r, :k], vh[:r, :k])
    assert np.array_equal(after[:, k:], before[:, k:])
    return True


@pytest.mark.parametrize("periodic,nx,ny", [(False, 24, 33), (True, 32, 33)])
def test_snapshot_roundtrip(emu, tmp_path, periodic, nx, ny):
    assert check_snapshot_roundtrip(emu, periodic, nx, ny, tmp_path)


def test_callback_writes_snapshot_and_info(emu, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # integrate() -> callback() writes to ./data like the reference (navier.rs:775-853)
    n = make(emu, False, 20, 20)
    R.integrate(n, 0.05, 0.02)
    assert sorted(p.name for p in (tmp_path / "data").iterdir())[-1] == "info.txt"
    # callback() ran at t = 0.02 and 0.04 (src/lib.rs:165-172)
    n.callback(data_dir=str(tmp_path))
    files = sorted(p.name for p in tmp_path.iterdir())
    assert "info.txt" in files and any(f.startswith("flow") and f.endswith(".rpsnap") for f in files)
    assert len(n.diagnostics["Nu"]) == 3
