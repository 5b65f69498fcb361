"""Solid-mask penalisation (navier.rs:552-608, solid_masks.rs:34-175) and Statistics (statistics.rs:10-247): device
path versus the oracle.  CPU emulation here; test_gpu_parity.py::test_solid_masks_gpu runs the same on the B200."""
import numpy as np
import pytest

import oracle as O
import parity_cases as pc
import rustpde_b200 as R


def test_mask_generators_match_the_restatement():
    x = -np.cos(np.pi * np.arange(33) / 32)
    y = -np.cos(np.pi * np.arange(41) / 40)
    for a, b in ((R.solid_cylinder_inner(x, y, 0.2, 0.0, 0.2), O.solid_cylinder_inner(x, y, 0.2, 0.0, 0.2)),
                 (R.solid_roughness_sinusoid(x, y, 0.1, 10.0), O.solid_roughness_sinusoid(x, y, 0.1, 10.0)),
                 (R.solid_porosity(x, y, 0.3, 0.6), O.solid_porosity(x, y, 0.3, 0.6))):
        assert np.abs(a[0] - b[0]).max() < 1e-14 and np.abs(a[1] - b[1]).max() < 1e-14
        assert a[0].max() > 0.5  # something is solid


def check_solid(lib, periodic, nx, ny, steps=6, tol=1e-9):
    n, o = pc.make_navier_pair(lib, periodic, nx, ny, 1e5, 1.0, 0.01, ics=False, own_eig=not periodic)
    x, y = o.temp.x[0][:nx], o.temp.x[1]
    if periodic:
        mask = O.solid_roughness_sinusoid(x, y, 0.1, 2.0)
    else:
        mask = O.solid_cylinder_inner(x, y, 0.2, 0.0, 0.2)
        mask[1] = mask[1] + 0.25 * mask[0]  # non-zero target temperature inside the body
    n.solid = mask
    o.solid = [mask[0].copy(), mask[1].copy()]
    for m in (n, o):
        m.set_velocity(0.2, 1.0, 1.0)
        m.set_temperature(0.2, 1.0, 1.0)
    assert n.kernel_path()[0]
    n.update(steps)
    for _ in range(steps):
        o.update()
    err = pc.navier_field_errors(n, o)
    assert max(err.values()) <= tol, err
    # the mask really acts: compare with a run without it
    n2, _ = pc.make_navier_pair(lib, periodic, nx, ny, 1e5, 1.0, 0.01, own_eig=not periodic)
    n2.update(steps)
    assert pc.rel(n2.ux.vhat, n.ux.vhat) > 1e-3
    return err


@pytest.mark.parametrize("periodic,nx,ny", [(False, 24, 33), (True, 32, 33)])
def test_solid_masks(emu, periodic, nx, ny):
    check_solid(emu, periodic, nx, ny)


def test_solid_needs_specialised_kernels(emu):
    n = R.Navier2D.new(20, 20, 1e5, 1.0, 0.01, 1.0, True, lib=emu)
    n.solid = [np.ones((20, 20)), np.zeros((20, 20))]
    with pytest.raises(R.RustpdeError):
        n.update(1)


def check_statistics(lib, periodic, nx, ny, tmp_path):
    n, o = pc.make_navier_pair(lib, periodic, nx, ny, 1e5, 1.0, 0.01, own_eig=not periodic)
    sn, so = R.Statistics(n, 0.02, 0.04), O.Statistics(o, 0.02, 0.04)
    for k in range(3):
        n.update(2)
        for _ in range(2):
            o.update()
        sn.update(n.temp.to_ortho() + n.tempbc_ortho(), n.ux.to_ortho(), n.uy.to_ortho(), n.time)
        so.update(o.temp.to_ortho() + o.fieldbc.to_ortho(), o.ux.to_ortho(), o.uy.to_ortho(), o.time)
    assert sn.num_save == so.num_save == 3 and abs(sn.avg_time - so.avg_time) < 1e-14
    for a, b in ((sn.t_avg, so.t_avg), (sn.ux_avg, so.ux_avg), (sn.uy_avg, so.uy_avg), (sn.nusselt, so.nusselt)):
        assert pc.rel(a.vhat, b.vhat) <= 1e-9
    # write / read round trip (statistics.rs:163-211)
    path = str(tmp_path / "statistics.rpsnap")
    sn.write(path)
    s2 = R.Statistics(n, 0.02, 0.04)
    s2.read(path)
    assert s2.num_save == 3 and np.array_equal(s2.nusselt.vhat, sn.nusselt.vhat)
    return True


@pytest.mark.parametrize("periodic,nx,ny", [(False, 24, 33), (True, 32, 33)])
def test_statistics(emu, tmp_path, periodic, nx, ny):
    assert check_statistics(emu, periodic, nx, ny, tmp_path)


def test_callback_updates_statistics(emu, tmp_path):
    n = R.Navier2D.new(24, 33, 1e5, 1.0, 0.01, 1.0, True, lib=emu)
    n.set_velocity(0.2, 1.0, 1.0)
    n.set_temperature(0.2, 1.0, 1.0)
    n.statistics = R.Statistics(n, 0.02, 0.02)
    n.update(2)
    n.callback(data_dir=str(tmp_path))
    assert n.statistics.num_save == 1 and (tmp_path / "statistics.rpsnap").exists()
