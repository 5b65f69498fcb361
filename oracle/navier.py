"""Oracle restatement of rustpde::navier::Navier2D (src/navier/navier.rs,
conv_term.rs, functions.rs) and the integrate() loop (src/lib.rs:132-187).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Same op order as the
reference; temporaries are plain numpy arrays.
"""
import math

import numpy as np

from .field import Field2
from .funspace import (
    Space2,
    cheb_dirichlet,
    cheb_dirichlet_bc,
    cheb_neumann,
    chebyshev,
    fourier_r2c,
)
from .solver import Hholtz, HholtzAdi, Poisson

MAX_TIMESTEP = 10_000_000  # src/lib.rs:132


def get_nu(ra, pr, height):  # navier.rs:46-49
    return math.sqrt(pr / (ra / height ** 3.0))


def get_ka(ra, pr, height):  # navier.rs:52-55
    return math.sqrt(1.0 / ((ra / height ** 3.0) * pr))


def conv_term(field, deriv_field, u, deriv, scale):  # conv_term.rs:22-42
    deriv_field.vhat = field.gradient(deriv, scale)
    deriv_field.backward()
    return u * deriv_field.v


def dealias(field):  # navier.rs:1022-1032
    n_x = field.vhat.shape[0] * 2 // 3
    n_y = field.vhat.shape[1] * 2 // 3
    field.vhat[n_x:, :] = 0.0
    field.vhat[:, n_y:] = 0.0


def _norm_coords(field):  # navier.rs:1042-1045
    x, y = field.x
    xs = (x - x[0]) / (x[len(x) - 1] - x[0])
    ys = (y - y[0]) / (y[len(y) - 1] - y[0])
    return xs, ys


def apply_sin_cos(field, amp, m, n):  # navier.rs:1035-1054
    nx, ny = field.v.shape
    xs, ys = _norm_coords(field)
    ax, ay = math.pi * m, math.pi * n
    field.v = amp * np.sin(ax * xs[:nx])[:, None] * np.cos(ay * ys[:ny])[None, :]
    field.forward()


def apply_cos_sin(field, amp, m, n):  # navier.rs:1057-1076
    nx, ny = field.v.shape
    xs, ys = _norm_coords(field)
    ax, ay = math.pi * m, math.pi * n
    field.v = amp * np.cos(ax * xs[:nx])[:, None] * np.sin(ay * ys[:ny])[None, :]
    field.forward()


class Navier2D:
    """navier.rs:153-195.  Use Navier2D.new / Navier2D.new_periodic."""

    def __init__(self):
        raise TypeError("use Navier2D.new(...) or Navier2D.new_periodic(...)")

    @classmethod
    def _bare(cls):
        return object.__new__(cls)

    # navier.rs:219-307
    @classmethod
    def new(cls, nx, ny, ra, pr, dt, aspect, adiabatic, banded=False, eig_data=None):
        s = cls._bare()
        s.periodic = False
        s.scale = [aspect, 1.0]
        s.nu = get_nu(ra, pr, s.scale[1] * 2.0)
        s.ka = get_ka(ra, pr, s.scale[1] * 2.0)
        s.ux = Field2(Space2(cheb_dirichlet(nx), cheb_dirichlet(ny)))
        s.uy = Field2(Space2(cheb_dirichlet(nx), cheb_dirichlet(ny)))
        if adiabatic:
            s.temp = Field2(Space2(cheb_neumann(nx), cheb_dirichlet(ny)))
        else:
            s.temp = Field2(Space2(cheb_dirichlet(nx), cheb_dirichlet(ny)))
        s.pres = [
            Field2(Space2(chebyshev(nx), chebyshev(ny))),
            Field2(Space2(cheb_neumann(nx), cheb_neumann(ny))),
        ]
        s.field = Field2(Space2(chebyshev(nx), chebyshev(ny)))
        sx2, sy2 = s.scale[0] ** 2.0, s.scale[1] ** 2.0
        s.solver = [
            HholtzAdi(s.ux, [dt * s.nu / sx2, dt * s.nu / sy2], banded=banded),
            HholtzAdi(s.uy, [dt * s.nu / sx2, dt * s.nu / sy2], banded=banded),
            HholtzAdi(s.temp, [dt * s.ka / sx2, dt * s.ka / sy2], banded=banded),
            Poisson(s.pres[1], [1.0 / sx2, 1.0 / sy2], banded=banded, eig_data=eig_data),
        ]
        s._finish(ra, pr, dt)
        s.fieldbc = cls.bc_rbc(nx, ny)
        # navier.rs:304 random_disturbance(0.1) is unseeded -> callers must set
        # deterministic ICs (set_velocity + set_temperature); fields start at 0.
        return s

    # navier.rs:384-467
    @classmethod
    def new_periodic(cls, nx, ny, ra, pr, dt, aspect, banded=False):
        s = cls._bare()
        s.periodic = True
        s.scale = [aspect, 1.0]
        s.nu = get_nu(ra, pr, s.scale[1] * 2.0)
        s.ka = get_ka(ra, pr, s.scale[1] * 2.0)
        s.ux = Field2(Space2(fourier_r2c(nx), cheb_dirichlet(ny)))
        s.uy = Field2(Space2(fourier_r2c(nx), cheb_dirichlet(ny)))
        s.temp = Field2(Space2(fourier_r2c(nx), cheb_dirichlet(ny)))
        s.pres = [
            Field2(Space2(fourier_r2c(nx), chebyshev(ny))),
            Field2(Space2(fourier_r2c(nx), cheb_neumann(ny))),
        ]
        s.field = Field2(Space2(fourier_r2c(nx), chebyshev(ny)))
        sx2, sy2 = s.scale[0] ** 2.0, s.scale[1] ** 2.0
        s.solver = [
            Hholtz(s.ux, [dt * s.nu / sx2, dt * s.nu / sy2], banded=banded),
            Hholtz(s.uy, [dt * s.nu / sx2, dt * s.nu / sy2], banded=banded),
            Hholtz(s.temp, [dt * s.ka / sx2, dt * s.ka / sy2], banded=banded),
            Poisson(s.pres[1], [1.0 / sx2, 1.0 / sy2], banded=banded),
        ]
        s._finish(ra, pr, dt)
        s.fieldbc = cls.bc_rbc_periodic(nx, ny)
        return s

    def _finish(self, ra, pr, dt):
        self.ra, self.pr, self.dt = ra, pr, dt
        self.time = 0.0
        self.dealias = True
        self.solid = None        # [mask, value] (navier.rs:191, solid_masks.rs)
        self.statistics = None   # navier.rs:195
        self.diagnostics = {"time": [], "Nu": [], "Nuvol": [], "Re": []}
        # navier.rs:502-514 _scale
        for f in (self.temp, self.ux, self.uy, self.pres[0]):
            f.x[0] = f.x[0] * self.scale[0]
            f.x[1] = f.x[1] * self.scale[1]
            f.dx[0] = f.dx[0] * self.scale[0]
            f.dx[1] = f.dx[1] * self.scale[1]

    # navier.rs:314-332
    @staticmethod
    def bc_rbc(nx, ny):
        x_base = chebyshev(nx)
        fieldbc = Field2(Space2(x_base, cheb_dirichlet_bc(ny)))
        bc = np.zeros_like(fieldbc.vhat)
        bc[:, 0] = 0.5
        bc[:, 1] = -0.5
        fieldbc.vhat = x_base.forward(bc, 0)
        fieldbc.backward()
        fieldbc.forward()
        return fieldbc

    # navier.rs:474-492
    @staticmethod
    def bc_rbc_periodic(nx, ny):
        x_base = fourier_r2c(nx)
        fieldbc = Field2(Space2(x_base, cheb_dirichlet_bc(ny)))
        bc = np.zeros((nx, 2))
        bc[:, 0] = 0.5
        bc[:, 1] = -0.5
        fieldbc.vhat = x_base.forward(bc, 0)
        fieldbc.backward()
        fieldbc.forward()
        return fieldbc

    # ---- initial conditions, navier.rs:927-936 ---------------------------
    def set_velocity(self, amp, m, n):
        apply_sin_cos(self.ux, amp, m, n)
        apply_cos_sin(self.uy, -amp, m, n)

    def set_temperature(self, amp, m, n):
        apply_cos_sin(self.temp, -amp, m, n)

    def reset_time(self):
        self.time = 0.0

    def new_work_field(self):
        """Field2::new(&navier.field.space) (statistics.rs:60-65)."""
        return Field2(self.field.space)

    # ---- convection, navier.rs:538-616 -----------------------------------
    def _finish_conv(self, conv):
        self.field.v = conv
        self.field.forward()
        if self.dealias:
            dealias(self.field)
        return self.field.vhat.copy()

    def conv_temp(self, ux, uy):
        conv = conv_term(self.temp, self.field, ux, [1, 0], self.scale)
        conv = conv + conv_term(self.temp, self.field, uy, [0, 1], self.scale)
        if self.fieldbc is not None:
            conv = conv + conv_term(self.fieldbc, self.field, ux, [1, 0], self.scale)
            conv = conv + conv_term(self.fieldbc, self.field, uy, [0, 1], self.scale)
        if self.solid is not None:  # navier.rs:552-560: volume penalisation, eta = 1e-2
            eta = 1e-2
            self.temp.backward()
            if self.fieldbc is None:
                damp = -1.0 / eta * self.solid[0] * (self.temp.v - self.solid[1])
            else:
                damp = -1.0 / eta * self.solid[0] * (self.temp.v + self.fieldbc.v - self.solid[1])
            conv = conv - damp
        return self._finish_conv(conv)

    def conv_ux(self, ux, uy):
        conv = conv_term(self.ux, self.field, ux, [1, 0], self.scale)
        conv = conv + conv_term(self.ux, self.field, uy, [0, 1], self.scale)
        if self.solid is not None:  # navier.rs:580-584
            conv = conv - (-1.0 / 1e-2 * self.solid[0] * ux)
        return self._finish_conv(conv)

    def conv_uy(self, ux, uy):
        conv = conv_term(self.uy, self.field, ux, [1, 0], self.scale)
        conv = conv + conv_term(self.uy, self.field, uy, [0, 1], self.scale)
        if self.solid is not None:  # navier.rs:604-608
            conv = conv - (-1.0 / 1e-2 * self.solid[0] * uy)
        return self._finish_conv(conv)

    # ---- implicit solves, navier.rs:622-674 ------------------------------
    def solve_ux(self, ux, uy):
        rhs = self.ux.to_ortho()
        rhs = rhs - self.pres[0].gradient([1, 0], self.scale) * self.dt
        rhs = rhs - self.conv_ux(ux, uy) * self.dt
        self.ux.vhat = self.solver[0].solve(rhs, 0)

    def solve_uy(self, ux, uy, buoy):
        rhs = self.uy.to_ortho()
        rhs = rhs - self.pres[0].gradient([0, 1], self.scale) * self.dt
        rhs = rhs + buoy * self.dt
        rhs = rhs - self.conv_uy(ux, uy) * self.dt
        self.uy.vhat = self.solver[1].solve(rhs, 0)

    def solve_temp(self, ux, uy):
        rhs = self.temp.to_ortho()
        if self.fieldbc is not None:
            rhs = rhs + self.fieldbc.gradient([2, 0], self.scale) * self.dt * self.ka
            rhs = rhs + self.fieldbc.gradient([0, 2], self.scale) * self.dt * self.ka
        rhs = rhs - self.conv_temp(ux, uy) * self.dt
        self.temp.vhat = self.solver[2].solve(rhs, 0)

    # ---- projection, navier.rs:683-721 -----------------------------------
    def divergence(self):
        return self.ux.gradient([1, 0], self.scale) + self.uy.gradient([0, 1], self.scale)

    def solve_pres(self, f):
        self.pres[1].vhat = self.solver[3].solve(f, 0)
        self.pres[1].vhat[0, 0] = 0.0

    def project_velocity(self, c):
        dpdx = self.pres[1].gradient([1, 0], self.scale)
        dpdy = self.pres[1].gradient([0, 1], self.scale)
        ux_old = self.ux.vhat.copy()
        uy_old = self.uy.vhat.copy()
        self.ux.from_ortho(dpdx)
        self.uy.from_ortho(dpdy)
        self.ux.vhat = self.ux.vhat * (-c) + ux_old
        self.uy.vhat = self.uy.vhat * (-c) + uy_old

    def update_pres(self, div):
        self.pres[0].vhat = self.pres[0].vhat - div * self.nu
        self.pres[0].vhat = self.pres[0].vhat + self.pres[1].to_ortho() * (1.0 / self.dt)

    # ---- Integrate, navier.rs:737-765 ------------------------------------
    def update(self):
        that = self.temp.to_ortho()
        if self.fieldbc is not None:
            that = that + self.fieldbc.to_ortho()
        self.ux.backward()
        self.uy.backward()
        ux = self.ux.v.copy()
        uy = self.uy.v.copy()
        self.solve_ux(ux, uy)
        self.solve_uy(ux, uy, that)
        div = self.divergence()
        self.solve_pres(div)
        self.project_velocity(1.0)
        self.update_pres(div)
        self.solve_temp(ux, uy)
        self.time += self.dt

    def get_time(self):
        return self.time

    def get_dt(self):
        return self.dt

    def div_norm(self):  # navier.rs:855-879
        d = self.divergence()
        return float(np.sqrt(np.sum(d.real ** 2 + d.imag ** 2)))

    def exit(self):
        return math.isnan(self.div_norm())

    def callback(self):  # navier.rs:816-842 (no HDF5 in this image)
        nu, nuvol, re = self.eval_nu(), self.eval_nuvol(), self.eval_re()
        self.diagnostics["time"].append(self.time)
        self.diagnostics["Nu"].append(nu)
        self.diagnostics["Nuvol"].append(nuvol)
        self.diagnostics["Re"].append(re)

    # ---- diagnostics, src/navier/functions.rs ----------------------------
    def eval_nu(self):  # functions.rs:12-36
        f = self.field
        f.vhat = self.temp.to_ortho()
        if self.fieldbc is not None:
            f.vhat = f.vhat + self.fieldbc.to_ortho()
        dtdz = f.gradient([0, 1], None) * -1.0
        dtdz = dtdz * (1.0 / (self.scale[1] / 2.0))
        f.vhat = dtdz
        f.backward()
        x_avg = f.average_axis(0)
        return float((x_avg[len(x_avg) - 1] + x_avg[0]) / 2.0)

    def eval_nuvol(self):  # functions.rs:42-75
        f = self.field
        f.vhat = self.temp.to_ortho()
        if self.fieldbc is not None:
            f.vhat = f.vhat + self.fieldbc.to_ortho()
        f.backward()
        self.uy.backward()
        uy_temp = f.v * self.uy.v
        dtdz = f.gradient([0, 1], None) / (self.scale[1] * -1.0)
        f.vhat = dtdz
        f.backward()
        f.v = (f.v + uy_temp / self.ka) * 2.0 * self.scale[1]
        return f.average()

    def eval_re(self):  # functions.rs:82-101
        self.ux.backward()
        self.uy.backward()
        ekin = self.ux.v ** 2 + self.uy.v ** 2
        self.field.v = np.sqrt(ekin)
        self.field.v = self.field.v * (2.0 * self.scale[1] / self.nu)
        return self.field.average()

    def eval_ekin(self):
        """<(ux^2+uy^2)/2> with the reference's Field::average weights
        (no dedicated reference function)."""
        self.ux.backward()
        self.uy.backward()
        self.field.v = 0.5 * (self.ux.v ** 2 + self.uy.v ** 2)
        return self.field.average()


def integrate(pde, max_time, save_intervall=None):  # src/lib.rs:155-187
    timestep = 0
    eps_dt = pde.get_dt() * 1e-4
    while True:
        pde.update()
        timestep += 1
        if save_intervall is not None:
            t, dt = pde.get_time(), pde.get_dt()
            if (t % save_intervall) < dt / 2.0 or (t % save_intervall) > save_intervall - dt / 2.0:
                pde.callback()
        if pde.get_time() + eps_dt >= max_time:
            break
        if timestep >= MAX_TIMESTEP:
            break
        if pde.exit():
            break
    return timestep
