"""Oracle restatement of rustpde::navier::Navier2DAdjoint (src/navier/navier_adjoint.rs:128-1068): steady-state
adjoint descent (Farazmand 2016).  TEST INFRASTRUCTURE (see oracle/__init__.py); same op order as the reference.

The constructors of the reference build the inner Navier2D with its unseeded random disturbance (navier.rs:304); the
inner solver's ux / uy / temp are overwritten before every use (update_residual, 739-766), its pressure starts at 0,
so nothing random survives -- fields start at zero here like everywhere else in this oracle.
"""
import math

import numpy as np

from .field import Field2
from .funspace import Space2, cheb_dirichlet, cheb_neumann, chebyshev, fourier_r2c
from .navier import Navier2D, apply_cos_sin, apply_sin_cos, conv_term, dealias, get_ka, get_nu
from .solver import Hholtz, Poisson

RES_TOL = 1e-8  # navier_adjoint.rs:122


def norm_l2(a):  # navier_adjoint.rs:916-926
    return float(np.sqrt(np.sum(a.real ** 2 + a.imag ** 2)))


class Navier2DAdjoint:
    def __init__(self):
        raise TypeError("use Navier2DAdjoint.new(...) or .new_periodic(...)")

    @classmethod
    def _make(cls, nx, ny, ra, pr, dt, aspect, adiabatic, periodic, banded, eig):
        """eig: None or dict with optional keys 'smooth_u', 'smooth_t', 'pres', 'navier_pres' -> (lam, Q, P)."""
        s = object.__new__(cls)
        eig = eig or {}
        s.periodic = periodic
        s.scale = [aspect, 1.0]
        s.nu = get_nu(ra, pr, s.scale[1] * 2.0)
        s.ka = get_ka(ra, pr, s.scale[1] * 2.0)
        bx_u = (lambda: fourier_r2c(nx)) if periodic else (lambda: cheb_dirichlet(nx))
        bx_t = (lambda: fourier_r2c(nx)) if periodic else ((lambda: cheb_neumann(nx)) if adiabatic else (lambda: cheb_dirichlet(nx)))
        bx_o = (lambda: fourier_r2c(nx)) if periodic else (lambda: chebyshev(nx))
        bx_n = (lambda: fourier_r2c(nx)) if periodic else (lambda: cheb_neumann(nx))
        mk = lambda bx, by: Field2(Space2(bx(), by(ny)))
        s.ux = [mk(bx_u, cheb_dirichlet), mk(bx_u, cheb_dirichlet)]
        s.uy = [mk(bx_u, cheb_dirichlet), mk(bx_u, cheb_dirichlet)]
        s.temp = [mk(bx_t, cheb_dirichlet), mk(bx_t, cheb_dirichlet)]
        s.dt_navier = 1e-2  # :231
        if periodic:
            s.navier = Navier2D.new_periodic(nx, ny, ra, pr, s.dt_navier, aspect, banded=banded)
        else:
            s.navier = Navier2D.new(nx, ny, ra, pr, s.dt_navier, aspect, adiabatic, banded=banded, eig_data=eig.get("navier_pres"))
        s.pres = [mk(bx_o, chebyshev), mk(bx_n, cheb_neumann)]
        s.field = mk(bx_o, chebyshev)
        sx2, sy2 = s.scale[0] ** 2.0, s.scale[1] ** 2.0
        s.solver = [Poisson(s.pres[1], [1.0 / sx2, 1.0 / sy2], banded=banded, eig_data=eig.get("pres"))]
        w = 1e0  # weight_laplacian, :251
        s.smoother = [
            Hholtz(s.ux[1], [w / sx2, w / sy2], banded=banded, eig_data=eig.get("smooth_u")),
            Hholtz(s.uy[1], [w / sx2, w / sy2], banded=banded, eig_data=eig.get("smooth_u")),
            Hholtz(s.temp[1], [w / sx2, w / sy2], banded=banded, eig_data=eig.get("smooth_t")),
        ]
        s.fields_unsmoothed = [np.zeros_like(s.field.vhat) for _ in range(3)]
        s.ra, s.pr, s.dt, s.time = ra, pr, dt, 0.0
        s.res_tol = RES_TOL
        s.dealias = True
        s.diagnostics = {"time": [], "Nu": [], "Nuvol": [], "Re": []}
        s.fieldbc = Navier2D.bc_rbc_periodic(nx, ny) if periodic else Navier2D.bc_rbc(nx, ny)
        return s

    @classmethod
    def new(cls, nx, ny, ra, pr, dt, aspect, adiabatic, banded=True, eig=None):  # :197-342
        return cls._make(nx, ny, ra, pr, dt, aspect, adiabatic, False, banded, eig)

    @classmethod
    def new_periodic(cls, nx, ny, ra, pr, dt, aspect, banded=True):  # :361-496
        return cls._make(nx, ny, ra, pr, dt, aspect, True, True, banded, None)

    # ---- :994-1008 ----
    def set_velocity(self, amp, m, n):
        apply_sin_cos(self.ux[0], amp, m, n)
        apply_cos_sin(self.uy[0], -amp, m, n)

    def set_temperature(self, amp, m, n):
        apply_cos_sin(self.temp[0], -amp, m, n)

    def reset_time(self):
        self.time = 0.0
        self.navier.time = 0.0

    # ---- convection, :547-630 ----
    def _finish(self, conv):
        self.field.v = conv
        self.field.forward()
        if self.dealias:
            dealias(self.field)
        return self.field.vhat.copy()

    def conv_ux(self, ux, uy, t):
        sc = self.scale
        conv = conv_term(self.ux[1], self.field, ux, [1, 0], sc)
        conv = conv + conv_term(self.ux[1], self.field, uy, [0, 1], sc)
        conv = conv + conv_term(self.ux[1], self.field, ux, [1, 0], sc)
        conv = conv + conv_term(self.uy[1], self.field, uy, [1, 0], sc)
        conv = conv + conv_term(self.temp[1], self.field, t, [1, 0], sc)
        if self.fieldbc is not None:
            conv = conv + conv_term(self.temp[1], self.field, self.fieldbc.v, [1, 0], sc)
        return self._finish(conv)

    def conv_uy(self, ux, uy, t):
        sc = self.scale
        conv = conv_term(self.uy[1], self.field, ux, [1, 0], sc)
        conv = conv + conv_term(self.uy[1], self.field, uy, [0, 1], sc)
        conv = conv + conv_term(self.ux[1], self.field, ux, [0, 1], sc)
        conv = conv + conv_term(self.uy[1], self.field, uy, [0, 1], sc)
        conv = conv + conv_term(self.temp[1], self.field, t, [0, 1], sc)
        if self.fieldbc is not None:
            conv = conv + conv_term(self.temp[1], self.field, self.fieldbc.v, [0, 1], sc)
        return self._finish(conv)

    def conv_temp(self, ux, uy):
        conv = conv_term(self.temp[1], self.field, ux, [1, 0], self.scale)
        conv = conv + conv_term(self.temp[1], self.field, uy, [0, 1], self.scale)
        return self._finish(conv)

    # ---- solves, :632-697 ----
    def solve_ux(self, ux, uy, t):
        rhs = self.ux[0].to_ortho()
        rhs = rhs - self.pres[0].gradient([1, 0], self.scale) * self.dt
        rhs = rhs + self.conv_ux(ux, uy, t) * self.dt
        rhs = rhs + self.ux[1].gradient([2, 0], self.scale) * self.dt * self.nu
        rhs = rhs + self.ux[1].gradient([0, 2], self.scale) * self.dt * self.nu
        self.ux[0].from_ortho(rhs)

    def solve_uy(self, ux, uy, t):
        rhs = self.uy[0].to_ortho()
        rhs = rhs - self.pres[0].gradient([0, 1], self.scale) * self.dt
        rhs = rhs + self.conv_uy(ux, uy, t) * self.dt
        rhs = rhs + self.uy[1].gradient([2, 0], self.scale) * self.dt * self.nu
        rhs = rhs + self.uy[1].gradient([0, 2], self.scale) * self.dt * self.nu
        self.uy[0].from_ortho(rhs)

    def solve_temp(self, ux, uy):
        rhs = self.temp[0].to_ortho()
        rhs = rhs + self.conv_temp(ux, uy) * self.dt
        rhs = rhs + self.uy[1].to_ortho() * self.dt
        rhs = rhs + self.temp[1].gradient([2, 0], self.scale) * self.dt * self.ka
        rhs = rhs + self.temp[1].gradient([0, 2], self.scale) * self.dt * self.ka
        self.temp[0].from_ortho(rhs)

    # ---- projection, :699-737 ----
    def divergence(self):
        return self.ux[0].gradient([1, 0], self.scale) + self.uy[0].gradient([0, 1], self.scale)

    def solve_pres(self, f):
        self.pres[1].vhat = self.solver[0].solve(f, 0)
        self.pres[1].vhat[0, 0] = 0.0

    def project_velocity(self, c):
        dpdx = self.pres[1].gradient([1, 0], self.scale)
        dpdy = self.pres[1].gradient([0, 1], self.scale)
        old_ux, old_uy = self.ux[0].vhat.copy(), self.uy[0].vhat.copy()
        self.ux[0].from_ortho(dpdx)
        self.uy[0].from_ortho(dpdy)
        self.ux[0].vhat = self.ux[0].vhat * (-c) + old_ux
        self.uy[0].vhat = self.uy[0].vhat * (-c) + old_uy

    def update_pres(self, _div):
        self.pres[0].vhat = self.pres[0].vhat + self.pres[1].to_ortho() * (1.0 / self.dt)

    # ---- :739-766 ----
    def update_residual(self):
        nv = self.navier
        nv.ux.vhat = self.ux[0].vhat.copy()
        nv.uy.vhat = self.uy[0].vhat.copy()
        nv.temp.vhat = self.temp[0].vhat.copy()
        nv.update()
        nv.ux.vhat = (nv.ux.vhat - self.ux[0].vhat) / nv.dt
        nv.uy.vhat = (nv.uy.vhat - self.uy[0].vhat) / nv.dt
        nv.temp.vhat = (nv.temp.vhat - self.temp[0].vhat) / nv.dt
        self.fields_unsmoothed[0] = nv.ux.to_ortho()
        self.fields_unsmoothed[1] = nv.uy.to_ortho()
        self.fields_unsmoothed[2] = nv.temp.to_ortho()
        self.ux[1].vhat = self.smoother[0].solve(self.fields_unsmoothed[0], 0) * -1.0
        self.uy[1].vhat = self.smoother[1].solve(self.fields_unsmoothed[1], 0) * -1.0
        self.temp[1].vhat = self.smoother[2].solve(self.fields_unsmoothed[2], 0) * -1.0

    # ---- Integrate, :778-913 ----
    def update(self):
        self.ux[0].backward()
        self.uy[0].backward()
        self.temp[0].backward()
        ux, uy, temp = self.ux[0].v.copy(), self.uy[0].v.copy(), self.temp[0].v.copy()
        self.update_residual()
        self.solve_ux(ux, uy, temp)
        self.solve_uy(ux, uy, temp)
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

    def residuals(self):
        """(|ux res|, |uy res|, |temp res|) smoothed (what exit() tests) and unsmoothed (callback, :868-876)."""
        sm = [norm_l2(f[1].vhat) for f in (self.ux, self.uy, self.temp)]
        un = [norm_l2(a) for a in self.fields_unsmoothed]
        return sm, un

    def div_norm(self):
        return norm_l2(self.divergence())

    def exit(self):  # :892-910
        if math.isnan(self.div_norm()):
            return True
        return sum(self.residuals()[0]) < self.res_tol

    # ---- :952-992 (functions.rs through the adjoint's own fields) ----
    def _as_navier(self):
        n = object.__new__(Navier2D)
        n.temp, n.ux, n.uy, n.field, n.fieldbc = self.temp[0], self.ux[0], self.uy[0], self.field, self.fieldbc
        n.scale, n.ka, n.nu = self.scale, self.ka, self.nu
        return n

    def eval_nu(self):
        return self._as_navier().eval_nu()

    def eval_nuvol(self):
        return self._as_navier().eval_nuvol()

    def eval_re(self):
        return self._as_navier().eval_re()
