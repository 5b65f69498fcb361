"""Oracle restatement of rustpde::field (src/field.rs, src/field/average.rs).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np

from .funspace import Chebyshev, CompositeChebyshev, FourierR2c


class Field2:
    """FieldBase<f64, f64, T2, S, 2>, src/field.rs:66-129."""

    def __init__(self, space):
        self.ndim = 2
        self.space = space
        self.v = space.ndarray_physical()
        self.vhat = space.ndarray_spectral()
        self.x = space.coords()
        self.dx = self._get_dx(self.x, [b.is_periodic for b in space.bases()])

    # src/field.rs:103-129
    def forward(self):
        self.vhat = self.space.forward(self.v)

    def backward(self):
        self.v = self.space.backward(self.vhat)

    def to_ortho(self):
        return self.space.to_ortho(self.vhat)

    def from_ortho(self, a):
        self.vhat = self.space.from_ortho(a)

    def gradient(self, deriv, scale=None):
        return self.space.gradient(self.vhat, deriv, scale)

    # src/field.rs:135-163
    @staticmethod
    def _get_dx(x_arr, is_periodic):
        out = []
        for x, per in zip(x_arr, is_periodic):
            if per:
                out.append(np.full(len(x), x[2] - x[1]))
            else:
                n = len(x)
                dx = np.zeros(n)
                for i in range(n):
                    xl = x[0] if i == 0 else (x[i] + x[i - 1]) / 2.0
                    xr = x[n - 1] if i == n - 1 else (x[i + 1] + x[i]) / 2.0
                    dx[i] = xr - xl
                out.append(dx)
        return out

    # src/field/average.rs:25-57
    def average_axis(self, axis):
        length = abs(self.x[axis][-1] - self.x[axis][0])
        dx = self.dx[axis]
        if axis == 0:
            w = self.v * dx[: self.v.shape[0], None] / length
        else:
            w = self.v * dx[None, : self.v.shape[1]] / length
        return w.sum(axis=axis)

    def average(self):
        length = abs(self.x[1][-1] - self.x[1][0])
        avg_x = self.average_axis(0) * self.dx[1] / length
        return float(avg_x.sum())

    # src/field.rs:194-225
    def ingredients_for_hholtz(self, axis, dense=None):
        """Returns (mat_a, mat_b, precond).  Literal dense construction."""
        b = self.space.bases()[axis]
        mass = b.mass()
        peye = b.laplace_inv_eye()
        pinv = peye @ b.laplace_inv()
        if isinstance(b, Chebyshev):
            ms = mass[:, 2:]
            mat_a, mat_b = pinv @ ms, peye @ ms
            precond = pinv
        elif isinstance(b, CompositeChebyshev):
            mat_a, mat_b = pinv @ mass, peye @ mass
            precond = pinv
        elif isinstance(b, FourierR2c):
            mat_a, mat_b = mass, b.laplace()
            precond = None
        else:
            raise TypeError(b)
        return mat_a, mat_b, precond

    # src/field.rs:234-252
    def ingredients_for_poisson(self, axis):
        mat_a, mat_b, precond = self.ingredients_for_hholtz(axis)
        is_diag = isinstance(self.space.bases()[axis], FourierR2c)
        return mat_a, mat_b, precond, is_diag
