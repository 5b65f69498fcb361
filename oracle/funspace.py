"""Oracle restatement of the funspace crate (bases + Space2).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are paths relative
to the reference checkout, e.g. funspace/src/chebyshev/ortho.rs:337-360.

All lane algorithms are written "lane axis first": the array is viewed with
the transformed axis as axis 0 and every step of the reference's per-lane
loop is applied to whole slices, i.e. vectorised over the batch axis.  The
arithmetic per lane is the reference's, in the reference's order.
"""
import math

import numpy as np
import scipy.fft as sfft

_WORKERS = -1  # scipy.fft threads (lane-parallel, like rayon's *_par variants)


def _lane_first(a, axis):
    return np.moveaxis(a, axis, 0)


def _resized(a, n, axis, dtype=None):
    shape = list(a.shape)
    shape[axis] = n
    return np.zeros(shape, dtype=dtype or a.dtype)


# --------------------------------------------------------------------------
# Chebyshev (orthogonal)                     funspace/src/chebyshev/ortho.rs
# --------------------------------------------------------------------------
class Chebyshev:
    """ortho.rs:20-80.  n physical points == m = n spectral coefficients."""

    kind = "chebyshev"
    is_periodic = False
    is_complex = False

    def __init__(self, n):
        self.n = n
        self.m = n
        # ortho.rs:51-58: correct_dct = (-1)^i ; fwd *= 1/(n-1) ; bwd /= 2
        sign = np.array([(-1.0) ** i for i in range(n)])
        self.correct_dct_forward = sign * (1.0 / (n - 1))
        self.correct_dct_backward = sign / 2.0
        # ortho.rs:71-80 nodes of the second kind, ascending -1..1
        mm = float(n - 1)
        k = np.arange(n, dtype=np.float64)
        self.x = -np.sin(math.pi * (mm - 2.0 * k) / (2.0 * mm))

    # -- Basics (traits.rs) ------------------------------------------------
    def len_phys(self):
        return self.n

    def len_spec(self):
        return self.m

    def coords(self):
        return self.x

    def mass(self):
        return np.eye(self.n)

    # -- Transform: ortho.rs:337-360 (forward), 383-407 (backward) ---------
    def forward(self, v, axis):
        out = sfft.dct(v, type=1, axis=axis, workers=_WORKERS)  # nddct1
        o = _lane_first(out, axis)
        shp = (-1,) + (1,) * (o.ndim - 1)
        o *= self.correct_dct_forward.reshape(shp)
        o[0] *= 0.5
        o[self.n - 1] *= 0.5
        return out

    def backward(self, vhat, axis):
        buf = np.array(vhat, dtype=vhat.dtype, copy=True)
        b = _lane_first(buf, axis)
        shp = (-1,) + (1,) * (b.ndim - 1)
        b *= self.correct_dct_backward.reshape(shp)
        b[0] *= 2.0
        b[self.n - 1] *= 2.0
        return sfft.dct(buf, type=1, axis=axis, workers=_WORKERS)

    # -- Differentiate: ortho.rs:107-125 -----------------------------------
    def differentiate(self, data, n_times, axis):
        out = np.array(data, copy=True)
        d = _lane_first(out, axis)
        n = self.n
        for _ in range(n_times):
            d[0] = d[1]
            for i in range(1, n - 1):
                d[i] = 2.0 * float(i + 1) * d[i + 1]
            d[n - 1] = 0.0
            for i in range(n - 3, 0, -1):
                d[i] = d[i] + d[i + 2]
            d[0] = d[0] + d[2] / 2.0
        return out

    # -- FromOrtho: ortho.rs:517-560 (identity) ----------------------------
    def to_ortho(self, a, axis):
        return np.array(a, copy=True)

    def from_ortho(self, a, axis):
        return np.array(a, copy=True)

    # -- LaplacianInverse: ortho.rs:147-180, 484-515 ------------------------
    def laplace(self):
        return diffmat_chebyshev(self.n, 2)

    def laplace_inv(self):
        return cheb_pinv(self.n, 2)

    def laplace_inv_eye(self):
        return np.eye(self.n)[2:, :]


def diffmat_chebyshev(n, deriv):
    """dmsuite.rs:24-49."""
    d = np.zeros((n, n))
    if deriv == 1:
        for p in range(n):
            for q in range(p + 1, n):
                if (p + q) % 2 != 0:
                    d[p, q] = float(q * 2)
    elif deriv == 2:
        for p in range(n):
            for q in range(p + 2, n):
                if (p + q) % 2 == 0:
                    d[p, q] = float(q * (q * q - p * p))
    else:
        raise NotImplementedError
    d[0, :] *= 0.5
    return d


def cheb_pinv(n, deriv):
    """ortho.rs:147-174 (pseudo-inverse B1 / B2)."""
    p = np.zeros((n, n))
    if deriv == 1:
        p[1, 0] = 1.0
        for i in range(2, n):
            p[i, i - 1] = 1.0 / (2.0 * i)
        for i in range(1, n - 2):
            p[i, i + 1] = -1.0 / (2.0 * i)
    elif deriv == 2:
        p[2, 0] = 0.25
        for i in range(3, n):
            p[i, i - 2] = 1.0 / float(4 * i * (i - 1))
        for i in range(2, n - 2):
            p[i, i] = -1.0 / float(2 * (i * i - 1))
        for i in range(2, n - 4):
            p[i, i + 2] = 1.0 / float(4 * i * (i + 1))
    else:
        raise ValueError("pinv does only support deriv's 1 & 2")
    return p


# --------------------------------------------------------------------------
# Stencils                          funspace/src/chebyshev/composite_stencil.rs
# --------------------------------------------------------------------------
def tdma(a, b, c, d):
    """linalg.rs:14-57: tridiagonal solve, offsets -2, 0, +2.  d: lane-first."""
    n = d.shape[0]
    x = np.zeros_like(d)
    w = np.zeros(n - 2)
    g = np.zeros_like(d)
    w[0] = c[0] / b[0]
    g[0] = d[0] / b[0]
    if len(c) > 1:
        w[1] = c[1] / b[1]
    g[1] = d[1] / b[1]
    for i in range(2, n - 2):
        w[i] = c[i] / (b[i] - a[i - 2] * w[i - 2])
    for i in range(2, n):
        g[i] = (d[i] - g[i - 2] * a[i - 2]) / (b[i] - a[i - 2] * w[i - 2])
    x[n - 1] = g[n - 1]
    x[n - 2] = g[n - 2]
    for i in range(n - 2, 0, -1):
        x[i - 1] = g[i - 1] - x[i + 1] * w[i - 1]
    return x


class StencilChebyshev:
    """composite_stencil.rs:90-276: phi_k = d_k T_k + l_k T_{k+2}."""

    def __init__(self, n, kind):
        self.n = n
        self.m = n - 2
        m = self.m
        self.diag = np.ones(m)
        if kind == "dirichlet":  # :117-130
            self.low2 = -np.ones(m)
        elif kind == "neumann":  # :139-157
            k = np.arange(m, dtype=np.float64)
            self.low2 = -1.0 * (k ** 2) / ((k + 2.0) ** 2)
        else:
            raise ValueError(kind)
        # :160-171
        self.main = self.diag * self.diag + self.low2 * self.low2
        self.off = self.diag[2:] * self.low2[: m - 2]

    def to_array(self):  # :181-188
        s = np.zeros((self.n, self.m))
        for i in range(self.m):
            s[i, i] = self.diag[i]
            s[i + 2, i] = self.low2[i]
        return s

    def multiply(self, c):  # :207-229, c lane-first (m, ...)
        n = self.n
        p = np.zeros((n,) + c.shape[1:], dtype=c.dtype)
        p[0] = c[0] * self.diag[0]
        p[1] = c[1] * self.diag[1]
        for i in range(2, n - 2):
            p[i] = c[i] * self.diag[i] + c[i - 2] * self.low2[i - 2]
        p[n - 2] = c[n - 4] * self.low2[n - 4]
        p[n - 1] = c[n - 3] * self.low2[n - 3]
        return p

    def solve(self, p):  # :250-276, p lane-first (n, ...)
        m = self.m
        c = np.zeros((m,) + p.shape[1:], dtype=p.dtype)
        for i in range(m):
            c[i] = p[i] * self.diag[i] + p[i + 2] * self.low2[i]
        return tdma(self.off, self.main, self.off, c)


class StencilChebyshevBoundary:
    """composite_stencil.rs:279-406: two boundary-lifting functions."""

    def __init__(self, n, kind):
        self.n = n
        self.m = 2
        if kind == "dirichlet":  # :287-293
            self.t0 = np.array([0.5, 0.5])
            self.t1 = np.array([-0.5, 0.5])
        elif kind == "neumann":  # :302-309
            self.t0 = np.array([0.5, 0.5])
            self.t1 = np.array([-1.0 / 8.0, 1.0 / 8.0])
        else:
            raise ValueError(kind)

    def to_array(self):  # :318-326
        s = np.zeros((self.n, 2))
        s[0, 0], s[0, 1] = self.t0
        s[1, 0], s[1, 1] = self.t1
        return s

    def multiply(self, c):  # :343-361
        p = np.zeros((self.n,) + c.shape[1:], dtype=c.dtype)
        p[0] = c[0] * self.t0[0] + c[1] * self.t0[1]
        p[1] = c[0] * self.t1[0] + c[1] * self.t1[1]
        return p

    def solve(self, p):  # :382-405
        t0, t1 = self.t0, self.t1
        c0 = p[0] * t0[0] + p[1] * t1[0]
        c1 = p[0] * t0[1] + p[1] * t1[1]
        a = t0[0] * t0[0] + t1[0] * t1[0]
        b = t0[0] * t0[1] + t1[0] * t1[1]
        c = t0[1] * t0[0] + t1[1] * t1[0]
        d = t0[1] * t0[1] + t1[1] * t1[1]
        det = 1.0 / (a * d - b * c)
        out = np.zeros((2,) + p.shape[1:], dtype=p.dtype)
        out[0] = (c0 * d - c1 * b) * det
        out[1] = (c1 * a - c0 * c) * det
        return out


# --------------------------------------------------------------------------
# CompositeChebyshev                  funspace/src/chebyshev/composite.rs
# --------------------------------------------------------------------------
class CompositeChebyshev:
    """composite.rs:20-115.  kind in dirichlet|neumann|dirichlet_bc|neumann_bc."""

    is_periodic = False
    is_complex = False

    def __init__(self, n, kind):
        self.n = n
        self.kind = kind
        self.ortho = Chebyshev(n)
        if kind in ("dirichlet", "neumann"):
            self.stencil = StencilChebyshev(n, kind)
        elif kind == "dirichlet_bc":
            self.stencil = StencilChebyshevBoundary(n, "dirichlet")
        elif kind == "neumann_bc":
            self.stencil = StencilChebyshevBoundary(n, "neumann")
        else:
            raise ValueError(kind)
        self.m = self.stencil.m
        self.x = self.ortho.x

    def len_phys(self):
        return self.n

    def len_spec(self):
        return self.m

    def coords(self):
        return self.x

    def mass(self):  # composite.rs:338-340
        return self.stencil.to_array()

    # composite.rs:255-278 / 293-316
    def to_ortho(self, a, axis):
        out = self.stencil.multiply(_lane_first(a, axis))
        return np.ascontiguousarray(np.moveaxis(out, 0, axis))

    def from_ortho(self, a, axis):
        out = self.stencil.solve(_lane_first(a, axis))
        return np.ascontiguousarray(np.moveaxis(out, 0, axis))

    # composite.rs:465-506
    def forward(self, v, axis):
        return self.from_ortho(self.ortho.forward(v, axis), axis)

    def backward(self, vhat, axis):
        return self.ortho.backward(self.to_ortho(vhat, axis), axis)

    # composite.rs:556-570
    def differentiate(self, data, n_times, axis):
        return self.ortho.differentiate(self.to_ortho(data, axis), n_times, axis)

    def laplace(self):
        return self.ortho.laplace()

    def laplace_inv(self):
        return self.ortho.laplace_inv()

    def laplace_inv_eye(self):
        return self.ortho.laplace_inv_eye()


# --------------------------------------------------------------------------
# FourierR2c                              funspace/src/fourier/r2c.rs
# --------------------------------------------------------------------------
class FourierR2c:
    """r2c.rs:24-99.  n real points -> m = n/2+1 complex coefficients."""

    kind = "fourier_r2c"
    is_periodic = True
    is_complex = True

    def __init__(self, n):
        self.n = n
        self.m = n // 2 + 1
        # c2c.rs:63-66: Array1::range(0, 2pi, 2pi/n): len = ceil((end-start)/step)
        step = 2.0 * math.pi / float(n)
        cnt = int(math.ceil((2.0 * math.pi - 0.0) / step))
        self.x = 0.0 + step * np.arange(cnt, dtype=np.float64)
        # r2c.rs:58-70
        self.k = 1j * np.arange(self.m, dtype=np.float64)

    def len_phys(self):
        return self.n

    def len_spec(self):
        return self.m

    def coords(self):
        return self.x

    def mass(self):
        return np.eye(self.m)

    # r2c.rs:250-303 (ndfft_r2c / ndifft_r2c == numpy rfft / irfft)
    def forward(self, v, axis):
        return sfft.rfft(v, axis=axis, workers=_WORKERS)

    def backward(self, vhat, axis):
        return sfft.irfft(vhat, n=self.n, axis=axis, workers=_WORKERS)

    # r2c.rs:88-99
    def differentiate(self, data, n_times, axis):
        out = np.array(data, dtype=np.complex128, copy=True)
        d = _lane_first(out, axis)
        shp = (-1,) + (1,) * (d.ndim - 1)
        k = self.k.reshape(shp)
        for _ in range(n_times):
            d *= k
        return out

    def to_ortho(self, a, axis):
        return np.array(a, copy=True)

    def from_ortho(self, a, axis):
        return np.array(a, copy=True)

    # r2c.rs:372-394
    def laplace(self):
        return np.diag(-(self.k.imag ** 2))

    def laplace_inv(self):
        p = self.laplace()
        for i in range(1, self.m):
            p[i, i] = 1.0 / p[i, i]
        return p

    def laplace_inv_eye(self):
        return np.eye(self.m)[1:, :]


# constructors, funspace/src/lib.rs:230-345
def chebyshev(n):
    return Chebyshev(n)


def cheb_dirichlet(n):
    return CompositeChebyshev(n, "dirichlet")


def cheb_neumann(n):
    return CompositeChebyshev(n, "neumann")


def cheb_dirichlet_bc(n):
    return CompositeChebyshev(n, "dirichlet_bc")


def cheb_neumann_bc(n):
    return CompositeChebyshev(n, "neumann_bc")


def fourier_r2c(n):
    return FourierR2c(n)


# --------------------------------------------------------------------------
# Space2                                        funspace/src/space2.rs
# --------------------------------------------------------------------------
class Space2:
    """space2.rs:42-363: forward = y then x, backward = x then y,
    to/from_ortho and gradient = axis 0 then axis 1."""

    def __init__(self, base0, base1):
        self.base0 = base0
        self.base1 = base1

    def bases(self):
        return [self.base0, self.base1]

    def shape_physical(self):
        return (self.base0.len_phys(), self.base1.len_phys())

    def shape_spectral(self):
        return (self.base0.len_spec(), self.base1.len_spec())

    @property
    def spectral_dtype(self):
        return np.complex128 if self.base0.is_complex else np.float64

    def ndarray_physical(self):
        return np.zeros(self.shape_physical())

    def ndarray_spectral(self):
        return np.zeros(self.shape_spectral(), dtype=self.spectral_dtype)

    def coords(self):
        return [self.base0.coords().copy(), self.base1.coords().copy()]

    def forward(self, v):  # :323-333
        buf = self.base1.forward(v, 1)
        return self.base0.forward(buf, 0)

    def backward(self, vhat):  # :346-356
        buf = self.base0.backward(vhat, 0)
        return self.base1.backward(buf, 1)

    def to_ortho(self, a):  # :182-201
        return self.base1.to_ortho(self.base0.to_ortho(a, 0), 1)

    def from_ortho(self, a):  # :204-226
        return self.base1.from_ortho(self.base0.from_ortho(a, 0), 1)

    def gradient(self, a, deriv, scale=None):  # :247-264
        buf = self.base0.differentiate(a, deriv[0], 0)
        out = self.base1.differentiate(buf, deriv[1], 1)
        if scale is not None:
            sc = (scale[0] ** deriv[0]) * (scale[1] ** deriv[1])
            out = out / sc
        return out
