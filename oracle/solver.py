"""Oracle restatement of rustpde::solver (src/solver/*.rs).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Lane algorithms are vectorised over the batch axis exactly like funspace.py.
`closed_form_*` give the banded operator entries of SURVEY.md section 8(a'')
so that large sizes never need the reference's dense n x n matrices; the
tests assert they agree with the literal dense construction.
"""
import numpy as np
import scipy.linalg as sla

from .funspace import Chebyshev, CompositeChebyshev, FourierR2c


def _lane_first(a, axis):
    return np.moveaxis(a, axis, 0)


# src/solver/utils.rs:17-44
def diag(a, offset):
    n = a.shape[0]
    m = abs(offset)
    if offset >= 0:
        return np.array([a[i, i + m] for i in range(n - m)])
    return np.array([a[i + m, i] for i in range(n - m)])


# src/solver/utils.rs:66-106
def inv(a):
    return sla.inv(a)


def eig(a):
    """utils.rs:66-98: dgeev, keep real parts, sort descending, Q^-1 by LAPACK."""
    ev, evec = sla.eig(a)
    ev = ev.real.copy()
    evec = evec.real.copy()
    perm = np.argsort(ev, kind="stable")[::-1]
    ev = ev[perm]
    evec = evec[:, perm]
    return ev, evec, inv(evec)


# --------------------------------------------------------------------------
class Fdma:
    """src/solver/fdma.rs:11-175: banded, offsets -2, 0, 2, 4."""

    def __init__(self, low, dia, up1, up2, sweep=True):
        self.n = len(dia)
        self.low = np.array(low, dtype=np.float64, copy=True)
        self.dia = np.array(dia, dtype=np.float64, copy=True)
        self.up1 = np.array(up1, dtype=np.float64, copy=True)
        self.up2 = np.array(up2, dtype=np.float64, copy=True)
        self.sweeped = False
        if sweep:
            self.sweep()

    @classmethod
    def from_matrix(cls, a):  # :33-37
        return cls(diag(a, -2), diag(a, 0), diag(a, 2), diag(a, 4), True)

    @classmethod
    def from_matrix_raw(cls, a):  # :44-53
        return cls(diag(a, -2), diag(a, 0), diag(a, 2), diag(a, 4), False)

    def sweep(self):  # :73-82
        n = self.n
        low, dia, up1, up2 = self.low, self.dia, self.up1, self.up2
        for i in range(2, n):
            low[i - 2] /= dia[i - 2]
            dia[i] -= low[i - 2] * up1[i - 2]
            if i < n - 2:
                up1[i] -= low[i - 2] * up2[i - 2]
        self.sweeped = True

    def fdma(self, x):  # :101-118, x lane-first, in place
        n = self.n
        low, dia, up1, up2 = self.low, self.dia, self.up1, self.up2
        for i in range(2, n):
            x[i] = x[i] - x[i - 2] * low[i - 2]
        x[n - 1] = x[n - 1] / dia[n - 1]
        x[n - 2] = x[n - 2] / dia[n - 2]
        x[n - 3] = (x[n - 3] - x[n - 1] * up1[n - 3]) / dia[n - 3]
        x[n - 4] = (x[n - 4] - x[n - 2] * up1[n - 4]) / dia[n - 4]
        for i in range(n - 5, -1, -1):
            x[i] = (x[i] - x[i + 2] * up1[i] - x[i + 4] * up2[i]) / dia[i]

    def solve(self, inp, axis):  # :161-175
        assert self.sweeped, "Fdma: Forward sweep must be performed for solve! Abort."
        out = np.array(inp, copy=True)
        self.fdma(_lane_first(out, axis))
        return out


def fdma_solve_multi(low, dia, up1, up2, x):
    """Per-row Fdma (hholtz.rs:182-190 / fdma_tensor.rs:219-227): row i of x
    is solved with its own raw diagonals low[i], dia[i], up1[i], up2[i]
    (shape (M, n) / (M, n-2) / (M, n-4)); sweep + solve, vectorised over rows.
    x has shape (M, n) and is solved in place along axis 1."""
    low = low.copy()
    dia = dia.copy()
    up1 = up1.copy()
    n = dia.shape[1]
    for i in range(2, n):  # sweep, fdma.rs:73-82
        low[:, i - 2] /= dia[:, i - 2]
        dia[:, i] -= low[:, i - 2] * up1[:, i - 2]
        if i < n - 2:
            up1[:, i] -= low[:, i - 2] * up2[:, i - 2]
    for i in range(2, n):  # fdma.rs:101-118
        x[:, i] = x[:, i] - x[:, i - 2] * low[:, i - 2]
    x[:, n - 1] = x[:, n - 1] / dia[:, n - 1]
    x[:, n - 2] = x[:, n - 2] / dia[:, n - 2]
    x[:, n - 3] = (x[:, n - 3] - x[:, n - 1] * up1[:, n - 3]) / dia[:, n - 3]
    x[:, n - 4] = (x[:, n - 4] - x[:, n - 2] * up1[:, n - 4]) / dia[:, n - 4]
    for i in range(n - 5, -1, -1):
        x[:, i] = (x[:, i] - x[:, i + 2] * up1[:, i] - x[:, i + 4] * up2[:, i]) / dia[:, i]


# --------------------------------------------------------------------------
class MatVecFdma:
    """src/solver/matvec.rs:125-230: banded (m x n) matvec, offsets -2,0,2,4."""

    def __init__(self, a=None, diags=None):
        if diags is not None:
            self.m, self.n, self.low, self.dia, self.up1, self.up2 = diags
            return
        m, n = a.shape
        self.m, self.n = m, n
        self.low = np.zeros(m)
        self.dia = np.zeros(m)
        self.up1 = np.zeros(m)
        self.up2 = np.zeros(m)
        for i in range(m):  # :149-160
            self.dia[i] = a[i, i]
            if i > 1:
                self.low[i] = a[i, i - 2]
            if i < m - 2:
                self.up1[i] = a[i, i + 2]
            if i < m - 4:
                self.up2[i] = a[i, i + 4]

    def solve(self, inp, axis):  # :172-193, 212-230
        x = _lane_first(inp, axis)
        n = self.m
        out = np.zeros((n,) + x.shape[1:], dtype=inp.dtype)
        for i in range(n):
            o = x[i] * self.dia[i]
            if i > 1:
                o = o + x[i - 2] * self.low[i]
            if i < n - 2:
                o = o + x[i + 2] * self.up1[i]
            if i < n - 4:
                o = o + x[i + 4] * self.up2[i]
            out[i] = o
        return np.ascontiguousarray(np.moveaxis(out, 0, axis))


# --------------------------------------------------------------------------
# Closed-form banded operator entries (SURVEY.md 8a'') -- fast path
# --------------------------------------------------------------------------
def _stencil_dl(base):
    """(d, l, ncols) of the composite stencil S; ortho Chebyshev = identity
    with the first two columns sliced off (field.rs:204-207)."""
    n = base.n
    if isinstance(base, Chebyshev):
        # columns c=0..n-3 <-> parent index c+2 : S[c+2, c] = 1 -> as "d at row c+2"
        return None
    st = base.stencil
    return st.diag, st.low2


def _b2(n):
    """B2 = _pinv(n,2) band entries as functions of row i (ortho.rs:160-171)."""
    lo = np.zeros(n)  # B2[i, i-2]
    di = np.zeros(n)  # B2[i, i]
    up = np.zeros(n)  # B2[i, i+2]
    lo[2] = 0.25
    for i in range(3, n):
        lo[i] = 1.0 / float(4 * i * (i - 1))
    for i in range(2, n - 2):
        di[i] = -1.0 / float(2 * (i * i - 1))
    for i in range(2, n - 4):
        up[i] = 1.0 / float(4 * i * (i + 1))
    return lo, di, up


def closed_form_precond(n):
    """MatVecFdma of the (n-2) x n preconditioner pinv = peye . B2."""
    lo, di, up = _b2(n)
    m = n - 2
    low = np.zeros(m)
    dia = np.zeros(m)
    up1 = np.zeros(m)
    up2 = np.zeros(m)
    for r in range(m):
        i = r + 2
        dia[r] = lo[i]  # pinv[r, r]   = B2[i, i-2]
        if r > 1:
            low[r] = 0.0  # pinv[r, r-2] = B2[i, i-4] = 0
        if r < m - 2:
            up1[r] = di[i]  # pinv[r, r+2] = B2[i, i]
        if r < m - 4:
            up2[r] = up[i]  # pinv[r, r+4] = B2[i, i+2]
    return MatVecFdma(diags=(m, n, low, dia, up1, up2))


def closed_form_a_c(base):
    """Banded diagonals of mat_b-like A = I2.S (offsets 0,+2) and
    mat_a-like C = B2.S (offsets -2,0,+2,+4) for a Chebyshev-family base.
    Returns dict of 1-D arrays (low, dia, up1, up2) for A and C; size n-2."""
    n = base.n
    m = n - 2
    if isinstance(base, Chebyshev):
        # S[:, c] = e_{c+2}: mass[:, 2:]
        def s_entry(row, col):
            return 1.0 if row == col + 2 else 0.0
    else:
        d, l = base.stencil.diag, base.stencil.low2

        def s_entry(row, col):
            if col < 0 or col >= m:
                return 0.0
            if row == col:
                return d[col]
            if row == col + 2:
                return l[col]
            return 0.0

    lo, di, up = _b2(n)
    A = {k: np.zeros(sz) for k, sz in (("low", m - 2), ("dia", m), ("up1", m - 2), ("up2", m - 4))}
    C = {k: np.zeros(sz) for k, sz in (("low", m - 2), ("dia", m), ("up1", m - 2), ("up2", m - 4))}

    def a_entry(r, c):  # (I2 S)[r, c] = S[r+2, c]
        return s_entry(r + 2, c)

    def c_entry(r, c):  # (B2 S)[r, c] = sum_k B2[i,k] S[k,c], i=r+2, k in {i-2,i,i+2}
        i = r + 2
        tot = 0.0
        for k, b in ((i - 2, lo[i]), (i, di[i]), (i + 2, up[i] if i + 2 < n else 0.0)):
            if 0 <= k < n and b != 0.0:
                tot += b * s_entry(k, c)
        return tot

    for r in range(m):
        A["dia"][r] = a_entry(r, r)
        C["dia"][r] = c_entry(r, r)
        if r + 2 < m:
            A["up1"][r] = a_entry(r, r + 2)
            C["up1"][r] = c_entry(r, r + 2)
            A["low"][r] = a_entry(r + 2, r)
            C["low"][r] = c_entry(r + 2, r)
        if r + 4 < m:
            A["up2"][r] = a_entry(r, r + 4)
            C["up2"][r] = c_entry(r, r + 4)
    return A, C


def _dense_from_diags(dg, m):
    a = np.zeros((m, m))
    for r in range(m):
        a[r, r] = dg["dia"][r]
        if r + 2 < m:
            a[r, r + 2] = dg["up1"][r]
            a[r + 2, r] = dg["low"][r]
        if r + 4 < m:
            a[r, r + 4] = dg["up2"][r]
    return a


def ingredients_banded(field, axis):
    """Banded equivalent of Field2.ingredients_for_hholtz for one axis.
    Returns (diags_a, diags_b, MatVecFdma|None, lam|None) where for a
    Fourier axis lam = -k^2 (laplace diagonal) and diags are None."""
    b = field.space.bases()[axis]
    if isinstance(b, FourierR2c):
        return None, None, None, -(b.k.imag ** 2)
    A, C = closed_form_a_c(b)
    # mat_a = pinv.mass = C ; mat_b = peye.mass = A
    return C, A, closed_form_precond(b.n), None


def _fdma_from_diags(dg, scale=1.0, sweep=False):
    return Fdma(dg["low"] * scale, dg["dia"] * scale, dg["up1"] * scale, dg["up2"] * scale, sweep)


# --------------------------------------------------------------------------
class HholtzAdi:
    """src/solver/hholtz_adi.rs:32-130:  (I - c D2) vhat = A f, ADI."""

    def __init__(self, field, c, banded=False):
        self.solver = []
        self.matvec = []
        for axis, ci in enumerate(c):
            if banded:
                da, db, mv, lam = ingredients_banded(field, axis)
                if lam is not None:
                    n = len(lam)
                    self.solver.append(Fdma(np.zeros(n - 2), 1.0 - lam * ci, np.zeros(n - 2), np.zeros(n - 4)))
                else:
                    self.solver.append(
                        Fdma(
                            da["low"] - db["low"] * ci,
                            da["dia"] - db["dia"] * ci,
                            da["up1"] - db["up1"] * ci,
                            da["up2"] - db["up2"] * ci,
                        )
                    )
                self.matvec.append(mv)
            else:
                mat_a, mat_b, precond = field.ingredients_for_hholtz(axis)
                mat = mat_a - mat_b * ci  # :54
                self.solver.append(Fdma.from_matrix(mat))
                self.matvec.append(None if precond is None else MatVecFdma(precond))

    def solve(self, inp, axis=0):  # :98-130
        rhs = inp.copy() if self.matvec[0] is None else self.matvec[0].solve(inp, 0)
        if len(self.solver) == 1:
            return self.solver[0].solve(rhs, 0)
        if self.matvec[1] is not None:
            rhs = self.matvec[1].solve(rhs, 1)
        out = self.solver[0].solve(rhs, 0)
        return self.solver[1].solve(out, 1)


# --------------------------------------------------------------------------
class FdmaTensor:
    """src/solver/fdma_tensor.rs:73-234 (N = 1 or 2)."""

    def __init__(self, a, c, a_is_diag, alpha, eig_data=None, diags=None):
        """a, c: lists of dense matrices (or None when `diags` gives the
        outermost banded diagonals and `eig_data`/lam the inner axis)."""
        self.alpha = alpha
        ndim = len(a_is_diag)
        self.fwd, self.bwd, self.lam = [], [], []
        for i in range(ndim - 1):  # :117-128
            if a_is_diag[i]:
                self.lam.append(np.array(a[i], dtype=np.float64) if np.ndim(a[i]) == 1 else diag(a[i], 0))
                self.fwd.append(None)
                self.bwd.append(None)
            elif eig_data is not None:
                lam, q, p = eig_data
                self.lam.append(np.array(lam, copy=True))
                self.fwd.append(p)
                self.bwd.append(q)
            else:
                cinv = inv(c[i])
                xmat = cinv @ a[i]
                lam, q, qi = eig(xmat)
                self.lam.append(lam)
                self.fwd.append(qi @ cinv)
                self.bwd.append(q)
        if diags is not None:
            da, dc = diags
            self.fdma = [_fdma_from_diags(da), _fdma_from_diags(dc)]
        else:
            self.fdma = [Fdma.from_matrix_raw(a[-1]), Fdma.from_matrix_raw(c[-1])]
        self.n = self.fdma[0].n
        self.ndim = ndim
        if ndim == 1:  # :146-150
            self.fdma[0].sweep()

    def solve(self, inp, axis=0):
        if self.ndim == 1:
            return self.fdma[0].solve(inp, axis)
        # :195-234
        assert inp.shape[0] == len(self.lam[0]) and inp.shape[1] == self.n, "Dimension mismatch in Tensor!"
        out = self.fwd[0] @ inp if self.fwd[0] is not None else inp.copy()
        l = (self.lam[0] + self.alpha)[:, None]
        f0, f1 = self.fdma
        fdma_solve_multi(
            f0.low[None, :] + f1.low[None, :] * l,
            f0.dia[None, :] + f1.dia[None, :] * l,
            f0.up1[None, :] + f1.up1[None, :] * l,
            f0.up2[None, :] + f1.up2[None, :] * l,
            out,
        )
        if self.bwd[0] is not None:
            out = self.bwd[0] @ out
        return out


def _tensor_from_field(field, c, sign, alpha, banded, eig_data):
    """Shared constructor body of Hholtz::new/new2 (hholtz.rs:42-115) and
    Poisson::new (poisson.rs:50-91).  laplacian = sign * mat_b * ci."""
    ndim = len(c)
    lap, mass, mv, isd = [], [], [], []
    diags = None
    for axis, ci in enumerate(c):
        if banded:
            da, db, m, lam = ingredients_banded(field, axis)
            if lam is not None:  # Fourier: mat_a = I, mat_b = diag(-k^2)
                lap.append(sign * lam * ci)
                mass.append(None)
                isd.append(True)
            else:
                isd.append(False)
                if axis == ndim - 1:
                    diags = ({k: sign * v * ci for k, v in db.items()}, da)
                    lap.append(None)
                    mass.append(None)
                else:
                    mm = field.space.bases()[axis].n - 2
                    lap.append(sign * _dense_from_diags(db, mm) * ci)
                    mass.append(_dense_from_diags(da, mm))
            mv.append(m)
        else:
            mat_a, mat_b, precond, is_diag = field.ingredients_for_poisson(axis)
            lap.append(sign * mat_b * ci)
            mass.append(mat_a)
            mv.append(None if precond is None else MatVecFdma(precond))
            isd.append(is_diag)
    tensor = FdmaTensor(lap, mass, isd, alpha, eig_data=eig_data, diags=diags)
    return tensor, mv


class _TensorSolver:
    def solve(self, inp, axis=0):
        # hholtz.rs:156-197 / poisson.rs:131-149
        rhs = inp.copy() if self.matvec[0] is None else self.matvec[0].solve(inp, 0)
        if len(self.matvec) > 1 and self.matvec[1] is not None:
            rhs = self.matvec[1].solve(rhs, 1)
        return self.solver.solve(rhs, 0)


class Hholtz(_TensorSolver):
    """src/solver/hholtz.rs:29-197: (alpha I - c D2) vhat = A f, fast diag."""

    def __init__(self, field, c, alpha=1.0, banded=False, eig_data=None):
        self.solver, self.matvec = _tensor_from_field(field, c, -1.0, alpha, banded, eig_data)

    @classmethod
    def new2(cls, field, c, alpha, **kw):  # hholtz.rs:81-115
        return cls(field, c, alpha, **kw)


class Poisson(_TensorSolver):
    """src/solver/poisson.rs:32-149: c D2 vhat = A f, fast diag."""

    def __init__(self, field, c, banded=False, eig_data=None):
        self.solver, self.matvec = _tensor_from_field(field, c, 1.0, 0.0, banded, eig_data)
        # poisson.rs:80-83
        if len(c) == 2 and abs(self.solver.lam[0][0]) < 1e-10:
            self.solver.lam[0] = self.solver.lam[0] - 1e-10
