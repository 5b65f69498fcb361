"""Host-side mirror of the reference's public API on the Navier2D path, over
the C ABI of librustpde_b200.so (same names, argument meaning and error
behaviour as rustpde; citations are reference file:line).

    funspace:  chebyshev, cheb_dirichlet, cheb_neumann, fourier_r2c, Space2
    field:     Field2 (v, vhat, x, dx, forward, backward, to_ortho, from_ortho, gradient)
    solver:    Hholtz (new / new2), HholtzAdi, Poisson  -- .solve(input, output, axis)
    navier:    Navier2D.new / new_periodic, set_velocity, set_temperature,
               update, callback, exit, eval_nu / eval_nuvol / eval_re; integrate()

Every object takes an optional keyword `lib=`; the default is the CUDA
library (`_ffi.product_lib()`), which raises if it cannot be loaded.
"""
import ctypes as C
import math

import numpy as np

from . import _ffi
from ._ffi import RustpdeError  # noqa: F401

BASE_CHEBYSHEV, BASE_CHEB_DIRICHLET, BASE_CHEB_NEUMANN, BASE_CHEB_DIRICHLET_BC, BASE_CHEB_NEUMANN_BC, BASE_FOURIER_R2C = range(6)
MAX_TIMESTEP = 10_000_000  # src/lib.rs:132


def _dp(a):
    return a.ctypes.data_as(_ffi.c_double_p)


def _as_f64(a, cplx):
    a = np.ascontiguousarray(a, dtype=np.complex128 if cplx else np.float64)
    return a, a.view(np.float64).reshape(-1)


class Base:
    """One funspace base (funspace/src/lib.rs:230-345)."""

    def __init__(self, kind, n):
        self.kind, self.n = kind, n

    def len_phys(self):
        return self.n

    def len_spec(self):
        if self.kind == BASE_CHEBYSHEV:
            return self.n
        if self.kind in (BASE_CHEB_DIRICHLET, BASE_CHEB_NEUMANN):
            return self.n - 2
        if self.kind == BASE_FOURIER_R2C:
            return self.n // 2 + 1
        return 2


def chebyshev(n):
    return Base(BASE_CHEBYSHEV, n)


def cheb_dirichlet(n):
    return Base(BASE_CHEB_DIRICHLET, n)


def cheb_neumann(n):
    return Base(BASE_CHEB_NEUMANN, n)


def cheb_dirichlet_bc(n):
    return Base(BASE_CHEB_DIRICHLET_BC, n)


def cheb_neumann_bc(n):
    return Base(BASE_CHEB_NEUMANN_BC, n)


def fourier_r2c(n):
    return Base(BASE_FOURIER_R2C, n)


class Space2:
    """funspace/src/space2.rs:42-53."""

    def __init__(self, base0, base1):
        self.base0, self.base1 = base0, base1


class Field2:
    """src/field.rs:66-129.  `v` / `vhat` are host mirrors of the device arrays:
    reading downloads, assigning uploads."""

    def __init__(self, space=None, lib=None, _handle=None, _owner=None):
        self._lib = lib or _ffi.product_lib()
        self._owner = _owner
        if _handle is None:
            h = C.c_void_p()
            self._lib.call("rp_field_create", space.base0.kind, space.base0.n, space.base1.kind, space.base1.n, C.byref(h))
            self._h, self._owned = h, True
        else:
            self._h, self._owned = _handle, False
        ph, sp, ort, cx = (C.c_int * 2)(), (C.c_int * 2)(), (C.c_int * 2)(), C.c_int()
        self._lib.call("rp_field_shape", self._h, ph, sp, ort, C.byref(cx))
        self.shape_physical, self.shape_spectral, self.shape_ortho = tuple(ph), tuple(sp), tuple(ort)
        self.is_complex = bool(cx.value)
        self.ndim = 2
        self.space = space

    def __del__(self):
        try:
            if getattr(self, "_owned", False):
                self._lib.c.rp_field_destroy(self._h)
        except Exception:
            pass

    @property
    def spectral_dtype(self):
        return np.complex128 if self.is_complex else np.float64

    def _coords(self, fn):
        out = []
        for axis in range(2):
            # the Fourier grid may have n or (rarely) n+1 points: ask with the physical size first
            n = self.shape_physical[axis]
            for trial in (n, n + 1):
                a = np.zeros(trial)
                if self._lib.c.__getattr__(fn)(self._h, axis, _dp(a), trial) == 0:
                    out.append(a)
                    break
            else:
                raise RuntimeError("coords query failed")
        return out

    @property
    def x(self):
        return self._coords("rp_field_coords")

    @property
    def dx(self):
        return self._coords("rp_field_dx")

    @property
    def v(self):
        a = np.zeros(self.shape_physical)
        self._lib.call("rp_field_download_v", self._h, _dp(a), a.size)
        return a

    @v.setter
    def v(self, val):
        a, flat = _as_f64(val, False)
        if a.shape != self.shape_physical:
            raise RustpdeError(2, "v: shape mismatch %s vs %s" % (a.shape, self.shape_physical))
        self._lib.call("rp_field_upload_v", self._h, _dp(flat), flat.size)

    @property
    def vhat(self):
        a = np.zeros(self.shape_spectral, dtype=self.spectral_dtype)
        flat = a.view(np.float64).reshape(-1)
        self._lib.call("rp_field_download_vhat", self._h, _dp(flat), flat.size)
        return a

    @vhat.setter
    def vhat(self, val):
        a, flat = _as_f64(val, self.is_complex)
        if a.shape != self.shape_spectral:
            raise RustpdeError(2, "vhat: shape mismatch %s vs %s" % (a.shape, self.shape_spectral))
        self._lib.call("rp_field_upload_vhat", self._h, _dp(flat), flat.size)

    def vhat_rows(self, row0, nrows):
        """Rows [row0, row0+nrows) of vhat (the kx slab a rank owns in the slab decomposition)."""
        a = np.zeros((nrows, self.shape_spectral[1]), dtype=self.spectral_dtype)
        flat = a.view(np.float64).reshape(-1)
        self._lib.call("rp_field_download_vhat_rows", self._h, int(row0), int(nrows), _dp(flat), flat.size)
        return a

    def set_vhat_rows(self, row0, val):
        a, flat = _as_f64(val, self.is_complex)
        self._lib.call("rp_field_upload_vhat_rows", self._h, int(row0), int(a.shape[0]), _dp(flat), flat.size)

    def forward(self):  # field.rs:103-105
        self._lib.call("rp_field_forward", self._h)

    def backward(self):  # field.rs:108-110
        self._lib.call("rp_field_backward", self._h)

    def to_ortho(self):  # field.rs:113-115
        a = np.zeros(self.shape_ortho, dtype=self.spectral_dtype)
        flat = a.view(np.float64).reshape(-1)
        self._lib.call("rp_field_to_ortho", self._h, _dp(flat), flat.size)
        return a

    def from_ortho(self, inp):  # field.rs:118-123
        a, flat = _as_f64(inp, self.is_complex)
        if a.shape != self.shape_ortho:
            raise RustpdeError(2, "from_ortho: shape mismatch %s vs %s" % (a.shape, self.shape_ortho))
        self._lib.call("rp_field_from_ortho", self._h, _dp(flat), flat.size)

    def gradient(self, deriv, scale=None):  # field.rs:127-129
        a = np.zeros(self.shape_ortho, dtype=self.spectral_dtype)
        flat = a.view(np.float64).reshape(-1)
        sc = None if scale is None else _dp(np.asarray(scale, dtype=np.float64))
        self._lib.call("rp_field_gradient", self._h, int(deriv[0]), int(deriv[1]), sc, _dp(flat), flat.size)
        return a

    def average(self):  # average.rs:51-57
        out = C.c_double()
        self._lib.call("rp_field_average", self._h, C.byref(out))
        return out.value

    def average_axis(self, axis):  # average.rs:25-33
        a = np.zeros(self.shape_physical[1])
        self._lib.call("rp_field_average_axis", self._h, axis, _dp(a), a.size)
        return a


class _Solver:
    def __init__(self):
        self._h = C.c_void_p()

    def __del__(self):
        try:
            if self._h:
                self._lib.c.rp_solver_destroy(self._h)
        except Exception:
            pass

    def _shapes(self, field):
        b0, b1 = field.space.base0, field.space.base1
        four = b0.kind == BASE_FOURIER_R2C
        self.shape_in = ((b0.n // 2 + 1) if four else b0.n, b1.n)
        self.shape_out = ((b0.n // 2 + 1) if four else b0.n - 2, b1.n - 2)

    def solve(self, inp, output=None, axis=0):
        """Solve::solve (src/solver.rs:60-66); shape mismatch raises (reference: panic!)."""
        cplx = np.iscomplexobj(inp)
        a, flat = _as_f64(inp, cplx)
        if a.shape != self.shape_in:
            raise RustpdeError(2, "Dimension mismatch in solver input: %s vs %s" % (a.shape, self.shape_in))
        out = np.zeros(self.shape_out, dtype=a.dtype)
        oflat = out.view(np.float64).reshape(-1)
        self._lib.call("rp_solver_solve", self._h, _dp(flat), flat.size, _dp(oflat), oflat.size, int(cplx))
        if output is not None:
            output[...] = out
        return out

    def solve_resident(self, reps=1, complex_data=False):
        """Repeat the last solve `reps` times on the device-resident rhs (asynchronous; measurement aid)."""
        self._lib.call("rp_solver_solve_resident", self._h, int(reps), int(bool(complex_data)))

    def sync(self):
        self._lib.call("rp_solver_sync", self._h)

    def path_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._lib.call("rp_solver_path", self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"specialised": bool(a.value), "split_gemm": bool(b.value), "launches": c.value}

    def export_eig(self):
        m, has = C.c_int(), C.c_int()
        self._lib.call("rp_solver_eig_size", self._h, C.byref(m), C.byref(has))
        lam = np.zeros(m.value)
        q = np.zeros((m.value, m.value)) if has.value else None
        p = np.zeros((m.value, m.value)) if has.value else None
        self._lib.call("rp_solver_export_eig", self._h, _dp(lam), _dp(q) if has.value else None, _dp(p) if has.value else None)
        return lam, q, p


def _eig_args(eig):
    lam, q, p = (np.ascontiguousarray(e, dtype=np.float64) for e in eig)
    return lam, q, p


class Hholtz(_Solver):
    """src/solver/hholtz.rs:42 (new) / :81 (new2): (alpha I - c D2) vhat = A f."""

    def __init__(self, field, c, alpha=1.0, eig=None, lib=None):
        super().__init__()
        self._lib = lib or field._lib
        self._shapes(field)
        if eig is None:
            self._lib.call("rp_hholtz_create", field._h, float(c[0]), float(c[1]), float(alpha), C.byref(self._h))
        else:
            lam, q, p = _eig_args(eig)
            self._lib.call("rp_hholtz_create_with_eig", field._h, float(c[0]), float(c[1]), float(alpha), _dp(lam), _dp(q), _dp(p), C.byref(self._h))

    @classmethod
    def new2(cls, field, c, alpha, **kw):
        return cls(field, c, alpha, **kw)


class HholtzAdi(_Solver):
    """src/solver/hholtz_adi.rs:44."""

    def __init__(self, field, c, lib=None):
        super().__init__()
        self._lib = lib or field._lib
        self._shapes(field)
        self._lib.call("rp_hholtz_adi_create", field._h, float(c[0]), float(c[1]), C.byref(self._h))


class Poisson(_Solver):
    """src/solver/poisson.rs:50."""

    def __init__(self, field, c, eig=None, lib=None):
        super().__init__()
        self._lib = lib or field._lib
        self._shapes(field)
        if eig is None:
            self._lib.call("rp_poisson_create", field._h, float(c[0]), float(c[1]), C.byref(self._h))
        else:
            lam, q, p = _eig_args(eig)
            self._lib.call("rp_poisson_create_with_eig", field._h, float(c[0]), float(c[1]), _dp(lam), _dp(q), _dp(p), C.byref(self._h))


class Navier2D:
    """src/navier/navier.rs:153-195.  Use Navier2D.new(...) / Navier2D.new_periodic(...)."""

    def __init__(self):
        raise TypeError("use Navier2D.new(...) or Navier2D.new_periodic(...)")

    @classmethod
    def _make(cls, nx, ny, ra, pr, dt, aspect, adiabatic, periodic, eig, lib):
        s = object.__new__(cls)
        s._lib = lib or _ffi.product_lib()
        s._h = C.c_void_p()
        if eig is not None and not periodic:
            lam, q, p = _eig_args(eig)
            s._lib.call("rp_navier_create_with_eig", nx, ny, ra, pr, dt, aspect, int(adiabatic), _dp(lam), _dp(q), _dp(p), C.byref(s._h))
        else:
            s._lib.call("rp_navier_create", nx, ny, ra, pr, dt, aspect, int(adiabatic), int(periodic), C.byref(s._h))
        s.nx, s.ny, s.ra, s.pr, s.dt, s.periodic = nx, ny, ra, pr, dt, bool(periodic)
        nu, ka, sc = C.c_double(), C.c_double(), (C.c_double * 2)()
        s._lib.call("rp_navier_params", s._h, C.byref(nu), C.byref(ka), sc)
        s.nu, s.ka, s.scale = nu.value, ka.value, [sc[0], sc[1]]
        kx = fourier_r2c if periodic else None
        spaces = {
            0: Space2((kx or (cheb_neumann if adiabatic else cheb_dirichlet))(nx), cheb_dirichlet(ny)),
            1: Space2((kx or cheb_dirichlet)(nx), cheb_dirichlet(ny)),
            2: Space2((kx or cheb_dirichlet)(nx), cheb_dirichlet(ny)),
            3: Space2((kx or chebyshev)(nx), chebyshev(ny)),
            4: Space2((kx or cheb_neumann)(nx), cheb_neumann(ny)),
            5: Space2((kx or chebyshev)(nx), chebyshev(ny)),
        }
        flds = []
        for i in range(6):
            fh = C.c_void_p()
            s._lib.call("rp_navier_field", s._h, i, C.byref(fh))
            flds.append(Field2(spaces[i], lib=s._lib, _handle=fh, _owner=s))
        s.temp, s.ux, s.uy = flds[0], flds[1], flds[2]
        s.pres = [flds[3], flds[4]]
        s.field = flds[5]
        s.diagnostics = {"time": [], "Nu": [], "Nuvol": [], "Re": []}
        s.write_intervall = None
        s._solid = None
        s.statistics = None
        s._dealias = True
        return s

    @classmethod
    def new(cls, nx, ny, ra, pr, dt, aspect, adiabatic, eig=None, lib=None):  # navier.rs:219-307
        return cls._make(nx, ny, ra, pr, dt, aspect, adiabatic, False, eig, lib)

    @classmethod
    def new_periodic(cls, nx, ny, ra, pr, dt, aspect, lib=None):  # navier.rs:384-467
        return cls._make(nx, ny, ra, pr, dt, aspect, True, True, None, lib)

    def __del__(self):
        try:
            if self._h:
                self._lib.c.rp_navier_destroy(self._h)
        except Exception:
            pass

    @property
    def dealias(self):
        return self._dealias

    @dealias.setter
    def dealias(self, on):
        self._lib.call("rp_navier_set_dealias", self._h, int(bool(on)))
        self._dealias = bool(on)

    @property
    def time(self):
        t = C.c_double()
        self._lib.call("rp_navier_get_time", self._h, C.byref(t))
        return t.value

    def set_velocity(self, amp, m, n):  # navier.rs:927-930
        self._lib.call("rp_navier_set_velocity", self._h, float(amp), float(m), float(n))

    def set_temperature(self, amp, m, n):  # navier.rs:934-936
        self._lib.call("rp_navier_set_temperature", self._h, float(amp), float(m), float(n))

    def set_temp_bc_ortho(self, that_bc):  # set_temp_bc, navier.rs:517-519 (ortho coefficients)
        a, flat = _as_f64(that_bc, self.periodic)
        self._lib.call("rp_navier_set_tempbc_ortho", self._h, _dp(flat), flat.size)

    @property
    def solid(self):
        return self._solid

    @solid.setter
    def solid(self, mask_value):
        """navier.solid = Some([mask, value]) (navier.rs:191; generators: rustpde_b200.solid_masks)."""
        if mask_value is None:
            self._lib.call("rp_navier_set_solid", self._h, None, None, 0)
            self._solid = None
            return
        mask, value = (np.ascontiguousarray(a, dtype=np.float64) for a in mask_value)
        if mask.shape != (self.nx, self.ny) or value.shape != mask.shape:
            raise RustpdeError(2, "solid: mask / value must be [nx, ny]")
        self._lib.call("rp_navier_set_solid", self._h, _dp(mask.reshape(-1)), _dp(value.reshape(-1)), mask.size)
        self._solid = [mask, value]

    def new_work_field(self):
        """Field2::new(&navier.field.space) (statistics.rs:60-65)."""
        return Field2(self.field.space, lib=self._lib)

    def tempbc_ortho(self):
        """fieldbc.to_ortho(): the Rayleigh-Benard boundary field of navier.rs:314-332 / 474-492 (only T_1(y) is present)."""
        a = np.zeros(self.field.shape_spectral, dtype=self.field.spectral_dtype)
        a[0, 1] = -0.5 * self.nx if self.periodic else -0.5  # the r2c forward transform is unnormalised
        return a

    def reset_time(self):  # navier.rs:951-953
        self._lib.call("rp_navier_reset_time", self._h)

    def set_graph(self, on):
        self._lib.call("rp_navier_set_graph", self._h, int(bool(on)))

    # ---- Integrate (src/lib.rs:135-146, navier.rs:737-862) ----------------
    def update(self, nsteps=1):
        """Integrate::update; nsteps > 1 queues several steps without a host sync."""
        self._lib.call("rp_navier_update", self._h, int(nsteps))

    def sync(self):
        self._lib.call("rp_navier_sync", self._h)

    def stage_state(self, temp, ux, uy, pres):
        """Queue the upload of a complete state (the `vhat` arrays of temp, ux, uy, pres[0] as flat float64
        buffers: numpy arrays or (address, n_doubles) pairs of page-locked memory) on the copy stream; returns
        at once.  The buffers must stay alive until commit_staged()."""
        args = []
        keep = []
        for a in (temp, ux, uy, pres):
            if isinstance(a, tuple):
                ptr, n = a
            else:
                a = np.ascontiguousarray(a)
                keep.append(a)
                ptr, n = a.ctypes.data, a.view(np.float64).size
            args += [C.cast(ptr, _ffi.c_double_p), C.c_size_t(n)]
        self._staged_keep = keep
        self._lib.call("rp_navier_stage_state", self._h, *args)

    def commit_staged(self):
        """Make the compute stream wait for the staged upload and move it into place (see stage_state)."""
        self._lib.call("rp_navier_commit_staged", self._h)
        self._staged_keep = None

    def _state_args(self, bufs, keep):
        args = []
        for a in bufs:
            if isinstance(a, tuple):
                ptr, n = a
            else:
                assert a.flags["C_CONTIGUOUS"]
                keep.append(a)
                ptr, n = a.ctypes.data, a.view(np.float64).size
            args += [C.cast(ptr, _ffi.c_double_p), C.c_size_t(n)]
        return args

    def fetch_state(self, temp, ux, uy, pres):
        """Queue the download of the complete state into host buffers (numpy arrays or (address, n_doubles) pairs of
        page-locked memory) on a second copy stream; overlaps the following update()s.  fetch_wait() completes it."""
        keep = []
        self._lib.call("rp_navier_fetch_state", self._h, *self._state_args((temp, ux, uy, pres), keep))
        self._fetch_keep = keep

    def fetch_wait(self):
        self._lib.call("rp_navier_fetch_wait", self._h)
        self._fetch_keep = None

    def div_async(self):
        """Queue |div u|_2 of the current state without waiting for it (see div_poll)."""
        self._lib.call("rp_navier_div_async", self._h)

    def div_poll(self, wait=False):
        """Most recent |div u|_2 that has arrived from div_async(), or None."""
        v, ready = C.c_double(), C.c_int()
        self._lib.call("rp_navier_div_poll", self._h, int(bool(wait)), C.byref(v), C.byref(ready))
        return v.value if ready.value else None

    def get_time(self):
        return self.time

    def get_dt(self):
        return self.dt

    def eval(self, nu=True, nuvol=True, re=True, div=True, ekin=True):
        vals = [C.c_double() for _ in range(5)]
        want = [nu, nuvol, re, div, ekin]
        args = [C.byref(v) if w else None for v, w in zip(vals, want)]
        self._lib.call("rp_navier_eval", self._h, *args)
        return [v.value if w else None for v, w in zip(vals, want)]

    def eval_nu(self):  # navier.rs:890-893
        return self.eval(True, False, False, False, False)[0]

    def eval_nuvol(self):  # navier.rs:899-909
        return self.eval(False, True, False, False, False)[1]

    def eval_re(self):  # navier.rs:912-921
        return self.eval(False, False, True, False, False)[2]

    def eval_ekin(self):
        return self.eval(False, False, False, False, True)[4]

    def div_norm(self):
        return self.eval(False, False, False, True, False)[3]

    def write(self, filename):  # navier.rs:975-981: errors are printed and swallowed
        try:
            self._lib.call("rp_navier_write_snapshot", self._h, str(filename).encode())
            print(" ==> %r" % str(filename))
        except RustpdeError:
            print("Error while writing file %r." % str(filename))

    def read(self, filename):  # navier.rs:963-972
        self._lib.call("rp_navier_read_snapshot", self._h, str(filename).encode())
        print(" <== %r" % str(filename))

    def callback(self, data_dir="data"):  # navier.rs:775-853
        import os
        t = self.time
        os.makedirs(data_dir, exist_ok=True)
        # "data/flow{:0>8.2}.h5" in the reference; same stem, the RPSNAP1 container (rustpde_b200/snapshot.py)
        fname = os.path.join(data_dir, "flow%s.rpsnap" % ("%.2f" % t).rjust(8, "0"))
        dt_save = self.write_intervall
        if dt_save is None or (t % dt_save) < self.dt / 2.0 or (t % dt_save) > dt_save - self.dt / 2.0:
            self.write(fname)
        st = self.statistics
        if st is not None:  # navier.rs:795-815
            if (t % st.save_stat) < self.dt / 2.0 or (t % st.save_stat) > st.save_stat - self.dt / 2.0:
                st.update(self.temp.to_ortho() + self.tempbc_ortho(), self.ux.to_ortho(), self.uy.to_ortho(), t)
            if (t % st.write_stat) < self.dt / 2.0 or (t % st.write_stat) > st.write_stat - self.dt / 2.0:
                st.write(os.path.join(data_dir, "statistics.rpsnap"))
        nu, nuvol, re, div, _ = self.eval(True, True, True, True, False)
        print("time = %4.2f      |div| = %4.2e     Nu = %5.3e     Nuv = %5.3e    Re = %5.3e" % (t, div, nu, nuvol, re))
        with open(os.path.join(data_dir, "info.txt"), "a") as f:  # navier.rs:843-852
            f.write("%r %r %r %r\n" % (t, nu, nuvol, re))
        self.diagnostics["time"].append(t)
        self.diagnostics["Nu"].append(nu)
        self.diagnostics["Nuvol"].append(nuvol)
        self.diagnostics["Re"].append(re)

    def exit(self):  # navier.rs:855-862: stop when |div| is NaN
        return math.isnan(self.div_norm())

    def exit_async(self):
        """exit() without a device sync: queues |div| of this step and tests the newest value that has already
        arrived (so a NaN is seen one check late)."""
        self.div_async()
        d = self.div_poll(False)
        return d is not None and math.isnan(d)

    def export_eig(self):
        m = self.nx - 2
        lam, q, p = np.zeros(m), np.zeros((m, m)), np.zeros((m, m))
        if self.periodic:
            raise RustpdeError(1, "periodic Navier2D has no eigen set-up data")
        self._lib.call("rp_navier_export_eig", self._h, _dp(lam), _dp(q), _dp(p))
        return lam, q, p

    def kernel_path(self):
        """(specialised, split_gemm): which kernels serve update() (see rp_navier_kernel_path)."""
        a, b = C.c_int(), C.c_int()
        self._lib.call("rp_navier_kernel_path", self._h, C.byref(a), C.byref(b))
        return bool(a.value), bool(b.value)

    def launches_per_step(self):
        n = C.c_int()
        self._lib.call("rp_navier_launches_per_step", self._h, C.byref(n))
        return n.value

    def profile(self, reps=5):
        """Per-launch device time of update() (ms, mean of `reps` eager steps) with the
        launch's name and algorithmic bytes / flops.  Advances the solution by `reps` steps."""
        ms = np.zeros(128)
        nops = C.c_int()
        self._lib.call("rp_navier_profile", self._h, int(reps), _dp(ms), ms.size, C.byref(nops))
        out = []
        for i in range(nops.value):
            name = C.create_string_buffer(64)
            by, fl = C.c_double(), C.c_double()
            self._lib.call("rp_navier_op_info", self._h, i, name, 64, C.byref(by), C.byref(fl))
            out.append({"name": name.value.decode(), "ms": float(ms[i]), "bytes": by.value, "flops": fl.value})
        return out


class _BorrowedSolver(_Solver):
    def __init__(self, lib, handle):
        self._lib = lib
        self._h = handle

    def __del__(self):
        pass


class Navier2DAdjoint:
    """src/navier/navier_adjoint.rs:128-176.  Use Navier2DAdjoint.new(...) / .new_periodic(...)."""

    def __init__(self):
        raise TypeError("use Navier2DAdjoint.new(...) or Navier2DAdjoint.new_periodic(...)")

    @classmethod
    def _make(cls, nx, ny, ra, pr, dt, aspect, adiabatic, periodic, lib):
        s = object.__new__(cls)
        s._lib = lib or _ffi.product_lib()
        s._h = C.c_void_p()
        s._lib.call("rp_adjoint_create", nx, ny, ra, pr, dt, aspect, int(adiabatic), int(periodic), C.byref(s._h))
        s.nx, s.ny, s.ra, s.pr, s.dt, s.periodic = nx, ny, ra, pr, dt, bool(periodic)
        s.dt_navier = 1e-2
        kx = fourier_r2c if periodic else None
        sp_u = Space2((kx or cheb_dirichlet)(nx), cheb_dirichlet(ny))
        sp_t = Space2((kx or (cheb_neumann if adiabatic else cheb_dirichlet))(nx), cheb_dirichlet(ny))
        spaces = [sp_t, sp_u, sp_u, Space2((kx or chebyshev)(nx), chebyshev(ny)), Space2((kx or cheb_neumann)(nx), cheb_neumann(ny)), sp_t, sp_u, sp_u]
        flds = []
        for i in range(8):
            fh = C.c_void_p()
            s._lib.call("rp_adjoint_field", s._h, i, C.byref(fh))
            flds.append(Field2(spaces[i], lib=s._lib, _handle=fh, _owner=s))
        # [adjoint field, Navier-Stokes residual]
        s.temp, s.ux, s.uy = [flds[0], flds[5]], [flds[1], flds[6]], [flds[2], flds[7]]
        s.pres = [flds[3], flds[4]]
        s.diagnostics = {"time": [], "Nu": [], "Nuvol": [], "Re": []}
        s.write_intervall = None
        return s

    @classmethod
    def new(cls, nx, ny, ra, pr, dt, aspect, adiabatic, lib=None):  # navier_adjoint.rs:197
        return cls._make(nx, ny, ra, pr, dt, aspect, adiabatic, False, lib)

    @classmethod
    def new_periodic(cls, nx, ny, ra, pr, dt, aspect, lib=None):  # navier_adjoint.rs:361
        return cls._make(nx, ny, ra, pr, dt, aspect, True, True, lib)

    def __del__(self):
        try:
            if self._h:
                self._lib.c.rp_adjoint_destroy(self._h)
        except Exception:
            pass

    def set_velocity(self, amp, m, n):
        self._lib.call("rp_adjoint_set_velocity", self._h, float(amp), float(m), float(n))

    def set_temperature(self, amp, m, n):
        self._lib.call("rp_adjoint_set_temperature", self._h, float(amp), float(m), float(n))

    def reset_time(self):
        self._lib.call("rp_adjoint_reset_time", self._h)

    @property
    def time(self):
        t = C.c_double()
        self._lib.call("rp_adjoint_get_time", self._h, C.byref(t))
        return t.value

    def update(self, nsteps=1):
        self._lib.call("rp_adjoint_update", self._h, int(nsteps))

    def get_time(self):
        return self.time

    def get_dt(self):
        return self.dt

    def eval(self):
        v = [C.c_double() for _ in range(4)]
        self._lib.call("rp_adjoint_eval", self._h, *[C.byref(x) for x in v])
        return [x.value for x in v]  # Nu, Nuvol, Re, |div|

    def eval_nu(self):
        return self.eval()[0]

    def eval_nuvol(self):
        return self.eval()[1]

    def eval_re(self):
        return self.eval()[2]

    def div_norm(self):
        return self.eval()[3]

    def residuals(self):
        sm, un = (C.c_double * 3)(), (C.c_double * 3)()
        self._lib.call("rp_adjoint_residuals", self._h, sm, un)
        return list(sm), list(un)

    def exit(self):  # navier_adjoint.rs:892-910
        stop = C.c_int()
        self._lib.call("rp_adjoint_exit", self._h, C.byref(stop))
        return bool(stop.value)

    def callback(self):  # navier_adjoint.rs:815-890 (diagnostics part)
        nu, nuvol, re, div = self.eval()
        t = self.time
        print("time = %4.2f      |div| = %4.2e     Nu = %5.3e     Nuv = %5.3e    Re = %5.3e" % (t, div, nu, nuvol, re))
        sm, _ = self.residuals()
        print("|U| = %10.2e\n|V| = %10.2e\n|T| = %10.2e" % tuple(sm))
        for k, v in (("time", t), ("Nu", nu), ("Nuvol", nuvol), ("Re", re)):
            self.diagnostics[k].append(v)

    def export_eig(self):
        """Eigen set-up data of the four fast-diagonalisation solvers (confined): smoother of ux / uy, smoother of temp,
        pressure Poisson, the inner Navier2D's pressure Poisson -- what the oracle consumes in the parity tests."""
        out = {}
        for i, key in enumerate(("smooth_u", "smooth_t", "pres", "navier_pres")):
            sh = C.c_void_p()
            self._lib.call("rp_adjoint_solver", self._h, i, C.byref(sh))
            out[key] = _BorrowedSolver(self._lib, sh).export_eig()
        return out


def integrate(pde, max_time, save_intervall=None, exit_every=1, async_exit=False):
    """src/lib.rs:155-187.  `exit_every` > 1 checks the NaN break criterion less often (the reference checks every
    step, which costs a device sync); `async_exit` uses exit_async() (no sync, NaN seen one check late)."""
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
            print("time limit reached: %r" % pde.get_time())
            break
        if timestep >= MAX_TIMESTEP:
            print("timestep limit reached: %r" % timestep)
            break
        if timestep % exit_every == 0 and (pde.exit_async() if async_exit else pde.exit()):
            print("break criteria triggered")
            break
    return timestep
