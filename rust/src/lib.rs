//! Thin Rust host layer over `librustpde_b200.so` (include/rustpde_b200.h).
//!
//! It keeps the reference's public surface on the Navier2D path with the same names and
//! signatures and forwards the arithmetic to the CUDA library:
//!
//! * funspace constructors `chebyshev`, `cheb_dirichlet`, `cheb_neumann`, `cheb_dirichlet_bc`,
//!   `cheb_neumann_bc`, `fourier_r2c` (funspace/src/lib.rs:230-345) and `Space2::new`
//!   (funspace/src/space2.rs:48),
//! * `Field2::{new, forward, backward, to_ortho, from_ortho, gradient, average, average_axis}`
//!   with pub `v`, `vhat`, `x`, `dx` (src/field.rs:67-129, src/field/average.rs:25-57),
//! * `Hholtz::{new, new2}`, `HholtzAdi::new`, `Poisson::new` and `trait Solve`
//!   (src/solver.rs:56-155, hholtz.rs:42,81, hholtz_adi.rs:44, poisson.rs:50),
//! * `Navier2D::{new, new_periodic, set_velocity, set_temperature, reset_time, eval_nu,
//!   eval_nuvol, eval_re, read, write}` with pub `temp, ux, uy, pres, field, nu, ka, ra, pr, time,
//!   dt, scale, diagnostics, dealias` (src/navier/navier.rs:153-195, 219-227, 384-391, 890-981),
//! * `trait Integrate` and `integrate` (src/lib.rs:135-187).
//!
//! Device residency: the arrays live in HBM.  The pub `v` / `vhat` members are HOST MIRRORS
//! (`ndarray::Array2`, row-major like the reference).  Every method that the reference defines on
//! host arrays keeps its meaning by syncing the mirror it reads to the device first and the mirror
//! it writes back afterwards (`forward()` = push `v`, transform, pull `vhat`); the `*_device`
//! variants skip the transfers for callers that stay on the GPU, and `push_*` / `pull_*` are the
//! explicit sync points.  `Navier2D::update()` does NOT touch the mirrors (it would cost 4 x 33 MB
//! per step at 2048 x 2049): call `sync_to_host()` before reading `navier.temp.vhat` and
//! `sync_to_device()` after writing it (`callback()` and `write()` do the former themselves).
//!
//! A non-zero status of the C ABI becomes `panic!`, the reference's error convention on this path
//! (funspace/src/utils.rs:49-76, src/solver/fdma_tensor.rs:201-209).
//!
//! NOT compiled in the development image (no Rust toolchain there); `tests/test_cabi.py` checks
//! that the `extern "C"` block below declares exactly the functions of the header, and the
//! executable verification goes through the same C ABI from C (tests/c_harness) and Python.
#![allow(non_camel_case_types)]
#![allow(clippy::too_many_arguments)]
use ndarray::{Array1, Array2, ArrayBase, Data, DataMut, Ix2};
use num_complex::Complex;
use std::collections::HashMap;
use std::os::raw::{c_char, c_double, c_int, c_uchar, c_void};

#[repr(C)]
pub struct rp_field_t {
    _private: [u8; 0],
}
#[repr(C)]
pub struct rp_solver_t {
    _private: [u8; 0],
}
#[repr(C)]
pub struct rp_navier_t {
    _private: [u8; 0],
}
#[repr(C)]
pub struct rp_adjoint_t {
    _private: [u8; 0],
}

// One line per function of include/rustpde_b200.h (checked by tests/test_cabi.py).
extern "C" {
    fn rp_init(device: c_int) -> c_int;
    fn rp_last_error() -> *const c_char;
    fn rp_version() -> c_int;
    fn rp_is_emulated() -> c_int;
    fn rp_set_lapack_library(path: *const c_char) -> c_int;
    fn rp_field_create(kind_x: c_int, nx: c_int, kind_y: c_int, ny: c_int, out: *mut *mut rp_field_t) -> c_int;
    fn rp_field_destroy(f: *mut rp_field_t) -> c_int;
    fn rp_field_shape(f: *mut rp_field_t, phys: *mut c_int, spec: *mut c_int, ortho: *mut c_int, is_complex: *mut c_int) -> c_int;
    fn rp_field_coords(f: *mut rp_field_t, axis: c_int, x: *mut c_double, len: usize) -> c_int;
    fn rp_field_dx(f: *mut rp_field_t, axis: c_int, dx: *mut c_double, len: usize) -> c_int;
    fn rp_field_upload_v(f: *mut rp_field_t, v: *const c_double, len: usize) -> c_int;
    fn rp_field_download_v(f: *mut rp_field_t, v: *mut c_double, len: usize) -> c_int;
    fn rp_field_upload_vhat(f: *mut rp_field_t, vhat: *const c_double, len: usize) -> c_int;
    fn rp_field_download_vhat(f: *mut rp_field_t, vhat: *mut c_double, len: usize) -> c_int;
    fn rp_field_upload_vhat_rows(f: *mut rp_field_t, row0: c_int, nrows: c_int, vhat: *const c_double, len: usize) -> c_int;
    fn rp_field_download_vhat_rows(f: *mut rp_field_t, row0: c_int, nrows: c_int, vhat: *mut c_double, len: usize) -> c_int;
    fn rp_field_forward(f: *mut rp_field_t) -> c_int;
    fn rp_field_backward(f: *mut rp_field_t) -> c_int;
    fn rp_field_to_ortho(f: *mut rp_field_t, out: *mut c_double, len: usize) -> c_int;
    fn rp_field_from_ortho(f: *mut rp_field_t, input: *const c_double, len: usize) -> c_int;
    fn rp_field_gradient(f: *mut rp_field_t, dx: c_int, dy: c_int, scale: *const c_double, out: *mut c_double, len: usize) -> c_int;
    fn rp_field_average(f: *mut rp_field_t, out: *mut c_double) -> c_int;
    fn rp_field_average_axis(f: *mut rp_field_t, axis: c_int, out: *mut c_double, len: usize) -> c_int;
    fn rp_hholtz_create(f: *mut rp_field_t, cx: c_double, cy: c_double, alpha: c_double, out: *mut *mut rp_solver_t) -> c_int;
    fn rp_hholtz_adi_create(f: *mut rp_field_t, cx: c_double, cy: c_double, out: *mut *mut rp_solver_t) -> c_int;
    fn rp_poisson_create(f: *mut rp_field_t, cx: c_double, cy: c_double, out: *mut *mut rp_solver_t) -> c_int;
    fn rp_hholtz_create_with_eig(f: *mut rp_field_t, cx: c_double, cy: c_double, alpha: c_double, lam: *const c_double,
                                 q: *const c_double, p: *const c_double, out: *mut *mut rp_solver_t) -> c_int;
    fn rp_poisson_create_with_eig(f: *mut rp_field_t, cx: c_double, cy: c_double, lam: *const c_double, q: *const c_double,
                                  p: *const c_double, out: *mut *mut rp_solver_t) -> c_int;
    fn rp_solver_eig_size(s: *mut rp_solver_t, m: *mut c_int, has_matrices: *mut c_int) -> c_int;
    fn rp_solver_export_eig(s: *mut rp_solver_t, lam: *mut c_double, q: *mut c_double, p: *mut c_double) -> c_int;
    fn rp_solver_solve(s: *mut rp_solver_t, input: *const c_double, in_len: usize, out: *mut c_double, out_len: usize,
                       is_complex: c_int) -> c_int;
    fn rp_solver_solve_resident(s: *mut rp_solver_t, reps: c_int, is_complex: c_int) -> c_int;
    fn rp_solver_sync(s: *mut rp_solver_t) -> c_int;
    fn rp_solver_path(s: *mut rp_solver_t, specialised: *mut c_int, split_gemm: *mut c_int, launches: *mut c_int) -> c_int;
    fn rp_solver_destroy(s: *mut rp_solver_t) -> c_int;
    fn rp_navier_create(nx: c_int, ny: c_int, ra: c_double, pr: c_double, dt: c_double, aspect: c_double, adiabatic: c_int,
                        periodic: c_int, out: *mut *mut rp_navier_t) -> c_int;
    fn rp_navier_create_with_eig(nx: c_int, ny: c_int, ra: c_double, pr: c_double, dt: c_double, aspect: c_double,
                                 adiabatic: c_int, lam: *const c_double, q: *const c_double, p: *const c_double,
                                 out: *mut *mut rp_navier_t) -> c_int;
    fn rp_navier_destroy(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_set_velocity(h: *mut rp_navier_t, amp: c_double, m: c_double, n: c_double) -> c_int;
    fn rp_navier_set_temperature(h: *mut rp_navier_t, amp: c_double, m: c_double, n: c_double) -> c_int;
    fn rp_navier_set_tempbc_ortho(h: *mut rp_navier_t, that_bc: *const c_double, len: usize) -> c_int;
    fn rp_navier_set_dealias(h: *mut rp_navier_t, on: c_int) -> c_int;
    fn rp_navier_set_solid(h: *mut rp_navier_t, mask: *const c_double, value: *const c_double, len: usize) -> c_int;
    fn rp_navier_update(h: *mut rp_navier_t, nsteps: c_int) -> c_int;
    fn rp_navier_sync(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_stage_state(h: *mut rp_navier_t, temp: *const c_double, n_temp: usize, ux: *const c_double, n_ux: usize,
                             uy: *const c_double, n_uy: usize, pres: *const c_double, n_pres: usize) -> c_int;
    fn rp_navier_commit_staged(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_fetch_state(h: *mut rp_navier_t, temp: *mut c_double, n_temp: usize, ux: *mut c_double, n_ux: usize,
                             uy: *mut c_double, n_uy: usize, pres: *mut c_double, n_pres: usize) -> c_int;
    fn rp_navier_fetch_wait(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_div_async(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_div_poll(h: *mut rp_navier_t, wait: c_int, div_norm: *mut c_double, ready: *mut c_int) -> c_int;
    fn rp_navier_get_time(h: *mut rp_navier_t, t: *mut c_double) -> c_int;
    fn rp_navier_get_dt(h: *mut rp_navier_t, dt: *mut c_double) -> c_int;
    fn rp_navier_reset_time(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_params(h: *mut rp_navier_t, nu: *mut c_double, ka: *mut c_double, scale: *mut c_double) -> c_int;
    fn rp_navier_eval(h: *mut rp_navier_t, nu: *mut c_double, nuvol: *mut c_double, re: *mut c_double, div: *mut c_double,
                      ekin: *mut c_double) -> c_int;
    fn rp_navier_field(h: *mut rp_navier_t, which: c_int, out: *mut *mut rp_field_t) -> c_int;
    fn rp_navier_export_eig(h: *mut rp_navier_t, lam: *mut c_double, q: *mut c_double, p: *mut c_double) -> c_int;
    fn rp_navier_launches_per_step(h: *mut rp_navier_t, n: *mut c_int) -> c_int;
    fn rp_navier_set_graph(h: *mut rp_navier_t, on: c_int) -> c_int;
    fn rp_navier_slab_phase1(h: *mut rp_navier_t, k0: c_int, mkl: c_int, out6: *const *mut c_double) -> c_int;
    fn rp_navier_slab_phase2(h: *mut rp_navier_t, j0: c_int, nyl: c_int, in6: *const *const c_double, work: *mut c_double,
                             out3: *const *mut c_double) -> c_int;
    fn rp_navier_slab_phase3(h: *mut rp_navier_t, k0: c_int, mkl: c_int, in3: *const *const c_double) -> c_int;
    fn rp_navier_slab_phase1_p2p(h: *mut rp_navier_t, k0: c_int, mkl: c_int, world: c_int, joff: *const c_int,
                                 peers: *const *mut c_double) -> c_int;
    fn rp_navier_slab_phase2_p2p(h: *mut rp_navier_t, j0: c_int, nyl: c_int, in6: *const *const c_double, work: *mut c_double,
                                 world: c_int, koff: *const c_int, peers: *const *mut c_double) -> c_int;
    fn rp_dev_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    fn rp_dev_free(p: *mut c_void) -> c_int;
    fn rp_ipc_export(p: *mut c_void, handle: *mut c_uchar) -> c_int;
    fn rp_ipc_open(handle: *const c_uchar, out: *mut *mut c_void) -> c_int;
    fn rp_ipc_close(p: *mut c_void) -> c_int;
    fn rp_navier_kernel_path(h: *mut rp_navier_t, specialised: *mut c_int, split_gemm: *mut c_int) -> c_int;
    fn rp_navier_profile(h: *mut rp_navier_t, reps: c_int, ms: *mut c_double, cap: usize, nops: *mut c_int) -> c_int;
    fn rp_navier_op_info(h: *mut rp_navier_t, i: c_int, name: *mut c_char, name_len: usize, bytes: *mut c_double,
                         flops: *mut c_double) -> c_int;
    fn rp_adjoint_create(nx: c_int, ny: c_int, ra: c_double, pr: c_double, dt: c_double, aspect: c_double, adiabatic: c_int,
                         periodic: c_int, out: *mut *mut rp_adjoint_t) -> c_int;
    fn rp_adjoint_destroy(h: *mut rp_adjoint_t) -> c_int;
    fn rp_adjoint_set_velocity(h: *mut rp_adjoint_t, amp: c_double, m: c_double, n: c_double) -> c_int;
    fn rp_adjoint_set_temperature(h: *mut rp_adjoint_t, amp: c_double, m: c_double, n: c_double) -> c_int;
    fn rp_adjoint_update(h: *mut rp_adjoint_t, nsteps: c_int) -> c_int;
    fn rp_adjoint_get_time(h: *mut rp_adjoint_t, time: *mut c_double) -> c_int;
    fn rp_adjoint_reset_time(h: *mut rp_adjoint_t) -> c_int;
    fn rp_adjoint_eval(h: *mut rp_adjoint_t, nu: *mut c_double, nuvol: *mut c_double, re: *mut c_double, div: *mut c_double) -> c_int;
    fn rp_adjoint_residuals(h: *mut rp_adjoint_t, smooth: *mut c_double, unsmooth: *mut c_double) -> c_int;
    fn rp_adjoint_exit(h: *mut rp_adjoint_t, stop: *mut c_int) -> c_int;
    fn rp_adjoint_field(h: *mut rp_adjoint_t, which: c_int, out: *mut *mut rp_field_t) -> c_int;
    fn rp_adjoint_solver(h: *mut rp_adjoint_t, which: c_int, out: *mut *mut rp_solver_t) -> c_int;
    fn rp_navier_write_snapshot(h: *mut rp_navier_t, path: *const c_char) -> c_int;
    fn rp_navier_read_snapshot(h: *mut rp_navier_t, path: *const c_char) -> c_int;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rp_last_error()) }.to_string_lossy().into_owned();
        panic!("rustpde_b200 error {}: {}", rc, msg);
    }
}

/// Selects the CUDA device (one process per GPU) and, optionally, the LAPACK provider used for the
/// solver set-up (`eig` / `inv`, src/solver/utils.rs:66-106).  Called implicitly with device 0.
pub fn init(device: i32, lapack: Option<&str>) {
    unsafe {
        if let Some(p) = lapack {
            let c = std::ffi::CString::new(p).unwrap();
            check(rp_set_lapack_library(c.as_ptr()));
        }
        check(rp_init(device as c_int));
    }
}
fn ensure_init() {
    use std::sync::Once;
    static START: Once = Once::new();
    START.call_once(|| {
        let lapack = std::env::var("RUSTPDE_B200_LAPACK").ok();
        init(std::env::var("LOCAL_RANK").ok().and_then(|s| s.parse().ok()).unwrap_or(0), lapack.as_deref());
    });
}
/// (library version, is the CPU emulation build)
pub fn library_info() -> (i32, bool) {
    unsafe { (rp_version() as i32, rp_is_emulated() != 0) }
}

// ------------------------------------------------------------------------------------------------
// funspace: bases and Space2 (funspace/src/lib.rs:230-345, space2.rs:42-53)
// ------------------------------------------------------------------------------------------------
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum BaseKind {
    Chebyshev = 0,
    ChebDirichlet = 1,
    ChebNeumann = 2,
    ChebDirichletBc = 3,
    ChebNeumannBc = 4,
    FourierR2c = 5,
}
/// One funspace base: a kind and its number of physical points.
#[derive(Clone, Copy, Debug)]
pub struct Base {
    pub kind: BaseKind,
    pub n: usize,
}
impl Base {
    pub fn len_phys(&self) -> usize {
        self.n
    }
    pub fn len_spec(&self) -> usize {
        match self.kind {
            BaseKind::Chebyshev => self.n,
            BaseKind::ChebDirichlet | BaseKind::ChebNeumann => self.n - 2,
            BaseKind::FourierR2c => self.n / 2 + 1,
            _ => 2,
        }
    }
}
pub fn chebyshev(n: usize) -> Base {
    Base { kind: BaseKind::Chebyshev, n }
}
pub fn cheb_dirichlet(n: usize) -> Base {
    Base { kind: BaseKind::ChebDirichlet, n }
}
pub fn cheb_neumann(n: usize) -> Base {
    Base { kind: BaseKind::ChebNeumann, n }
}
pub fn cheb_dirichlet_bc(n: usize) -> Base {
    Base { kind: BaseKind::ChebDirichletBc, n }
}
pub fn cheb_neumann_bc(n: usize) -> Base {
    Base { kind: BaseKind::ChebNeumannBc, n }
}
pub fn fourier_r2c(n: usize) -> Base {
    Base { kind: BaseKind::FourierR2c, n }
}
/// `Space2::new(&base0, &base1)` (space2.rs:48)
#[derive(Clone, Copy, Debug)]
pub struct Space2 {
    pub base0: Base,
    pub base1: Base,
}
impl Space2 {
    pub fn new(base0: &Base, base1: &Base) -> Self {
        Space2 { base0: *base0, base1: *base1 }
    }
}

// ------------------------------------------------------------------------------------------------
// Field2 (src/field.rs:66-129)
// ------------------------------------------------------------------------------------------------
/// Spectral scalar: `f64` (Chebyshev x Chebyshev, `Space2R2r`) or `Complex<f64>` (Fourier x Chebyshev, `Space2R2c`).
/// Complex arrays cross the ABI as interleaved (re, im) doubles, which is `Complex<f64>`'s layout.
pub trait SpectralScalar: Clone + num_traits::Zero + 'static {
    const WIDTH: usize;
    const IS_COMPLEX: bool;
}
impl SpectralScalar for f64 {
    const WIDTH: usize = 1;
    const IS_COMPLEX: bool = false;
}
impl SpectralScalar for Complex<f64> {
    const WIDTH: usize = 2;
    const IS_COMPLEX: bool = true;
}

pub struct Field2<T: SpectralScalar> {
    h: *mut rp_field_t,
    owned: bool,
    /// Number of dimensions
    pub ndim: usize,
    pub space: Space2,
    /// Host mirror of the field in physical space
    pub v: Array2<f64>,
    /// Host mirror of the field in spectral space
    pub vhat: Array2<T>,
    /// Grid coordinates
    pub x: [Array1<f64>; 2],
    /// Grid deltas
    pub dx: [Array1<f64>; 2],
    shape_ortho: (usize, usize),
}
pub type Field2Real = Field2<f64>;
pub type Field2Complex = Field2<Complex<f64>>;

impl<T: SpectralScalar> Field2<T> {
    /// `Field2::new(&space)` (field.rs:91-100)
    pub fn new(space: &Space2) -> Self {
        ensure_init();
        let mut h = std::ptr::null_mut();
        unsafe {
            check(rp_field_create(space.base0.kind as c_int, space.base0.n as c_int, space.base1.kind as c_int,
                                  space.base1.n as c_int, &mut h));
        }
        Self::from_handle(h, true, *space)
    }
    fn from_handle(h: *mut rp_field_t, owned: bool, space: Space2) -> Self {
        let (mut ph, mut sp, mut or, mut cx) = ([0 as c_int; 2], [0 as c_int; 2], [0 as c_int; 2], 0 as c_int);
        unsafe { check(rp_field_shape(h, ph.as_mut_ptr(), sp.as_mut_ptr(), or.as_mut_ptr(), &mut cx)) };
        assert_eq!(cx != 0, T::IS_COMPLEX, "spectral scalar type does not match the space (real vs complex)");
        let coords = |fun: unsafe extern "C" fn(*mut rp_field_t, c_int, *mut c_double, usize) -> c_int, axis: usize| {
            // the Fourier grid may have n or n + 1 points (c2c.rs:63-66): ask with the physical size first
            for trial in [ph[axis] as usize, ph[axis] as usize + 1] {
                let mut a = Array1::<f64>::zeros(trial);
                if unsafe { fun(h, axis as c_int, a.as_mut_ptr(), trial) } == 0 {
                    return a;
                }
            }
            panic!("coordinate query failed");
        };
        Field2 {
            h,
            owned,
            ndim: 2,
            space,
            v: Array2::zeros((ph[0] as usize, ph[1] as usize)),
            vhat: Array2::zeros((sp[0] as usize, sp[1] as usize)),
            x: [coords(rp_field_coords, 0), coords(rp_field_coords, 1)],
            dx: [coords(rp_field_dx, 0), coords(rp_field_dx, 1)],
            shape_ortho: (or[0] as usize, or[1] as usize),
        }
    }
    // ---- explicit sync points between the host mirrors and the device arrays ----
    pub fn push_v(&self) {
        let a = self.v.as_standard_layout();
        unsafe { check(rp_field_upload_v(self.h, a.as_ptr(), a.len())) }
    }
    pub fn pull_v(&mut self) {
        unsafe { check(rp_field_download_v(self.h, self.v.as_mut_ptr(), self.v.len())) }
    }
    pub fn push_vhat(&self) {
        let a = self.vhat.as_standard_layout();
        unsafe { check(rp_field_upload_vhat(self.h, a.as_ptr() as *const c_double, a.len() * T::WIDTH)) }
    }
    pub fn pull_vhat(&mut self) {
        unsafe { check(rp_field_download_vhat(self.h, self.vhat.as_mut_ptr() as *mut c_double, self.vhat.len() * T::WIDTH)) }
    }
    /// Rows `[row0, row0 + rows.nrows())` of the device `vhat` (the kx slab a rank owns in the slab decomposition).
    pub fn push_vhat_rows(&self, row0: usize, rows: &Array2<T>) {
        let a = rows.as_standard_layout();
        unsafe {
            check(rp_field_upload_vhat_rows(self.h, row0 as c_int, a.nrows() as c_int, a.as_ptr() as *const c_double,
                                            a.len() * T::WIDTH))
        }
    }
    pub fn pull_vhat_rows(&self, row0: usize, nrows: usize) -> Array2<T> {
        let mut a = Array2::<T>::zeros((nrows, self.vhat.ncols()));
        unsafe {
            check(rp_field_download_vhat_rows(self.h, row0 as c_int, nrows as c_int, a.as_mut_ptr() as *mut c_double,
                                              a.len() * T::WIDTH))
        }
        a
    }
    // ---- the reference's methods (host semantics) ----
    /// `forward()`: v -> vhat (field.rs:103-105)
    pub fn forward(&mut self) {
        self.push_v();
        self.forward_device();
        self.pull_vhat();
    }
    /// `backward()`: vhat -> v (field.rs:108-110)
    pub fn backward(&mut self) {
        self.push_vhat();
        self.backward_device();
        self.pull_v();
    }
    /// `to_ortho()` (field.rs:113-115)
    pub fn to_ortho(&self) -> Array2<T> {
        self.push_vhat();
        self.to_ortho_device()
    }
    /// `from_ortho(&input)` (field.rs:118-123)
    pub fn from_ortho<S: Data<Elem = T>>(&mut self, input: &ArrayBase<S, Ix2>) {
        assert_eq!(input.dim(), self.shape_ortho, "from_ortho: shape mismatch");
        let a = input.as_standard_layout();
        unsafe { check(rp_field_from_ortho(self.h, a.as_ptr() as *const c_double, a.len() * T::WIDTH)) };
        self.pull_vhat();
    }
    /// `gradient(deriv, scale)` (field.rs:127-129)
    pub fn gradient(&self, deriv: [usize; 2], scale: Option<[f64; 2]>) -> Array2<T> {
        self.push_vhat();
        self.gradient_device(deriv, scale)
    }
    /// `average()` of `v` (average.rs:51-57)
    pub fn average(&self) -> f64 {
        self.push_v();
        let mut out = 0.0;
        unsafe { check(rp_field_average(self.h, &mut out)) };
        out
    }
    /// `average_axis(axis)` of `v` (average.rs:25-33); axis 0 is the one on the Navier2D path
    pub fn average_axis(&self, axis: usize) -> Array1<f64> {
        self.push_v();
        let mut out = Array1::<f64>::zeros(self.v.ncols());
        unsafe { check(rp_field_average_axis(self.h, axis as c_int, out.as_mut_ptr(), out.len())) };
        out
    }
    // ---- device-resident variants: no mirror traffic ----
    pub fn forward_device(&mut self) {
        unsafe { check(rp_field_forward(self.h)) }
    }
    pub fn backward_device(&mut self) {
        unsafe { check(rp_field_backward(self.h)) }
    }
    pub fn to_ortho_device(&self) -> Array2<T> {
        let mut out = Array2::<T>::zeros(self.shape_ortho);
        unsafe { check(rp_field_to_ortho(self.h, out.as_mut_ptr() as *mut c_double, out.len() * T::WIDTH)) };
        out
    }
    pub fn gradient_device(&self, deriv: [usize; 2], scale: Option<[f64; 2]>) -> Array2<T> {
        let mut out = Array2::<T>::zeros(self.shape_ortho);
        let sc = scale.as_ref().map_or(std::ptr::null(), |s| s.as_ptr());
        unsafe {
            check(rp_field_gradient(self.h, deriv[0] as c_int, deriv[1] as c_int, sc, out.as_mut_ptr() as *mut c_double,
                                    out.len() * T::WIDTH))
        };
        out
    }
    pub(crate) fn handle(&self) -> *mut rp_field_t {
        self.h
    }
}
impl<T: SpectralScalar> Drop for Field2<T> {
    fn drop(&mut self) {
        if self.owned {
            unsafe { rp_field_destroy(self.h) };
        }
    }
}

// ------------------------------------------------------------------------------------------------
// solvers (src/solver.rs:56-155)
// ------------------------------------------------------------------------------------------------
/// `trait Solve<A, D>` (src/solver.rs:56-66); `axis` is ignored for the 2-D solvers, like in the reference.
pub trait Solve<A> {
    fn solve<S1: Data<Elem = A>, S2: Data<Elem = A> + DataMut>(&self, input: &ArrayBase<S1, Ix2>, output: &mut ArrayBase<S2, Ix2>,
                                                                axis: usize);
}
/// Eigen set-up data of the fast diagonalisation (lam, Q, P = Q^-1 Cx^-1), row-major m x m, lam unshifted.
pub struct EigData {
    pub lam: Array1<f64>,
    pub q: Array2<f64>,
    pub p: Array2<f64>,
}
pub struct SolverHandle {
    h: *mut rp_solver_t,
}
impl SolverHandle {
    fn solve_raw<A: SpectralScalar, S1: Data<Elem = A>, S2: Data<Elem = A> + DataMut>(&self, input: &ArrayBase<S1, Ix2>,
                                                                                        output: &mut ArrayBase<S2, Ix2>) {
        let a = input.as_standard_layout();
        let mut out = Array2::<A>::zeros(output.dim());
        unsafe {
            check(rp_solver_solve(self.h, a.as_ptr() as *const c_double, a.len() * A::WIDTH, out.as_mut_ptr() as *mut c_double,
                                  out.len() * A::WIDTH, A::IS_COMPLEX as c_int))
        };
        output.assign(&out);
    }
    /// (m, Some((lam, Q, P))) of a fast-diagonalisation solver with a Chebyshev x axis
    pub fn export_eig(&self) -> Option<EigData> {
        let (mut m, mut has) = (0 as c_int, 0 as c_int);
        unsafe { check(rp_solver_eig_size(self.h, &mut m, &mut has)) };
        if has == 0 {
            return None;
        }
        let m = m as usize;
        let mut e = EigData { lam: Array1::zeros(m), q: Array2::zeros((m, m)), p: Array2::zeros((m, m)) };
        unsafe { check(rp_solver_export_eig(self.h, e.lam.as_mut_ptr(), e.q.as_mut_ptr(), e.p.as_mut_ptr())) };
        Some(e)
    }
    /// (runs on the specialised kernels, parity-split GEMMs, launches per real-data solve)
    pub fn path(&self) -> (bool, bool, usize) {
        let (mut a, mut b, mut c) = (0 as c_int, 0 as c_int, 0 as c_int);
        unsafe { check(rp_solver_path(self.h, &mut a, &mut b, &mut c)) };
        (a != 0, b != 0, c as usize)
    }
    /// Measurement aid: repeat the last solve on the device-resident rhs.
    pub fn solve_resident(&self, reps: usize, complex_data: bool) {
        unsafe {
            check(rp_solver_solve_resident(self.h, reps as c_int, complex_data as c_int));
            check(rp_solver_sync(self.h));
        }
    }
}
impl Drop for SolverHandle {
    fn drop(&mut self) {
        unsafe { rp_solver_destroy(self.h) };
    }
}
macro_rules! impl_solve {
    ($name:ident) => {
        impl Solve<f64> for $name {
            fn solve<S1: Data<Elem = f64>, S2: Data<Elem = f64> + DataMut>(&self, input: &ArrayBase<S1, Ix2>,
                                                                            output: &mut ArrayBase<S2, Ix2>, _axis: usize) {
                self.0.solve_raw(input, output)
            }
        }
        impl Solve<Complex<f64>> for $name {
            fn solve<S1: Data<Elem = Complex<f64>>, S2: Data<Elem = Complex<f64>> + DataMut>(&self, input: &ArrayBase<S1, Ix2>,
                                                                                              output: &mut ArrayBase<S2, Ix2>,
                                                                                              _axis: usize) {
                self.0.solve_raw(input, output)
            }
        }
        impl std::ops::Deref for $name {
            type Target = SolverHandle;
            fn deref(&self) -> &SolverHandle {
                &self.0
            }
        }
    };
}
/// `Hholtz` (src/solver/hholtz.rs:29-197): (alpha I - c D2) vhat = A f by fast diagonalisation
pub struct Hholtz(SolverHandle);
impl Hholtz {
    /// `Hholtz::new(&field, c)` (hholtz.rs:42)
    pub fn new<T: SpectralScalar>(field: &Field2<T>, c: [f64; 2]) -> Self {
        Self::new2(field, c, 1.0)
    }
    /// `Hholtz::new2(&field, c, alpha)` (hholtz.rs:81)
    pub fn new2<T: SpectralScalar>(field: &Field2<T>, c: [f64; 2], alpha: f64) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { check(rp_hholtz_create(field.handle(), c[0], c[1], alpha, &mut h)) };
        Hholtz(SolverHandle { h })
    }
    /// Same with caller-supplied eigen set-up data (what makes <= 1e-10 parity of this solve well defined).
    pub fn with_eig<T: SpectralScalar>(field: &Field2<T>, c: [f64; 2], alpha: f64, eig: &EigData) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe {
            check(rp_hholtz_create_with_eig(field.handle(), c[0], c[1], alpha, eig.lam.as_ptr(), eig.q.as_ptr(), eig.p.as_ptr(), &mut h))
        };
        Hholtz(SolverHandle { h })
    }
}
impl_solve!(Hholtz);
/// `HholtzAdi` (src/solver/hholtz_adi.rs:32-130)
pub struct HholtzAdi(SolverHandle);
impl HholtzAdi {
    /// `HholtzAdi::new(&field, c)` (hholtz_adi.rs:44)
    pub fn new<T: SpectralScalar>(field: &Field2<T>, c: [f64; 2]) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { check(rp_hholtz_adi_create(field.handle(), c[0], c[1], &mut h)) };
        HholtzAdi(SolverHandle { h })
    }
}
impl_solve!(HholtzAdi);
/// `Poisson` (src/solver/poisson.rs:32-149)
pub struct Poisson(SolverHandle);
impl Poisson {
    /// `Poisson::new(&field, c)` (poisson.rs:50)
    pub fn new<T: SpectralScalar>(field: &Field2<T>, c: [f64; 2]) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { check(rp_poisson_create(field.handle(), c[0], c[1], &mut h)) };
        Poisson(SolverHandle { h })
    }
    pub fn with_eig<T: SpectralScalar>(field: &Field2<T>, c: [f64; 2], eig: &EigData) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { check(rp_poisson_create_with_eig(field.handle(), c[0], c[1], eig.lam.as_ptr(), eig.q.as_ptr(), eig.p.as_ptr(), &mut h)) };
        Poisson(SolverHandle { h })
    }
}
impl_solve!(Poisson);
/// `enum SolverField` (src/solver.rs:100-155)
pub enum SolverField {
    Hholtz(Hholtz),
    HholtzAdi(HholtzAdi),
    Poisson(Poisson),
}

// ------------------------------------------------------------------------------------------------
// Navier2D (src/navier/navier.rs:153-195)
// ------------------------------------------------------------------------------------------------
/// Integrate trait, identical to rustpde's (src/lib.rs:135-146).
pub trait Integrate {
    fn update(&mut self);
    fn get_time(&self) -> f64;
    fn get_dt(&self) -> f64;
    fn callback(&mut self);
    fn exit(&mut self) -> bool;
}

/// `Navier2D<T, S>`: `Navier2D<f64>` from `new` (confined), `Navier2D<Complex<f64>>` from `new_periodic`.
pub struct Navier2D<T: SpectralScalar> {
    h: *mut rp_navier_t,
    /// Field for temperature
    pub temp: Field2<T>,
    /// Horizontal velocity
    pub ux: Field2<T>,
    /// Vertical velocity
    pub uy: Field2<T>,
    /// Pressure [pres, pseudo pressure]
    pub pres: [Field2<T>; 2],
    /// Intermediate work field (ortho x ortho)
    pub field: Field2<T>,
    pub nu: f64,
    pub ka: f64,
    pub ra: f64,
    pub pr: f64,
    /// mirror of the device clock; refreshed by `get_time()` / `update()`
    pub time: f64,
    pub dt: f64,
    pub scale: [f64; 2],
    pub diagnostics: HashMap<String, Vec<f64>>,
    pub write_intervall: Option<f64>,
    dealias: bool,
    /// `integrate()` checks the NaN break criterion without a device sync (the value of the previous check is tested)
    pub async_exit: bool,
}
pub type Navier2DConfined = Navier2D<f64>;
pub type Navier2DPeriodic = Navier2D<Complex<f64>>;

impl Navier2D<f64> {
    /// `Navier2D::new(nx, ny, ra, pr, dt, aspect, adiabatic)` (navier.rs:219-227)
    pub fn new(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64, adiabatic: bool) -> Self {
        ensure_init();
        let mut h = std::ptr::null_mut();
        unsafe { check(rp_navier_create(nx as c_int, ny as c_int, ra, pr, dt, aspect, adiabatic as c_int, 0, &mut h)) };
        let tx = if adiabatic { cheb_neumann(nx) } else { cheb_dirichlet(nx) };
        Self::wrap(h, ra, pr, dt, [
            Space2::new(&tx, &cheb_dirichlet(ny)),
            Space2::new(&cheb_dirichlet(nx), &cheb_dirichlet(ny)),
            Space2::new(&cheb_dirichlet(nx), &cheb_dirichlet(ny)),
            Space2::new(&chebyshev(nx), &chebyshev(ny)),
            Space2::new(&cheb_neumann(nx), &cheb_neumann(ny)),
            Space2::new(&chebyshev(nx), &chebyshev(ny)),
        ])
    }
    /// Same with the pressure-Poisson eigen set-up data supplied by the caller.
    pub fn new_with_eig(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64, adiabatic: bool, eig: &EigData) -> Self {
        ensure_init();
        let mut h = std::ptr::null_mut();
        unsafe {
            check(rp_navier_create_with_eig(nx as c_int, ny as c_int, ra, pr, dt, aspect, adiabatic as c_int, eig.lam.as_ptr(),
                                            eig.q.as_ptr(), eig.p.as_ptr(), &mut h))
        };
        let tx = if adiabatic { cheb_neumann(nx) } else { cheb_dirichlet(nx) };
        Self::wrap(h, ra, pr, dt, [
            Space2::new(&tx, &cheb_dirichlet(ny)),
            Space2::new(&cheb_dirichlet(nx), &cheb_dirichlet(ny)),
            Space2::new(&cheb_dirichlet(nx), &cheb_dirichlet(ny)),
            Space2::new(&chebyshev(nx), &chebyshev(ny)),
            Space2::new(&cheb_neumann(nx), &cheb_neumann(ny)),
            Space2::new(&chebyshev(nx), &chebyshev(ny)),
        ])
    }
    /// (lam, Q, P) of the pressure Poisson solver
    pub fn export_eig(&self) -> EigData {
        let m = self.pres[1].vhat.nrows();
        let mut e = EigData { lam: Array1::zeros(m), q: Array2::zeros((m, m)), p: Array2::zeros((m, m)) };
        unsafe { check(rp_navier_export_eig(self.h, e.lam.as_mut_ptr(), e.q.as_mut_ptr(), e.p.as_mut_ptr())) };
        e
    }
}
impl Navier2D<Complex<f64>> {
    /// `Navier2D::new_periodic(nx, ny, ra, pr, dt, aspect)` (navier.rs:384-391)
    pub fn new_periodic(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64) -> Self {
        ensure_init();
        let mut h = std::ptr::null_mut();
        unsafe { check(rp_navier_create(nx as c_int, ny as c_int, ra, pr, dt, aspect, 1, 1, &mut h)) };
        let fx = fourier_r2c(nx);
        Self::wrap(h, ra, pr, dt, [
            Space2::new(&fx, &cheb_dirichlet(ny)),
            Space2::new(&fx, &cheb_dirichlet(ny)),
            Space2::new(&fx, &cheb_dirichlet(ny)),
            Space2::new(&fx, &chebyshev(ny)),
            Space2::new(&fx, &cheb_neumann(ny)),
            Space2::new(&fx, &chebyshev(ny)),
        ])
    }
}
impl<T: SpectralScalar> Navier2D<T> {
    fn wrap(h: *mut rp_navier_t, ra: f64, pr: f64, dt: f64, spaces: [Space2; 6]) -> Self {
        let view = |i: usize| {
            let mut f = std::ptr::null_mut();
            unsafe { check(rp_navier_field(h, i as c_int, &mut f)) };
            Field2::<T>::from_handle(f, false, spaces[i])
        };
        let (mut nu, mut ka, mut scale) = (0.0, 0.0, [0.0; 2]);
        unsafe { check(rp_navier_params(h, &mut nu, &mut ka, scale.as_mut_ptr())) };
        let mut diagnostics = HashMap::new();
        for k in ["time", "Nu", "Nuvol", "Re"] {
            diagnostics.insert(k.to_string(), Vec::new());
        }
        Navier2D {
            h,
            temp: view(0),
            ux: view(1),
            uy: view(2),
            pres: [view(3), view(4)],
            field: view(5),
            nu,
            ka,
            ra,
            pr,
            time: 0.0,
            dt,
            scale,
            diagnostics,
            write_intervall: None,
            dealias: true,
            async_exit: false,
        }
    }
    /// `set_velocity(amp, m, n)` (navier.rs:927-930); device side, mirrors untouched
    pub fn set_velocity(&mut self, amp: f64, m: f64, n: f64) {
        unsafe { check(rp_navier_set_velocity(self.h, amp, m, n)) }
    }
    /// `set_temperature(amp, m, n)` (navier.rs:934-936)
    pub fn set_temperature(&mut self, amp: f64, m: f64, n: f64) {
        unsafe { check(rp_navier_set_temperature(self.h, amp, m, n)) }
    }
    /// `set_temp_bc` (navier.rs:517-519) with the ortho coefficients of the boundary field
    pub fn set_temp_bc_ortho(&mut self, that_bc: &Array2<T>) {
        let a = that_bc.as_standard_layout();
        unsafe { check(rp_navier_set_tempbc_ortho(self.h, a.as_ptr() as *const c_double, a.len() * T::WIDTH)) }
    }
    /// pub `dealias` (navier.rs:188); must be set before the first `update()`
    pub fn set_dealias(&mut self, on: bool) {
        unsafe { check(rp_navier_set_dealias(self.h, on as c_int)) };
        self.dealias = on;
    }
    pub fn dealias(&self) -> bool {
        self.dealias
    }
    /// pub `solid = Some([mask, value])` (navier.rs:191): volume penalisation masks on the physical grid
    /// (solid_masks.rs:34-175); before the first `update()`
    pub fn set_solid(&mut self, solid: Option<&[Array2<f64>; 2]>) {
        match solid {
            Some(s) => {
                let (m, v) = (s[0].as_standard_layout(), s[1].as_standard_layout());
                unsafe { check(rp_navier_set_solid(self.h, m.as_ptr(), v.as_ptr(), m.len())) }
            }
            None => unsafe { check(rp_navier_set_solid(self.h, std::ptr::null(), std::ptr::null(), 0)) },
        }
    }
    /// `reset_time()` (navier.rs:951-953)
    pub fn reset_time(&mut self) {
        unsafe { check(rp_navier_reset_time(self.h)) };
        self.time = 0.0;
    }
    /// Several steps without returning to the host in between (one CUDA-graph launch per step).
    pub fn update_n(&mut self, nsteps: usize) {
        unsafe { check(rp_navier_update(self.h, nsteps as c_int)) };
        self.time = self.get_time_device();
    }
    pub fn sync(&self) {
        unsafe { check(rp_navier_sync(self.h)) }
    }
    fn get_time_device(&self) -> f64 {
        let mut t = 0.0;
        unsafe { check(rp_navier_get_time(self.h, &mut t)) };
        t
    }
    /// Download temp / ux / uy / pres `vhat` into the pub mirrors (asynchronously, on the second copy stream).
    pub fn sync_to_host(&mut self) {
        unsafe {
            check(rp_navier_fetch_state(self.h,
                                        self.temp.vhat.as_mut_ptr() as *mut c_double, self.temp.vhat.len() * T::WIDTH,
                                        self.ux.vhat.as_mut_ptr() as *mut c_double, self.ux.vhat.len() * T::WIDTH,
                                        self.uy.vhat.as_mut_ptr() as *mut c_double, self.uy.vhat.len() * T::WIDTH,
                                        self.pres[0].vhat.as_mut_ptr() as *mut c_double, self.pres[0].vhat.len() * T::WIDTH));
            check(rp_navier_fetch_wait(self.h));
        }
    }
    /// Upload the pub `vhat` mirrors of temp / ux / uy / pres[0] (what `read()` assigns, navier.rs:963-972).
    pub fn sync_to_device(&mut self) {
        assert!(self.temp.vhat.is_standard_layout() && self.ux.vhat.is_standard_layout() && self.uy.vhat.is_standard_layout()
                && self.pres[0].vhat.is_standard_layout());
        unsafe {
            check(rp_navier_stage_state(self.h,
                                        self.temp.vhat.as_ptr() as *const c_double, self.temp.vhat.len() * T::WIDTH,
                                        self.ux.vhat.as_ptr() as *const c_double, self.ux.vhat.len() * T::WIDTH,
                                        self.uy.vhat.as_ptr() as *const c_double, self.uy.vhat.len() * T::WIDTH,
                                        self.pres[0].vhat.as_ptr() as *const c_double, self.pres[0].vhat.len() * T::WIDTH));
            check(rp_navier_commit_staged(self.h));
            check(rp_navier_sync(self.h));
        }
    }
    fn eval(&mut self, want: [bool; 5]) -> [f64; 5] {
        let mut v = [0.0f64; 5];
        let p = |i: usize, v: &mut [f64; 5]| if want[i] { &mut v[i] as *mut f64 } else { std::ptr::null_mut() };
        let (a, b, c, d, e) = (p(0, &mut v), p(1, &mut v), p(2, &mut v), p(3, &mut v), p(4, &mut v));
        unsafe { check(rp_navier_eval(self.h, a, b, c, d, e)) };
        v
    }
    /// `eval_nu()` (navier.rs:890-893)
    pub fn eval_nu(&mut self) -> f64 {
        self.eval([true, false, false, false, false])[0]
    }
    /// `eval_nuvol()` (navier.rs:899-909)
    pub fn eval_nuvol(&mut self) -> f64 {
        self.eval([false, true, false, false, false])[1]
    }
    /// `eval_re()` (navier.rs:912-921)
    pub fn eval_re(&mut self) -> f64 {
        self.eval([false, false, true, false, false])[2]
    }
    /// `<(ux^2 + uy^2) / 2>` with the averaging weights of average.rs:51-57
    pub fn eval_ekin(&mut self) -> f64 {
        self.eval([false, false, false, false, true])[4]
    }
    /// `|div u|_2` (navier.rs:698-703, 855-862)
    pub fn div_norm(&mut self) -> f64 {
        self.eval([false, false, false, true, false])[3]
    }
    /// `write(filename)` (navier.rs:975-1013): snapshot of temp / ux / uy / pres (v, vhat), grids and scalars in the
    /// reference's group / dataset layout (see INTEGRATION.md for the container format)
    pub fn write(&mut self, filename: &str) {
        let c = std::ffi::CString::new(filename).unwrap();
        let rc = unsafe { rp_navier_write_snapshot(self.h, c.as_ptr()) };
        if rc != 0 {
            // I/O errors are printed and swallowed in the reference (navier.rs:975-981)
            println!("Error while writing file {:?}.", filename);
        } else {
            println!(" ==> {:?}", filename);
        }
    }
    /// `read(filename)` (navier.rs:963-972): vhat of temp / ux / uy / pres and the time; shapes that differ are
    /// truncated / zero-padded like field/read.rs:113-122
    pub fn read(&mut self, filename: &str) {
        let c = std::ffi::CString::new(filename).unwrap();
        unsafe { check(rp_navier_read_snapshot(self.h, c.as_ptr())) };
        self.time = self.get_time_device();
        println!(" <== {:?}", filename);
    }
    /// (specialised kernels, parity-split GEMMs) serve update()
    pub fn kernel_path(&self) -> (bool, bool) {
        let (mut a, mut b) = (0 as c_int, 0 as c_int);
        unsafe { check(rp_navier_kernel_path(self.h, &mut a, &mut b)) };
        (a != 0, b != 0)
    }
    pub fn launches_per_step(&self) -> usize {
        let mut n = 0 as c_int;
        unsafe { check(rp_navier_launches_per_step(self.h, &mut n)) };
        n as usize
    }
    pub fn set_graph(&mut self, on: bool) {
        unsafe { check(rp_navier_set_graph(self.h, on as c_int)) }
    }
    /// Per-launch device time (ms) of one update(), with name / algorithmic bytes / flops.  Advances the solution.
    pub fn profile(&mut self, reps: usize) -> Vec<(String, f64, f64, f64)> {
        let mut ms = vec![0.0f64; 128];
        let mut nops = 0 as c_int;
        unsafe { check(rp_navier_profile(self.h, reps as c_int, ms.as_mut_ptr(), ms.len(), &mut nops)) };
        (0..nops as usize)
            .map(|i| {
                let mut name = vec![0 as c_char; 64];
                let (mut by, mut fl) = (0.0, 0.0);
                unsafe { check(rp_navier_op_info(self.h, i as c_int, name.as_mut_ptr(), 64, &mut by, &mut fl)) };
                let s = unsafe { std::ffi::CStr::from_ptr(name.as_ptr()) }.to_string_lossy().into_owned();
                (s, ms[i], by, fl)
            })
            .collect()
    }
    /// Raw handle for the slab-decomposition driver (see `slab` below).
    pub fn raw(&self) -> *mut rp_navier_t {
        self.h
    }
}
impl<T: SpectralScalar> Drop for Navier2D<T> {
    fn drop(&mut self) {
        unsafe { rp_navier_destroy(self.h) };
    }
}

impl<T: SpectralScalar> Integrate for Navier2D<T> {
    /// navier.rs:737-765
    fn update(&mut self) {
        unsafe { check(rp_navier_update(self.h, 1)) };
        self.time += self.dt;
    }
    fn get_time(&self) -> f64 {
        self.get_time_device()
    }
    fn get_dt(&self) -> f64 {
        let mut t = 0.0;
        unsafe { check(rp_navier_get_dt(self.h, &mut t)) };
        t
    }
    /// navier.rs:775-853: snapshot file + Nu / Nuvol / Re, printed and stored in `diagnostics`
    fn callback(&mut self) {
        let t = self.get_time_device();
        std::fs::create_dir_all("data").ok();
        let fname = format!("data/flow{:0>8.2}.h5", t);
        self.write(&fname);
        let v = self.eval([true, true, true, true, false]);
        println!("time = {:4.2}      |div| = {:4.2e}     Nu = {:5.3e}     Nuv = {:5.3e}    Re = {:5.3e}", t, v[3], v[0], v[1], v[2]);
        for (k, x) in [("time", t), ("Nu", v[0]), ("Nuvol", v[1]), ("Re", v[2])] {
            self.diagnostics.get_mut(k).unwrap().push(x);
        }
    }
    /// navier.rs:855-862: stop when |div| is NaN.  Only |div| is evaluated (no Nu / Nuvol / Re transforms); with
    /// `async_exit` the request is queued and the value of the PREVIOUS request is tested, so the step loop never
    /// waits for the device.
    fn exit(&mut self) -> bool {
        if self.async_exit {
            let (mut d, mut ready) = (0.0, 0 as c_int);
            unsafe {
                check(rp_navier_div_async(self.h));
                check(rp_navier_div_poll(self.h, 0, &mut d, &mut ready));
            }
            return ready != 0 && d.is_nan();
        }
        self.div_norm().is_nan()
    }
}

/// `Navier2DAdjoint` (src/navier/navier_adjoint.rs:128-176): steady-state adjoint descent on the device.
/// `temp`, `ux`, `uy` are `[adjoint field, Navier-Stokes residual]` like in the reference.
pub struct Navier2DAdjoint<T: SpectralScalar> {
    h: *mut rp_adjoint_t,
    pub temp: [Field2<T>; 2],
    pub ux: [Field2<T>; 2],
    pub uy: [Field2<T>; 2],
    pub pres: [Field2<T>; 2],
    pub ra: f64,
    pub pr: f64,
    pub time: f64,
    pub dt: f64,
    pub dt_navier: f64,
    pub diagnostics: HashMap<String, Vec<f64>>,
}
impl<T: SpectralScalar> Navier2DAdjoint<T> {
    fn make(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64, adiabatic: bool, periodic: bool) -> Self {
        ensure_init();
        let mut h = std::ptr::null_mut();
        unsafe { check(rp_adjoint_create(nx as c_int, ny as c_int, ra, pr, dt, aspect, adiabatic as c_int, periodic as c_int, &mut h)) };
        let bx = |k: BaseKind| Base { kind: if periodic { BaseKind::FourierR2c } else { k }, n: nx };
        let sp_u = Space2::new(&bx(BaseKind::ChebDirichlet), &cheb_dirichlet(ny));
        let sp_t = Space2::new(&bx(if adiabatic { BaseKind::ChebNeumann } else { BaseKind::ChebDirichlet }), &cheb_dirichlet(ny));
        let spaces = [sp_t, sp_u, sp_u, Space2::new(&bx(BaseKind::Chebyshev), &chebyshev(ny)),
                      Space2::new(&bx(BaseKind::ChebNeumann), &cheb_neumann(ny)), sp_t, sp_u, sp_u];
        let view = |i: usize| {
            let mut f = std::ptr::null_mut();
            unsafe { check(rp_adjoint_field(h, i as c_int, &mut f)) };
            Field2::<T>::from_handle(f, false, spaces[i])
        };
        let mut diagnostics = HashMap::new();
        for k in ["time", "Nu", "Nuvol", "Re"] {
            diagnostics.insert(k.to_string(), Vec::new());
        }
        Navier2DAdjoint { h, temp: [view(0), view(5)], ux: [view(1), view(6)], uy: [view(2), view(7)], pres: [view(3), view(4)],
                          ra, pr, time: 0.0, dt, dt_navier: 1e-2, diagnostics }
    }
    pub fn set_velocity(&mut self, amp: f64, m: f64, n: f64) {
        unsafe { check(rp_adjoint_set_velocity(self.h, amp, m, n)) }
    }
    pub fn set_temperature(&mut self, amp: f64, m: f64, n: f64) {
        unsafe { check(rp_adjoint_set_temperature(self.h, amp, m, n)) }
    }
    pub fn reset_time(&mut self) {
        unsafe { check(rp_adjoint_reset_time(self.h)) };
        self.time = 0.0;
    }
    /// (Nu, Nuvol, Re, |div|) -- `eval_nu`, `eval_nuvol`, `eval_re` (navier_adjoint.rs:952-992)
    pub fn eval(&mut self) -> (f64, f64, f64, f64) {
        let (mut a, mut b, mut c, mut d) = (0.0, 0.0, 0.0, 0.0);
        unsafe { check(rp_adjoint_eval(self.h, &mut a, &mut b, &mut c, &mut d)) };
        (a, b, c, d)
    }
    pub fn eval_nu(&mut self) -> f64 {
        self.eval().0
    }
    pub fn eval_nuvol(&mut self) -> f64 {
        self.eval().1
    }
    pub fn eval_re(&mut self) -> f64 {
        self.eval().2
    }
    /// |.|_2 of the smoothed and of the unsmoothed residual fields (ux, uy, temp)
    pub fn residuals(&mut self) -> ([f64; 3], [f64; 3]) {
        let (mut s, mut u) = ([0.0; 3], [0.0; 3]);
        unsafe { check(rp_adjoint_residuals(self.h, s.as_mut_ptr(), u.as_mut_ptr())) };
        (s, u)
    }
    /// Eigen set-up data of solver `which` (0 smoother ux|uy, 1 smoother temp, 2 pressure, 3 inner Navier2D pressure)
    pub fn export_eig(&self, which: usize) -> Option<EigData> {
        let mut s = std::ptr::null_mut();
        unsafe { check(rp_adjoint_solver(self.h, which as c_int, &mut s)) };
        let borrowed = std::mem::ManuallyDrop::new(SolverHandle { h: s });
        borrowed.export_eig()
    }
}
impl Navier2DAdjoint<f64> {
    /// `Navier2DAdjoint::new(nx, ny, ra, pr, dt, aspect, adiabatic)` (navier_adjoint.rs:197)
    pub fn new(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64, adiabatic: bool) -> Self {
        Self::make(nx, ny, ra, pr, dt, aspect, adiabatic, false)
    }
}
impl Navier2DAdjoint<Complex<f64>> {
    /// `Navier2DAdjoint::new_periodic(nx, ny, ra, pr, dt, aspect)` (navier_adjoint.rs:361)
    pub fn new_periodic(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64) -> Self {
        Self::make(nx, ny, ra, pr, dt, aspect, true, true)
    }
}
impl<T: SpectralScalar> Drop for Navier2DAdjoint<T> {
    fn drop(&mut self) {
        unsafe { rp_adjoint_destroy(self.h) };
    }
}
impl<T: SpectralScalar> Integrate for Navier2DAdjoint<T> {
    fn update(&mut self) {
        unsafe { check(rp_adjoint_update(self.h, 1)) };
        self.time += self.dt;
    }
    fn get_time(&self) -> f64 {
        let mut t = 0.0;
        unsafe { check(rp_adjoint_get_time(self.h, &mut t)) };
        t
    }
    fn get_dt(&self) -> f64 {
        self.dt
    }
    fn callback(&mut self) {
        let (nu, nuvol, re, div) = self.eval();
        let t = self.get_time();
        println!("time = {:4.2}      |div| = {:4.2e}     Nu = {:5.3e}     Nuv = {:5.3e}    Re = {:5.3e}", t, div, nu, nuvol, re);
        let (s, _) = self.residuals();
        println!("|U| = {:10.2e}", s[0]);
        println!("|V| = {:10.2e}", s[1]);
        println!("|T| = {:10.2e}", s[2]);
        for (k, x) in [("time", t), ("Nu", nu), ("Nuvol", nuvol), ("Re", re)] {
            self.diagnostics.get_mut(k).unwrap().push(x);
        }
    }
    /// NaN divergence or residual below 1e-8 (navier_adjoint.rs:892-910)
    fn exit(&mut self) -> bool {
        let mut stop = 0 as c_int;
        unsafe { check(rp_adjoint_exit(self.h, &mut stop)) };
        stop != 0
    }
}

const MAX_TIMESTEP: usize = 10_000_000;

/// `integrate` (src/lib.rs:155-187), unchanged.
pub fn integrate<T: Integrate>(pde: &mut T, max_time: f64, save_intervall: Option<f64>) {
    let mut timestep: usize = 0;
    let eps_dt = pde.get_dt() * 1e-4;
    loop {
        pde.update();
        timestep += 1;
        if let Some(dt_save) = &save_intervall {
            if (pde.get_time() % dt_save) < pde.get_dt() / 2. || (pde.get_time() % dt_save) > dt_save - pde.get_dt() / 2. {
                pde.callback();
            }
        }
        if pde.get_time() + eps_dt >= max_time {
            println!("time limit reached: {:?}", pde.get_time());
            break;
        }
        if timestep >= MAX_TIMESTEP {
            println!("timestep limit reached: {:?}", timestep);
            break;
        }
        if pde.exit() {
            println!("break criteria triggered");
            break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// slab decomposition over kx (periodic path, one process per GPU); the caller owns the exchange
// buffers (device pointers) and the all-to-all / peer mapping -- see INTEGRATION.md section 4
// ------------------------------------------------------------------------------------------------
pub mod slab {
    use super::*;
    pub fn phase1(nav: &Navier2D<Complex<f64>>, k0: usize, mkl: usize, out6: &[*mut f64; 6]) {
        unsafe { check(rp_navier_slab_phase1(nav.raw(), k0 as c_int, mkl as c_int, out6.as_ptr())) }
    }
    pub fn phase2(nav: &Navier2D<Complex<f64>>, j0: usize, nyl: usize, in6: &[*const f64; 6], work: *mut f64, out3: &[*mut f64; 3]) {
        unsafe { check(rp_navier_slab_phase2(nav.raw(), j0 as c_int, nyl as c_int, in6.as_ptr(), work, out3.as_ptr())) }
    }
    pub fn phase3(nav: &Navier2D<Complex<f64>>, k0: usize, mkl: usize, in3: &[*const f64; 3]) {
        unsafe { check(rp_navier_slab_phase3(nav.raw(), k0 as c_int, mkl as c_int, in3.as_ptr())) }
    }
    /// Fused transposes: the kernels store straight into the peers' buffers (`peers[a * world + q]`).
    pub fn phase1_p2p(nav: &Navier2D<Complex<f64>>, k0: usize, mkl: usize, joff: &[c_int], peers: &[*mut f64]) {
        let world = joff.len() - 1;
        assert_eq!(peers.len(), 6 * world);
        unsafe { check(rp_navier_slab_phase1_p2p(nav.raw(), k0 as c_int, mkl as c_int, world as c_int, joff.as_ptr(), peers.as_ptr())) }
    }
    pub fn phase2_p2p(nav: &Navier2D<Complex<f64>>, j0: usize, nyl: usize, in6: &[*const f64; 6], work: *mut f64, koff: &[c_int],
                      peers: &[*mut f64]) {
        let world = koff.len() - 1;
        assert_eq!(peers.len(), 3 * world);
        unsafe {
            check(rp_navier_slab_phase2_p2p(nav.raw(), j0 as c_int, nyl as c_int, in6.as_ptr(), work, world as c_int, koff.as_ptr(),
                                            peers.as_ptr()))
        }
    }
    /// Device memory that can be shared between the processes of one node (cudaMalloc + CUDA IPC).
    pub fn dev_alloc(bytes: usize) -> *mut c_void {
        let mut p = std::ptr::null_mut();
        unsafe { check(rp_dev_alloc(bytes, &mut p)) };
        p
    }
    pub fn dev_free(p: *mut c_void) {
        unsafe { check(rp_dev_free(p)) }
    }
    pub fn ipc_export(p: *mut c_void) -> [u8; 64] {
        let mut h = [0u8; 64];
        unsafe { check(rp_ipc_export(p, h.as_mut_ptr())) };
        h
    }
    pub fn ipc_open(handle: &[u8; 64]) -> *mut c_void {
        let mut p = std::ptr::null_mut();
        unsafe { check(rp_ipc_open(handle.as_ptr(), &mut p)) };
        p
    }
    pub fn ipc_close(p: *mut c_void) {
        unsafe { check(rp_ipc_close(p)) }
    }
}
