//! Thin Rust host layer over `librustpde_b200.so` (include/rustpde_b200.h).
//!
//! It keeps the signatures of the reference on the Navier2D path
//! (`Navier2D::new / new_periodic`, `set_velocity`, `set_temperature`,
//! `Integrate::{update, get_time, get_dt, callback, exit}`, `integrate`) and forwards the
//! arithmetic to the CUDA library.  A non-zero status becomes `panic!`, the reference's error
//! convention on this path.  NOT compiled in the development image (no Rust toolchain there);
//! the executable verification goes through the same C ABI from Python (tests/).
#![allow(non_camel_case_types)]
use ndarray::Array2;
use std::os::raw::{c_char, c_double, c_int};

#[repr(C)]
pub struct rp_field_t {
    _private: [u8; 0],
}
#[repr(C)]
pub struct rp_navier_t {
    _private: [u8; 0],
}

extern "C" {
    fn rp_init(device: c_int) -> c_int;
    fn rp_last_error() -> *const c_char;
    fn rp_navier_create(nx: c_int, ny: c_int, ra: c_double, pr: c_double, dt: c_double, aspect: c_double,
                        adiabatic: c_int, periodic: c_int, out: *mut *mut rp_navier_t) -> c_int;
    fn rp_navier_destroy(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_set_velocity(h: *mut rp_navier_t, amp: c_double, m: c_double, n: c_double) -> c_int;
    fn rp_navier_set_temperature(h: *mut rp_navier_t, amp: c_double, m: c_double, n: c_double) -> c_int;
    fn rp_navier_update(h: *mut rp_navier_t, nsteps: c_int) -> c_int;
    fn rp_navier_stage_state(h: *mut rp_navier_t, temp: *const c_double, n_temp: usize, ux: *const c_double, n_ux: usize,
                             uy: *const c_double, n_uy: usize, pres: *const c_double, n_pres: usize) -> c_int;
    fn rp_navier_commit_staged(h: *mut rp_navier_t) -> c_int;
    fn rp_navier_get_time(h: *mut rp_navier_t, t: *mut c_double) -> c_int;
    fn rp_navier_get_dt(h: *mut rp_navier_t, dt: *mut c_double) -> c_int;
    fn rp_navier_eval(h: *mut rp_navier_t, nu: *mut c_double, nuvol: *mut c_double, re: *mut c_double,
                      div: *mut c_double, ekin: *mut c_double) -> c_int;
    fn rp_navier_field(h: *mut rp_navier_t, which: c_int, out: *mut *mut rp_field_t) -> c_int;
    fn rp_field_shape(f: *mut rp_field_t, phys: *mut c_int, spec: *mut c_int, ortho: *mut c_int, is_complex: *mut c_int) -> c_int;
    fn rp_field_download_v(f: *mut rp_field_t, v: *mut c_double, len: usize) -> c_int;
    fn rp_field_upload_v(f: *mut rp_field_t, v: *const c_double, len: usize) -> c_int;
    fn rp_field_forward(f: *mut rp_field_t) -> c_int;
    fn rp_field_backward(f: *mut rp_field_t) -> c_int;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rp_last_error()) }.to_string_lossy().into_owned();
        panic!("rustpde_b200 error {}: {}", rc, msg);
    }
}

/// Integrate trait, identical to rustpde's (src/lib.rs:135-146).
pub trait Integrate {
    fn update(&mut self);
    fn get_time(&self) -> f64;
    fn get_dt(&self) -> f64;
    fn callback(&mut self);
    fn exit(&mut self) -> bool;
}

/// Which of Navier2D's pub fields (src/navier/navier.rs:153-195).
#[derive(Clone, Copy)]
pub enum Which {
    Temp = 0,
    Ux = 1,
    Uy = 2,
    Pres = 3,
    PseudoPres = 4,
}

/// Device-resident Navier2D; host mirrors of `v` are fetched on demand.
pub struct Navier2D {
    h: *mut rp_navier_t,
    pub diagnostics: std::collections::HashMap<String, Vec<f64>>,
}

impl Navier2D {
    /// `Navier2D::new(nx, ny, ra, pr, dt, aspect, adiabatic)` (navier.rs:219-227)
    pub fn new(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64, adiabatic: bool) -> Self {
        Self::make(nx, ny, ra, pr, dt, aspect, adiabatic, false)
    }
    /// `Navier2D::new_periodic(nx, ny, ra, pr, dt, aspect)` (navier.rs:384-391)
    pub fn new_periodic(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64) -> Self {
        Self::make(nx, ny, ra, pr, dt, aspect, true, true)
    }
    fn make(nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64, adiabatic: bool, periodic: bool) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe {
            check(rp_init(0));
            check(rp_navier_create(nx as c_int, ny as c_int, ra, pr, dt, aspect, adiabatic as c_int, periodic as c_int, &mut h));
        }
        let mut diagnostics = std::collections::HashMap::new();
        for k in ["time", "Nu", "Nuvol", "Re"] {
            diagnostics.insert(k.to_string(), Vec::new());
        }
        Navier2D { h, diagnostics }
    }
    pub fn set_velocity(&mut self, amp: f64, m: f64, n: f64) {
        unsafe { check(rp_navier_set_velocity(self.h, amp, m, n)) }
    }
    pub fn set_temperature(&mut self, amp: f64, m: f64, n: f64) {
        unsafe { check(rp_navier_set_temperature(self.h, amp, m, n)) }
    }
    /// Physical-space mirror of a field (`field.v` after `backward()`).
    pub fn v(&mut self, which: Which) -> Array2<f64> {
        unsafe {
            let mut f = std::ptr::null_mut();
            check(rp_navier_field(self.h, which as c_int, &mut f));
            let (mut ph, mut sp, mut or, mut cx) = ([0 as c_int; 2], [0 as c_int; 2], [0 as c_int; 2], 0 as c_int);
            check(rp_field_shape(f, ph.as_mut_ptr(), sp.as_mut_ptr(), or.as_mut_ptr(), &mut cx));
            check(rp_field_backward(f));
            let mut a = Array2::<f64>::zeros((ph[0] as usize, ph[1] as usize));
            check(rp_field_download_v(f, a.as_mut_ptr(), a.len()));
            a
        }
    }
    /// Overwrite a field from physical space (`field.v = a; field.forward()`).
    pub fn set_v(&mut self, which: Which, a: &Array2<f64>) {
        unsafe {
            let mut f = std::ptr::null_mut();
            check(rp_navier_field(self.h, which as c_int, &mut f));
            let a = a.as_standard_layout();
            check(rp_field_upload_v(f, a.as_ptr(), a.len()));
            check(rp_field_forward(f));
        }
    }
    fn eval(&mut self) -> (f64, f64, f64, f64) {
        let (mut nu, mut nuvol, mut re, mut div) = (0.0, 0.0, 0.0, 0.0);
        unsafe { check(rp_navier_eval(self.h, &mut nu, &mut nuvol, &mut re, &mut div, std::ptr::null_mut())) };
        (nu, nuvol, re, div)
    }
    pub fn eval_nu(&mut self) -> f64 { self.eval().0 }
    pub fn eval_nuvol(&mut self) -> f64 { self.eval().1 }
    pub fn eval_re(&mut self) -> f64 { self.eval().2 }
}

impl Drop for Navier2D {
    fn drop(&mut self) {
        unsafe { rp_navier_destroy(self.h) };
    }
}

impl Navier2D {
    /// Queue the upload of a full state (the `vhat` buffers of temp, ux, uy, pres[0], row-major as in the
    /// reference's `Array2`) on the copy stream; it overlaps the kernels of the step in flight.  The slices must
    /// outlive the matching `commit_staged` (page-locked memory keeps the copy asynchronous).
    pub fn stage_state(&mut self, temp: &[f64], ux: &[f64], uy: &[f64], pres: &[f64]) {
        unsafe {
            check(rp_navier_stage_state(self.h, temp.as_ptr(), temp.len(), ux.as_ptr(), ux.len(), uy.as_ptr(), uy.len(),
                                        pres.as_ptr(), pres.len()))
        }
    }
    /// Make the compute stream wait for the staged upload and move it into place.
    pub fn commit_staged(&mut self) {
        unsafe { check(rp_navier_commit_staged(self.h)) }
    }
}

impl Integrate for Navier2D {
    fn update(&mut self) {
        unsafe { check(rp_navier_update(self.h, 1)) }
    }
    fn get_time(&self) -> f64 {
        let mut t = 0.0;
        unsafe { check(rp_navier_get_time(self.h, &mut t)) };
        t
    }
    fn get_dt(&self) -> f64 {
        let mut t = 0.0;
        unsafe { check(rp_navier_get_dt(self.h, &mut t)) };
        t
    }
    fn callback(&mut self) {
        let (nu, nuvol, re, div) = self.eval();
        let t = self.get_time();
        println!("time = {:4.2}      |div| = {:4.2e}     Nu = {:5.3e}     Nuv = {:5.3e}    Re = {:5.3e}", t, div, nu, nuvol, re);
        for (k, v) in [("time", t), ("Nu", nu), ("Nuvol", nuvol), ("Re", re)] {
            self.diagnostics.get_mut(k).unwrap().push(v);
        }
    }
    fn exit(&mut self) -> bool {
        self.eval().3.is_nan()
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
