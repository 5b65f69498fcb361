/* rustpde_b200.h -- C ABI of librustpde_b200.so
 *
 * Drop-in boundary for the Navier2D time-step path of preiter93/rustpde
 * (SURVEY.md section 8b).  The reference is pure Rust with no FFI boundary of
 * its own; these are the entry points a thin Rust shim (see INTEGRATION.md and
 * rust/src/lib.rs) binds so that the reference's public API keeps its
 * signatures while the arithmetic runs in hand-written sm_100a CUDA kernels.
 *
 * Conventions
 *   - every function returns an int status (RP_OK == 0); on failure
 *     rp_last_error() holds a message.  The Rust shim turns a non-zero status
 *     into panic!(), the reference's error convention on this path
 *     (funspace/src/utils.rs:49-76, src/solver/fdma_tensor.rs:201-209).
 *   - arrays crossing the ABI are caller-owned HOST buffers, dense row-major
 *     [n0, n1] (ndarray's default layout, funspace/src/space2.rs:75-83), f64;
 *     complex data is interleaved (re, im) like num_complex::Complex<f64>.
 *     `len` arguments count doubles and are checked (RP_ERR_SHAPE).
 *   - device memory, streams and CUDA graphs are owned by the handles.
 *     A handle is not thread-safe (the reference API is &mut self).
 *   - there is no CPU fallback: every call below runs on the GPU.
 */
#ifndef RUSTPDE_B200_H
#define RUSTPDE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  RP_OK = 0,
  RP_ERR_INVALID = 1,  /* bad argument / unsupported combination */
  RP_ERR_SHAPE = 2,    /* size mismatch (reference: panic!)       */
  RP_ERR_CUDA = 3,
  RP_ERR_LAPACK = 4,   /* set-up eigendecomposition unavailable/failed */
  RP_ERR_INTERNAL = 5
};

/* funspace constructors (funspace/src/lib.rs:230-345) */
enum {
  RP_BASE_CHEBYSHEV = 0,         /* chebyshev(n)          */
  RP_BASE_CHEB_DIRICHLET = 1,    /* cheb_dirichlet(n)     */
  RP_BASE_CHEB_NEUMANN = 2,      /* cheb_neumann(n)       */
  RP_BASE_CHEB_DIRICHLET_BC = 3, /* cheb_dirichlet_bc(n)  (set-up helper, host side) */
  RP_BASE_CHEB_NEUMANN_BC = 4,   /* cheb_neumann_bc(n)    (set-up helper, host side) */
  RP_BASE_FOURIER_R2C = 5        /* fourier_r2c(n)        */
};

/* Navier2D pub fields (src/navier/navier.rs:153-195) */
enum { RP_FIELD_TEMP = 0, RP_FIELD_UX = 1, RP_FIELD_UY = 2, RP_FIELD_PRES = 3, RP_FIELD_PSEUDO_PRES = 4, RP_FIELD_WORK = 5 };

typedef struct rp_field rp_field_t;
typedef struct rp_solver rp_solver_t;
typedef struct rp_navier rp_navier_t;
typedef struct rp_adjoint rp_adjoint_t;

/* ---- library ----------------------------------------------------------- */
int rp_init(int device);                 /* selects the CUDA device, sets kernel attributes */
const char* rp_last_error(void);
int rp_version(void);
int rp_is_emulated(void);                /* 0 in the product library */
int rp_set_lapack_library(const char* path); /* dgeev/dgetri provider for solver set-up (utils.rs:66-106) */

/* ---- Field2 (src/field.rs:66-129) --------------------------------------- */
/* Field2::new(&Space2::new(&base_x, &base_y)) */
int rp_field_create(int kind_x, int nx, int kind_y, int ny, rp_field_t** out);
int rp_field_destroy(rp_field_t* f);
/* shapes: physical, spectral, ortho-spectral; is_complex = spectral type is Complex<f64> */
int rp_field_shape(rp_field_t* f, int phys[2], int spec[2], int ortho[2], int* is_complex);
int rp_field_coords(rp_field_t* f, int axis, double* x, size_t len);   /* pub x  */
int rp_field_dx(rp_field_t* f, int axis, double* dx, size_t len);      /* pub dx */
int rp_field_upload_v(rp_field_t* f, const double* v, size_t len);      /* pub v    */
int rp_field_download_v(rp_field_t* f, double* v, size_t len);
int rp_field_upload_vhat(rp_field_t* f, const double* vhat, size_t len);/* pub vhat */
int rp_field_download_vhat(rp_field_t* f, double* vhat, size_t len);
/* rows [row0, row0+nrows) of the pub vhat (a kx slab of the spectral array, dense host block) */
int rp_field_upload_vhat_rows(rp_field_t* f, int row0, int nrows, const double* vhat, size_t len);
int rp_field_download_vhat_rows(rp_field_t* f, int row0, int nrows, double* vhat, size_t len);
int rp_field_forward(rp_field_t* f);                                    /* field.rs:103-105 */
int rp_field_backward(rp_field_t* f);                                   /* field.rs:108-110 */
int rp_field_to_ortho(rp_field_t* f, double* out, size_t len);          /* field.rs:113-115 */
int rp_field_from_ortho(rp_field_t* f, const double* in, size_t len);   /* field.rs:118-123 */
/* field.rs:127-129; scale may be NULL (None) or point to [sx, sy] */
int rp_field_gradient(rp_field_t* f, int dx, int dy, const double* scale, double* out, size_t len);
int rp_field_average(rp_field_t* f, double* out);                       /* average.rs:51-57 */
int rp_field_average_axis(rp_field_t* f, int axis, double* out, size_t len); /* average.rs:25-33 (axis 0) */

/* ---- solvers (src/solver.rs:56-155) -------------------------------------- */
int rp_hholtz_create(rp_field_t* f, double cx, double cy, double alpha, rp_solver_t** out); /* Hholtz::new (alpha=1) / new2, hholtz.rs:42,81 */
int rp_hholtz_adi_create(rp_field_t* f, double cx, double cy, rp_solver_t** out);          /* HholtzAdi::new, hholtz_adi.rs:44 */
int rp_poisson_create(rp_field_t* f, double cx, double cy, rp_solver_t** out);             /* Poisson::new, poisson.rs:50 */
/* Same, with the eigen set-up data (lam[m], Q[m*m], P = Q^-1 Cx^-1 [m*m], row-major,
 * m = nx-2) supplied by the caller instead of LAPACK -- this is what makes
 * <=1e-10 parity of the fast-diagonalisation solve well defined. */
int rp_hholtz_create_with_eig(rp_field_t* f, double cx, double cy, double alpha, const double* lam, const double* q,
                              const double* p, rp_solver_t** out);
int rp_poisson_create_with_eig(rp_field_t* f, double cx, double cy, const double* lam, const double* q, const double* p,
                               rp_solver_t** out);
int rp_solver_eig_size(rp_solver_t* s, int* m, int* has_matrices);
/* The exported eigenvalues are those of inv(Cx) Ax in the reference order (descending, utils.rs:80-94) BEFORE the
 * Poisson singularity shift (poisson.rs:80-83); *_create_with_eig expects the same and applies the shift itself,
 * so create_with_eig(export_eig()) reproduces the solver exactly. */
int rp_solver_export_eig(rp_solver_t* s, double* lam, double* q, double* p);
/* Solve::solve(&self, input, output, axis) ; input [n0,n1] ortho, output composite */
int rp_solver_solve(rp_solver_t* s, const double* in, size_t in_len, double* out, size_t out_len, int is_complex);
/* Measurement aids (no reference counterpart): repeat the solve `reps` times on the rhs staged by the last
 * rp_solver_solve (device-resident, asynchronous; rp_solver_sync waits), and report which kernels serve real-data
 * solves (specialised = 1: hand-specialised x/y kernels + DMMA GEMMs, else generic lane programs). */
int rp_solver_solve_resident(rp_solver_t* s, int reps, int is_complex);
int rp_solver_sync(rp_solver_t* s);
int rp_solver_path(rp_solver_t* s, int* specialised, int* split_gemm, int* launches);
int rp_solver_destroy(rp_solver_t* s);

/* ---- Navier2D (src/navier/navier.rs) -------------------------------------- */
/* Navier2D::new (219-307) when periodic == 0, Navier2D::new_periodic (384-467) otherwise.
 * Unlike the reference constructors, which add an UNSEEDED random disturbance
 * (navier.rs:304,464), all fields start at zero: set deterministic ICs. */
int rp_navier_create(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic, int periodic,
                     rp_navier_t** out);
int rp_navier_create_with_eig(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic,
                              const double* lam, const double* q, const double* p, rp_navier_t** out);
int rp_navier_destroy(rp_navier_t* h);
int rp_navier_set_velocity(rp_navier_t* h, double amp, double m, double n);    /* 927-930 */
int rp_navier_set_temperature(rp_navier_t* h, double amp, double m, double n); /* 934-936 */
int rp_navier_set_tempbc_ortho(rp_navier_t* h, const double* that_bc, size_t len); /* set_temp_bc, 517-519 (ortho coefficients) */
int rp_navier_set_dealias(rp_navier_t* h, int on);                             /* pub dealias */
/* pub solid = Some([mask, value]) (navier.rs:191, solid_masks.rs:34-175): volume penalisation -1/eta * mask * (u - value),
 * eta = 1e-2 (navier.rs:552-608); both [nx, ny] on the physical grid, value may be NULL (zeros), mask NULL switches it off.
 * Before the first update(); specialised kernels only. */
int rp_navier_set_solid(rp_navier_t* h, const double* mask, const double* value, size_t len);
int rp_navier_update(rp_navier_t* h, int nsteps);                              /* Integrate::update, 737-765 (x nsteps, asynchronous) */
int rp_navier_sync(rp_navier_t* h);
/* Double-buffered upload of the pub `vhat` arrays of temp, ux, uy, pres[0] (navier.rs:153-160; what `read()`
 * 963-981 assigns): stage_state queues the host-to-device copies on a copy stream and returns (the host buffers
 * must stay valid -- and should be page-locked -- until the commit); commit_staged makes the compute stream wait
 * for them and moves the staged state into place.  The copies of the next state overlap update() of the current one. */
int rp_navier_stage_state(rp_navier_t* h, const double* temp, size_t len_temp, const double* ux, size_t len_ux,
                          const double* uy, size_t len_uy, const double* pres, size_t len_pres);
int rp_navier_commit_staged(rp_navier_t* h);
/* Asynchronous download of the same four pub `vhat` arrays (what `write()` 975-1013 stores): the state is
 * snapshotted on the compute stream and copied to the (page-locked) host buffers on a second copy stream, so the
 * copies overlap the following update()s; fetch_wait blocks until the buffers are complete. */
int rp_navier_fetch_state(rp_navier_t* h, double* temp, size_t len_temp, double* ux, size_t len_ux, double* uy, size_t len_uy,
                          double* pres, size_t len_pres);
int rp_navier_fetch_wait(rp_navier_t* h);
/* Integrate::exit (855-862) without a host sync per step: div_async queues |div u|_2 of the current state and its
 * copy to a page-locked slot; div_poll returns the most recent value that has arrived (ready = 0: none yet;
 * wait != 0 blocks for the outstanding request).  integrate() then sees a NaN one check late. */
int rp_navier_div_async(rp_navier_t* h);
int rp_navier_div_poll(rp_navier_t* h, int wait, double* div_norm, int* ready);
/* write (975-1013) / read (963-972): checkpoint and restart with the reference's group / dataset layout
 * (temp, ux, uy, pres: v, vhat | vhat_re + vhat_im; x, dx, y, dy; time, ra, pr, nu, kappa) in the RPSNAP1 container
 * (csrc/snapshot.cu; rustpde_b200/snapshot.py converts to / from HDF5 where h5py exists).  read keeps the
 * truncate / keep-the-rest "broadcast" of field/read.rs:113-122 when the stored shape differs. */
int rp_navier_write_snapshot(rp_navier_t* h, const char* path);
int rp_navier_read_snapshot(rp_navier_t* h, const char* path);
int rp_navier_get_time(rp_navier_t* h, double* time);                          /* get_time */
int rp_navier_get_dt(rp_navier_t* h, double* dt);                              /* get_dt */
int rp_navier_reset_time(rp_navier_t* h);                                      /* 951-953 */
int rp_navier_params(rp_navier_t* h, double* nu, double* ka, double scale[2]); /* pub nu, ka, scale */
/* eval_nu / eval_nuvol / eval_re (890-921), |div|_2 (exit(), 855-879) and <(ux^2+uy^2)/2>;
 * any pointer may be NULL */
int rp_navier_eval(rp_navier_t* h, double* nu, double* nuvol, double* re, double* div_norm, double* ekin);
int rp_navier_field(rp_navier_t* h, int which, rp_field_t** out);              /* borrowed handle */
int rp_navier_export_eig(rp_navier_t* h, double* lam, double* q, double* p);   /* pressure Poisson set-up data */
int rp_navier_launches_per_step(rp_navier_t* h, int* n);
int rp_navier_set_graph(rp_navier_t* h, int on);                               /* CUDA-graph replay of update() (default on) */
/* Slab decomposition of the periodic step over the Fourier modes kx (one process per GPU; the caller owns the
   exchange buffers -- DEVICE pointers, dense row-major, interleaved complex -- and performs the all-to-all between
   the phases; rustpde_b200/slab.py is the driver).  The object's own state arrays hold the rows [k0, k0+mkl).
     phase1: out6[f]   = B_y S_y u_f,  out6[3+f] = B_y D_y S_y u_f / sy   f = ux, uy, temp   each [mkl, ny] complex
     phase2: in6[.]    = the same six arrays after the exchange, [nx/2+1, nyl] complex (columns j0..j0+nyl of physical y);
             work      = 8 * nx * nyl doubles;  out3[f] = r2c(u . grad f) with the kx dealias cut, [nx/2+1, nyl] complex
     phase3: in3[f]    = out3[f] after the exchange, [mkl, ny] complex; forward DCT-y, rhs, solves, projection; time += dt */
int rp_navier_slab_phase1(rp_navier_t* h, int k0, int mkl, double* const* out6);
int rp_navier_slab_phase2(rp_navier_t* h, int j0, int nyl, const double* const* in6, double* work, double* const* out3);
int rp_navier_slab_phase3(rp_navier_t* h, int k0, int mkl, const double* const* in3);
/* Fused transposes over NVLink peer memory: instead of dense outputs + all-to-all, the phase-1 / phase-2 kernels
   store every element straight into the buffer of the rank that owns it (buffers mapped with CUDA IPC).
     joff[world+1] / koff[world+1]: first physical-y column / Fourier mode owned by each rank;
     peers[a * world + q]: device address of array a (phase 1: the six in6 arrays, phase 2: the three in3 arrays)
     on rank q, as mapped into THIS process.  The caller separates the phases with a cross-rank barrier. */
int rp_navier_slab_phase1_p2p(rp_navier_t* h, int k0, int mkl, int world, const int* joff, double* const* peers);
int rp_navier_slab_phase2_p2p(rp_navier_t* h, int j0, int nyl, const double* const* in6, double* work, int world,
                              const int* koff, double* const* peers);
/* device memory shared between the processes of one node (cudaMalloc + cudaIpc*MemHandle) */
int rp_dev_alloc(size_t bytes, void** out);
int rp_dev_free(void* p);
int rp_ipc_export(void* p, unsigned char handle[64]);
int rp_ipc_open(const unsigned char handle[64], void** out);
int rp_ipc_close(void* p);
/* which kernels serve update(): specialised = 1 -> hand-specialised x/y pass kernels (else generic lane programs);
   split_gemm = 1 -> pressure Poisson runs the even/odd parity-split GEMM pairs (exactly checkerboard set-up data) */
int rp_navier_kernel_path(rp_navier_t* h, int* specialised, int* split_gemm);
/* Measurement aids (no reference counterpart): per-launch device time of update(), averaged over
 * `reps` eagerly launched steps (CUDA events on the launching stream; advances the solution), and
 * the name / algorithmic bytes / flops of launch i. */
int rp_navier_profile(rp_navier_t* h, int reps, double* ms, size_t cap, int* nops);
int rp_navier_op_info(rp_navier_t* h, int i, char* name, size_t name_len, double* bytes, double* flops);

/* ---- Navier2DAdjoint (src/navier/navier_adjoint.rs:128-1068) --------------------- */
/* Navier2DAdjoint::new (197) / new_periodic (361): steady-state adjoint descent; every update() runs one update() of an
 * inner Navier2D (dt_navier = 1e-2), three Hholtz smoother solves and one pressure Poisson solve.  Fields start at zero. */
int rp_adjoint_create(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic, int periodic, rp_adjoint_t** out);
int rp_adjoint_destroy(rp_adjoint_t* h);
int rp_adjoint_set_velocity(rp_adjoint_t* h, double amp, double m, double n);     /* 994-999 */
int rp_adjoint_set_temperature(rp_adjoint_t* h, double amp, double m, double n);  /* 1001-1003 */
int rp_adjoint_update(rp_adjoint_t* h, int nsteps);                               /* Integrate::update, 778-805 */
int rp_adjoint_get_time(rp_adjoint_t* h, double* time);
int rp_adjoint_reset_time(rp_adjoint_t* h);                                       /* 1006-1009 */
int rp_adjoint_eval(rp_adjoint_t* h, double* nu, double* nuvol, double* re, double* div_norm); /* 952-992; any pointer may be NULL */
/* |.|_2 of the smoothed residual fields ux[1], uy[1], temp[1] (what exit() sums, 903-906) and of the unsmoothed ones (868-870) */
int rp_adjoint_residuals(rp_adjoint_t* h, double smooth[3], double unsmooth[3]);
int rp_adjoint_exit(rp_adjoint_t* h, int* stop);                                  /* 892-910: NaN or residual < 1e-8 */
/* borrowed handles: fields 0 temp, 1 ux, 2 uy, 3 pres, 4 pseudo pressure, 5 / 6 / 7 residual temp / ux / uy;
 * solvers 0 smoother of ux and uy, 1 smoother of temp, 2 pressure Poisson, 3 the inner Navier2D's Poisson (for export_eig) */
int rp_adjoint_field(rp_adjoint_t* h, int which, rp_field_t** out);
int rp_adjoint_solver(rp_adjoint_t* h, int which, rp_solver_t** out);

#ifdef __cplusplus
}
#endif
#endif /* RUSTPDE_B200_H */
