// capi.cu -- extern "C" boundary (include/rustpde_b200.h).
#include <cstring>
#include <string>

#include "../../include/rustpde_b200.h"
#include "model.h"

using namespace rp;

struct rp_field {
  Field2* f;
  bool owned;
};
struct rp_solver {
  Solver2* s;
};
struct rp_navier {
  Navier2D* n;
  rp_field views[6];
};
struct rp_adjoint {
  Navier2DAdjoint* a;
  rp_field views[8];
  rp_solver solvers[4];
};

static thread_local std::string g_last_error;

template <class F>
static int guard(F&& fn) {
  try {
    fn();
    return RP_OK;
  } catch (const rp::Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return RP_ERR_INTERNAL;
  } catch (...) {
    g_last_error = "unknown error";
    return RP_ERR_INTERNAL;
  }
}

static void need(bool c, int code, const char* msg) {
  if (!c) throw rp::Error(code, msg);
}

#pragma GCC visibility push(default)
extern "C" {

int rp_init(int device) {
  return guard([&] {
#ifndef RP_EMU
    RP_CUDA_CHECK(cudaSetDevice(device));
    RP_CUDA_CHECK(cudaFree(0));
#else
    (void)device;
#endif
    init_kernels();
  });
}
const char* rp_last_error(void) { return g_last_error.c_str(); }
int rp_version(void) { return 100; }
int rp_is_emulated(void) {
#ifdef RP_EMU
  return 1;
#else
  return 0;
#endif
}
int rp_set_lapack_library(const char* path) {
  return guard([&] { lapack_set_library(path); });
}

// ---- Field2 ---------------------------------------------------------------
int rp_field_create(int kind_x, int nx, int kind_y, int ny, rp_field_t** out) {
  return guard([&] {
    need(out != nullptr, RP_ERR_INVALID, "null out pointer");
    Space2 sp{get_base(kind_x, nx), get_base(kind_y, ny)};
    *out = new rp_field{new Field2(sp), true};
  });
}
int rp_field_destroy(rp_field_t* f) {
  return guard([&] {
    if (!f) return;
    if (f->owned) {
      delete f->f;
      delete f;
    }
  });
}
int rp_field_shape(rp_field_t* f, int phys[2], int spec[2], int ortho[2], int* is_complex) {
  return guard([&] {
    need(f, RP_ERR_INVALID, "null field");
    if (phys) phys[0] = f->f->n0, phys[1] = f->f->n1;
    if (spec) spec[0] = f->f->m0, spec[1] = f->f->m1;
    if (ortho) ortho[0] = f->f->o0, ortho[1] = f->f->o1;
    if (is_complex) *is_complex = f->f->cplx ? 1 : 0;
  });
}
static int copy_vec(const std::vector<double>& v, double* out, size_t len) {
  return guard([&] {
    need(len == v.size(), RP_ERR_SHAPE, "coordinate length mismatch");
    std::copy(v.begin(), v.end(), out);
  });
}
int rp_field_coords(rp_field_t* f, int axis, double* x, size_t len) {
  if (!f || axis < 0 || axis > 1) return RP_ERR_INVALID;
  return copy_vec(f->f->x[axis], x, len);
}
int rp_field_dx(rp_field_t* f, int axis, double* dx, size_t len) {
  if (!f || axis < 0 || axis > 1) return RP_ERR_INVALID;
  return copy_vec(f->f->dx[axis], dx, len);
}
static size_t arr_len(const Arr& a) { return (size_t)a.rows * a.cols * (a.cplx ? 2 : 1); }

int rp_field_upload_v(rp_field_t* f, const double* v, size_t len) {
  return guard([&] {
    need(f && v, RP_ERR_INVALID, "null argument");
    need(len == arr_len(f->f->v), RP_ERR_SHAPE, "v: size mismatch");
    f->f->v.upload(v, f->f->stream);
    rt::sync(f->f->stream);
  });
}
int rp_field_download_v(rp_field_t* f, double* v, size_t len) {
  return guard([&] {
    need(f && v, RP_ERR_INVALID, "null argument");
    need(len == arr_len(f->f->v), RP_ERR_SHAPE, "v: size mismatch");
    f->f->v.download(v, f->f->stream);
    rt::sync(f->f->stream);
  });
}
int rp_field_upload_vhat(rp_field_t* f, const double* v, size_t len) {
  return guard([&] {
    need(f && v, RP_ERR_INVALID, "null argument");
    need(len == arr_len(f->f->vhat), RP_ERR_SHAPE, "vhat: size mismatch");
    f->f->vhat.upload(v, f->f->stream);
    ++f->f->vhat_version;
    rt::sync(f->f->stream);
  });
}
int rp_field_download_vhat(rp_field_t* f, double* v, size_t len) {
  return guard([&] {
    need(f && v, RP_ERR_INVALID, "null argument");
    need(len == arr_len(f->f->vhat), RP_ERR_SHAPE, "vhat: size mismatch");
    f->f->vhat.download(v, f->f->stream);
    rt::sync(f->f->stream);
  });
}
// rows [row0, row0 + nrows) of vhat <-> a dense host block (the slab a rank owns in the kx decomposition)
static void vhat_rows(rp_field_t* f, int row0, int nrows, double* host, size_t len, bool up) {
  need(f && host, RP_ERR_INVALID, "null argument");
  Arr& a = f->f->vhat;
  need(row0 >= 0 && nrows > 0 && row0 + nrows <= a.rows, RP_ERR_INVALID, "vhat_rows: bad row range");
  need(len == (size_t)nrows * a.cols * (a.cplx ? 2 : 1), RP_ERR_SHAPE, "vhat_rows: size mismatch");
  const size_t w = (size_t)a.cols * a.elem_bytes(), pitch = (size_t)a.ld * a.elem_bytes();
  char* d = (char*)a.buf.p + (size_t)row0 * pitch;
  if (up) {
    rt::h2d_2d(d, pitch, host, w, w, nrows, f->f->stream);
    ++f->f->vhat_version;
  } else {
    rt::d2h_2d(host, w, d, pitch, w, nrows, f->f->stream);
  }
  rt::sync(f->f->stream);
}
int rp_field_upload_vhat_rows(rp_field_t* f, int row0, int nrows, const double* vhat, size_t len) {
  return guard([&] { vhat_rows(f, row0, nrows, const_cast<double*>(vhat), len, true); });
}
int rp_field_download_vhat_rows(rp_field_t* f, int row0, int nrows, double* vhat, size_t len) {
  return guard([&] { vhat_rows(f, row0, nrows, vhat, len, false); });
}
int rp_field_forward(rp_field_t* f) {
  return guard([&] {
    need(f, RP_ERR_INVALID, "null field");
    f->f->forward();
  });
}
int rp_field_backward(rp_field_t* f) {
  return guard([&] {
    need(f, RP_ERR_INVALID, "null field");
    f->f->backward();
  });
}
int rp_field_to_ortho(rp_field_t* f, double* out, size_t len) {
  return guard([&] {
    need(f && out, RP_ERR_INVALID, "null argument");
    need(len == arr_len(f->f->ortho), RP_ERR_SHAPE, "to_ortho: output size mismatch");
    f->f->to_ortho();
    f->f->ortho.download(out, f->f->stream);
    rt::sync(f->f->stream);
  });
}
int rp_field_from_ortho(rp_field_t* f, const double* in, size_t len) {
  return guard([&] {
    need(f && in, RP_ERR_INVALID, "null argument");
    need(len == arr_len(f->f->ortho), RP_ERR_SHAPE, "from_ortho: input size mismatch");
    f->f->ortho.upload(in, f->f->stream);
    f->f->from_ortho();
    rt::sync(f->f->stream);
  });
}
int rp_field_gradient(rp_field_t* f, int dx, int dy, const double* scale, double* out, size_t len) {
  return guard([&] {
    need(f && out, RP_ERR_INVALID, "null argument");
    need(dx >= 0 && dy >= 0 && dx <= 4 && dy <= 4, RP_ERR_INVALID, "derivative order out of range");
    need(len == arr_len(f->f->ortho), RP_ERR_SHAPE, "gradient: output size mismatch");
    f->f->gradient(dx, dy, scale);
    f->f->ortho.download(out, f->f->stream);
    rt::sync(f->f->stream);
  });
}
int rp_field_average(rp_field_t* f, double* out) {
  return guard([&] {
    need(f && out, RP_ERR_INVALID, "null argument");
    *out = f->f->average();
  });
}
int rp_field_average_axis(rp_field_t* f, int axis, double* out, size_t len) {
  return guard([&] {
    need(f && out, RP_ERR_INVALID, "null argument");
    need(axis == 0, RP_ERR_INVALID, "average_axis: only axis 0 is on the Navier2D path (functions.rs:33)");
    need(len == (size_t)f->f->n1, RP_ERR_SHAPE, "average_axis: output size mismatch");
    std::vector<double> v;
    f->f->average_axis0(v);
    std::copy(v.begin(), v.end(), out);
  });
}

// ---- solvers --------------------------------------------------------------
static int make_solver(int kind, rp_field_t* f, double cx, double cy, double alpha, const double* lam, const double* q,
                       const double* p, rp_solver_t** out) {
  return guard([&] {
    need(f && out, RP_ERR_INVALID, "null argument");
    EigData e;
    e.lam = lam, e.q = q, e.p = p;
    *out = new rp_solver{new Solver2(kind, f->f->sp, cx, cy, alpha, (lam && q && p) ? &e : nullptr)};
  });
}
int rp_hholtz_create(rp_field_t* f, double cx, double cy, double alpha, rp_solver_t** out) {
  return make_solver(SOLVER_HHOLTZ, f, cx, cy, alpha, nullptr, nullptr, nullptr, out);
}
int rp_hholtz_adi_create(rp_field_t* f, double cx, double cy, rp_solver_t** out) {
  return make_solver(SOLVER_HHOLTZ_ADI, f, cx, cy, 1.0, nullptr, nullptr, nullptr, out);
}
int rp_poisson_create(rp_field_t* f, double cx, double cy, rp_solver_t** out) {
  return make_solver(SOLVER_POISSON, f, cx, cy, 0.0, nullptr, nullptr, nullptr, out);
}
int rp_hholtz_create_with_eig(rp_field_t* f, double cx, double cy, double alpha, const double* lam, const double* q,
                              const double* p, rp_solver_t** out) {
  return make_solver(SOLVER_HHOLTZ, f, cx, cy, alpha, lam, q, p, out);
}
int rp_poisson_create_with_eig(rp_field_t* f, double cx, double cy, const double* lam, const double* q, const double* p,
                               rp_solver_t** out) {
  return make_solver(SOLVER_POISSON, f, cx, cy, 0.0, lam, q, p, out);
}
int rp_solver_eig_size(rp_solver_t* s, int* m, int* has_matrices) {
  return guard([&] {
    need(s, RP_ERR_INVALID, "null solver");
    if (m) *m = s->s->m0;
    if (has_matrices) *has_matrices = (s->s->kind != SOLVER_HHOLTZ_ADI && !s->s->x_fourier) ? 1 : 0;
  });
}
int rp_solver_export_eig(rp_solver_t* s, double* lam, double* q, double* p) {
  return guard([&] {
    need(s, RP_ERR_INVALID, "null solver");
    s->s->export_eig(lam, q, p);
  });
}
int rp_solver_solve(rp_solver_t* s, const double* in, size_t in_len, double* out, size_t out_len, int is_complex) {
  return guard([&] {
    need(s && in && out, RP_ERR_INVALID, "null argument");
    Solver2& S = *s->s;
    const bool cd = is_complex != 0;
    need(!S.x_fourier || cd, RP_ERR_INVALID, "a Fourier axis needs complex data");
    const size_t w = cd ? 2 : 1;
    // reference: panic!("Dimension mismatch in Tensor! ...")  (fdma_tensor.rs:201-209)
    need(in_len == (size_t)S.n0 * S.n1 * w, RP_ERR_SHAPE, "Dimension mismatch in solver input");
    need(out_len == (size_t)S.m0 * S.m1 * w, RP_ERR_SHAPE, "Dimension mismatch in solver output");
    const bool lanes_c = cd || S.x_fourier;
    Arr& ain = lanes_c ? S.in_c : S.in_r;
    Arr& aout = lanes_c ? S.out_c : S.out_r;
    if (!ain.buf.p) {
      ain.alloc(S.n0, S.n1, lanes_c);
      aout.alloc(S.m0, S.m1, lanes_c);
    }
    ain.upload(in, S.stream);
    S.solve(cd);
    aout.download(out, S.stream);
    rt::sync(S.stream);
  });
}
int rp_solver_solve_resident(rp_solver_t* s, int reps, int is_complex) {
  return guard([&] {
    need(s && reps >= 0, RP_ERR_INVALID, "bad argument");
    Solver2& S = *s->s;
    const bool cd = is_complex != 0;
    const bool lanes_c = cd || S.x_fourier;
    need((lanes_c ? S.in_c : S.in_r).buf.p != nullptr, RP_ERR_INVALID, "solve_resident: call rp_solver_solve once first (stages the rhs)");
    for (int i = 0; i < reps; ++i) S.solve(cd);
  });
}
int rp_solver_sync(rp_solver_t* s) {
  return guard([&] {
    need(s, RP_ERR_INVALID, "null solver");
    rt::sync(s->s->stream);
  });
}
int rp_solver_path(rp_solver_t* s, int* specialised, int* split_gemm, int* launches) {
  return guard([&] {
    need(s, RP_ERR_INVALID, "null solver");
    if (specialised) *specialised = s->s->fast_path() ? 1 : 0;
    if (split_gemm) *split_gemm = s->s->ts.split ? 1 : 0;
    if (launches) *launches = s->s->launches_per_solve(false);
  });
}
int rp_solver_destroy(rp_solver_t* s) {
  return guard([&] {
    if (!s) return;
    delete s->s;
    delete s;
  });
}

// ---- Navier2D ---------------------------------------------------------------
static int make_navier(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic, int periodic,
                       const double* lam, const double* q, const double* p, rp_navier_t** out) {
  return guard([&] {
    need(out != nullptr, RP_ERR_INVALID, "null out pointer");
    need(nx >= 8 && ny >= 8, RP_ERR_INVALID, "grid too small");
    need(ra > 0 && pr > 0 && dt > 0 && aspect > 0, RP_ERR_INVALID, "ra, pr, dt, aspect must be positive");
    EigData e;
    e.lam = lam, e.q = q, e.p = p;
    init_kernels();
    auto* h = new rp_navier;
    h->n = new Navier2D(nx, ny, ra, pr, dt, aspect, adiabatic != 0, periodic != 0, (lam && q && p) ? &e : nullptr);
    for (int i = 0; i < 6; ++i) h->views[i] = rp_field{h->n->field_by_index(i), false};
    *out = h;
  });
}
int rp_navier_create(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic, int periodic,
                     rp_navier_t** out) {
  return make_navier(nx, ny, ra, pr, dt, aspect, adiabatic, periodic, nullptr, nullptr, nullptr, out);
}
int rp_navier_create_with_eig(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic,
                              const double* lam, const double* q, const double* p, rp_navier_t** out) {
  return make_navier(nx, ny, ra, pr, dt, aspect, adiabatic, 0, lam, q, p, out);
}
int rp_navier_destroy(rp_navier_t* h) {
  return guard([&] {
    if (!h) return;
    delete h->n;
    delete h;
  });
}
#define NAV_GUARD(...)                       \
  return guard([&] {                         \
    need(h, RP_ERR_INVALID, "null handle");  \
    Navier2D& N = *h->n;                     \
    (void)N;                                 \
    __VA_ARGS__;                             \
  })
int rp_navier_set_velocity(rp_navier_t* h, double amp, double m, double n) { NAV_GUARD(N.set_velocity(amp, m, n)); }
int rp_navier_set_temperature(rp_navier_t* h, double amp, double m, double n) { NAV_GUARD(N.set_temperature(amp, m, n)); }
int rp_navier_set_tempbc_ortho(rp_navier_t* h, const double* tb, size_t len) {
  NAV_GUARD({
    need(tb, RP_ERR_INVALID, "null argument");
    need(len == arr_len(N.field->ortho), RP_ERR_SHAPE, "set_tempbc_ortho: size mismatch");
    N.set_tempbc_ortho(tb);
  });
}
int rp_navier_set_solid(rp_navier_t* h, const double* mask, const double* value, size_t len) {
  NAV_GUARD({
    need(mask == nullptr || len == (size_t)N.nx * N.ny, RP_ERR_SHAPE, "set_solid: size mismatch");
    N.set_solid(mask, value);
  });
}
int rp_navier_set_dealias(rp_navier_t* h, int on) {
  NAV_GUARD({
    need(N.launches_per_step() == 0, RP_ERR_INVALID, "set_dealias must be called before the first update()");
    N.dealias = on != 0;
  });
}
int rp_navier_update(rp_navier_t* h, int nsteps) {
  NAV_GUARD({
    need(nsteps >= 0, RP_ERR_INVALID, "nsteps < 0");
    N.update(nsteps);
  });
}
int rp_navier_stage_state(rp_navier_t* h, const double* temp, size_t len_temp, const double* ux, size_t len_ux,
                          const double* uy, size_t len_uy, const double* pres, size_t len_pres) {
  return guard([&] {
    need(h && temp && ux && uy && pres, RP_ERR_INVALID, "null argument");
    Navier2D& n = *h->n;
    need(len_temp == arr_len(n.temp->vhat) && len_ux == arr_len(n.ux->vhat) && len_uy == arr_len(n.uy->vhat) &&
             len_pres == arr_len(n.pres0->vhat),
         RP_ERR_SHAPE, "stage_state: size mismatch");
    n.stage_state(temp, ux, uy, pres);
  });
}
int rp_navier_commit_staged(rp_navier_t* h) {
  return guard([&] {
    need(h, RP_ERR_INVALID, "null handle");
    h->n->commit_staged();
  });
}
int rp_navier_fetch_state(rp_navier_t* h, double* temp, size_t len_temp, double* ux, size_t len_ux, double* uy, size_t len_uy,
                          double* pres, size_t len_pres) {
  return guard([&] {
    need(h && temp && ux && uy && pres, RP_ERR_INVALID, "null argument");
    Navier2D& n = *h->n;
    need(len_temp == arr_len(n.temp->vhat) && len_ux == arr_len(n.ux->vhat) && len_uy == arr_len(n.uy->vhat) &&
             len_pres == arr_len(n.pres0->vhat),
         RP_ERR_SHAPE, "fetch_state: size mismatch");
    n.fetch_state(temp, ux, uy, pres);
  });
}
int rp_navier_fetch_wait(rp_navier_t* h) { NAV_GUARD(N.fetch_wait()); }
int rp_navier_div_async(rp_navier_t* h) { NAV_GUARD(N.div_async()); }
int rp_navier_div_poll(rp_navier_t* h, int wait, double* div_norm, int* ready) {
  NAV_GUARD({
    const bool ok = N.div_poll(div_norm, wait != 0);
    if (ready) *ready = ok ? 1 : 0;
  });
}
int rp_navier_write_snapshot(rp_navier_t* h, const char* path) {
  NAV_GUARD({
    need(path != nullptr, RP_ERR_INVALID, "null path");
    N.write_snapshot(path);
  });
}
int rp_navier_read_snapshot(rp_navier_t* h, const char* path) {
  NAV_GUARD({
    need(path != nullptr, RP_ERR_INVALID, "null path");
    N.read_snapshot(path);
  });
}
int rp_navier_sync(rp_navier_t* h) { NAV_GUARD(N.sync()); }
int rp_navier_get_time(rp_navier_t* h, double* t) { NAV_GUARD(if (t) *t = N.time); }
int rp_navier_get_dt(rp_navier_t* h, double* dt) { NAV_GUARD(if (dt) *dt = N.dt); }
int rp_navier_reset_time(rp_navier_t* h) { NAV_GUARD(N.time = 0.0); }
int rp_navier_params(rp_navier_t* h, double* nu, double* ka, double scale[2]) {
  NAV_GUARD({
    if (nu) *nu = N.nu;
    if (ka) *ka = N.ka;
    if (scale) scale[0] = N.scale[0], scale[1] = N.scale[1];
  });
}
int rp_navier_eval(rp_navier_t* h, double* nu, double* nuvol, double* re, double* div_norm, double* ekin) {
  NAV_GUARD(N.eval(nu, nuvol, re, div_norm, ekin));
}
int rp_navier_field(rp_navier_t* h, int which, rp_field_t** out) {
  NAV_GUARD({
    need(out && which >= 0 && which < 6, RP_ERR_INVALID, "bad field index");
    *out = &h->views[which];
  });
}
int rp_navier_export_eig(rp_navier_t* h, double* lam, double* q, double* p) {
  NAV_GUARD(N.solver[3]->export_eig(lam, q, p));
}
int rp_navier_launches_per_step(rp_navier_t* h, int* n) { NAV_GUARD(if (n) *n = N.launches_per_step()); }
int rp_navier_set_graph(rp_navier_t* h, int on) { NAV_GUARD(N.set_graph(on != 0)); }
int rp_navier_slab_phase1(rp_navier_t* h, int k0, int mkl, double* const* out6) {
  NAV_GUARD({
    need(out6 != nullptr && mkl > 0 && k0 >= 0 && k0 + mkl <= N.nx / 2 + 1, RP_ERR_INVALID, "slab_phase1: bad row range");
    N.slab_phase1(k0, mkl, out6);
  });
}
int rp_navier_slab_phase2(rp_navier_t* h, int j0, int nyl, const double* const* in6, double* work, double* const* out3) {
  NAV_GUARD({
    need(in6 && work && out3 && nyl > 0 && j0 >= 0 && j0 + nyl <= N.ny, RP_ERR_INVALID, "slab_phase2: bad column range");
    N.slab_phase2(j0, nyl, in6, work, out3);
  });
}
int rp_navier_slab_phase1_p2p(rp_navier_t* h, int k0, int mkl, int world, const int* joff, double* const* peers) {
  NAV_GUARD({
    need(world >= 1 && world <= 8 && joff && peers && mkl > 0 && k0 >= 0 && k0 + mkl <= N.nx / 2 + 1, RP_ERR_INVALID, "slab_phase1_p2p: bad arguments");
    N.slab_phase1(k0, mkl, nullptr, world, joff, peers);
  });
}
int rp_navier_slab_phase2_p2p(rp_navier_t* h, int j0, int nyl, const double* const* in6, double* work, int world, const int* koff,
                              double* const* peers) {
  NAV_GUARD({
    need(world >= 1 && world <= 8 && koff && peers && in6 && work && nyl > 0 && j0 >= 0 && j0 + nyl <= N.ny, RP_ERR_INVALID, "slab_phase2_p2p: bad arguments");
    N.slab_phase2(j0, nyl, in6, work, nullptr, world, koff, peers);
  });
}
int rp_dev_alloc(size_t bytes, void** out) {
  return guard([&] {
    need(out != nullptr, RP_ERR_INVALID, "null out pointer");
    *out = rt::dmalloc(bytes);
  });
}
int rp_dev_free(void* p) {
  return guard([&] { rt::dfree(p); });
}
int rp_ipc_export(void* p, unsigned char handle[64]) {
  return guard([&] {
    need(p && handle, RP_ERR_INVALID, "null argument");
#ifndef RP_EMU
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hd;
    RP_CUDA_CHECK(cudaIpcGetMemHandle(&hd, p));
    memcpy(handle, &hd, 64);
#else
    throw rp::Error(RP_ERR_INVALID, "CUDA IPC is not available in the emulation");
#endif
  });
}
int rp_ipc_open(const unsigned char handle[64], void** out) {
  return guard([&] {
    need(out && handle, RP_ERR_INVALID, "null argument");
#ifndef RP_EMU
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    RP_CUDA_CHECK(cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess));
#else
    throw rp::Error(RP_ERR_INVALID, "CUDA IPC is not available in the emulation");
#endif
  });
}
int rp_ipc_close(void* p) {
  return guard([&] {
#ifndef RP_EMU
    if (p) RP_CUDA_CHECK(cudaIpcCloseMemHandle(p));
#else
    (void)p;
#endif
  });
}
int rp_navier_slab_phase3(rp_navier_t* h, int k0, int mkl, const double* const* in3) {
  NAV_GUARD({
    need(in3 != nullptr && mkl > 0 && k0 >= 0 && k0 + mkl <= N.nx / 2 + 1, RP_ERR_INVALID, "slab_phase3: bad row range");
    N.slab_phase3(k0, mkl, in3);
  });
}
int rp_navier_kernel_path(rp_navier_t* h, int* specialised, int* split_gemm) {
  NAV_GUARD({
    if (specialised) *specialised = N.uses_specialised_kernels() ? 1 : 0;
    if (split_gemm) *split_gemm = (!N.periodic && N.solver[3]->ts.split) ? 1 : 0;
  });
}
int rp_navier_profile(rp_navier_t* h, int reps, double* ms, size_t cap, int* nops) {
  NAV_GUARD({
    need(reps > 0 && ms, RP_ERR_INVALID, "bad profile arguments");
    std::vector<double> v;
    N.profile(reps, v);
    if (nops) *nops = (int)v.size();
    for (size_t i = 0; i < v.size() && i < cap; ++i) ms[i] = v[i];
  });
}
int rp_navier_op_info(rp_navier_t* h, int i, char* name, size_t name_len, double* bytes, double* flops) {
  NAV_GUARD({
    const auto& info = N.op_info();
    need(i >= 0 && i < (int)info.size(), RP_ERR_INVALID, "op index out of range");
    if (name && name_len) {
      strncpy(name, info[i].name.c_str(), name_len - 1);
      name[name_len - 1] = 0;
    }
    if (bytes) *bytes = info[i].bytes;
    if (flops) *flops = info[i].flops;
  });
}

// ---- Navier2DAdjoint ---------------------------------------------------------
int rp_adjoint_create(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic, int periodic, rp_adjoint_t** out) {
  return guard([&] {
    need(out != nullptr, RP_ERR_INVALID, "null out pointer");
    need(nx >= 8 && ny >= 8, RP_ERR_INVALID, "grid too small");
    need(ra > 0 && pr > 0 && dt > 0 && aspect > 0, RP_ERR_INVALID, "ra, pr, dt, aspect must be positive");
    init_kernels();
    auto* h = new rp_adjoint;
    h->a = new Navier2DAdjoint(nx, ny, ra, pr, dt, aspect, adiabatic != 0, periodic != 0);
    for (int i = 0; i < 8; ++i) h->views[i] = rp_field{h->a->field_by_index(i), false};
    for (int i = 0; i < 4; ++i) h->solvers[i] = rp_solver{h->a->solver_by_index(i)};
    *out = h;
  });
}
int rp_adjoint_destroy(rp_adjoint_t* h) {
  return guard([&] {
    if (!h) return;
    delete h->a;
    delete h;
  });
}
#define ADJ_GUARD(...)                       \
  return guard([&] {                         \
    need(h, RP_ERR_INVALID, "null handle");  \
    Navier2DAdjoint& A = *h->a;              \
    (void)A;                                 \
    __VA_ARGS__;                             \
  })
int rp_adjoint_set_velocity(rp_adjoint_t* h, double amp, double m, double n) { ADJ_GUARD(A.set_velocity(amp, m, n)); }
int rp_adjoint_set_temperature(rp_adjoint_t* h, double amp, double m, double n) { ADJ_GUARD(A.set_temperature(amp, m, n)); }
int rp_adjoint_update(rp_adjoint_t* h, int nsteps) {
  ADJ_GUARD({
    need(nsteps >= 0, RP_ERR_INVALID, "nsteps < 0");
    A.update(nsteps);
  });
}
int rp_adjoint_get_time(rp_adjoint_t* h, double* t) { ADJ_GUARD(if (t) *t = A.time); }
int rp_adjoint_reset_time(rp_adjoint_t* h) {
  ADJ_GUARD({
    A.time = 0.0;
    A.navier->time = 0.0;
  });
}
int rp_adjoint_eval(rp_adjoint_t* h, double* nu, double* nuvol, double* re, double* div_norm) { ADJ_GUARD(A.eval(nu, nuvol, re, div_norm)); }
int rp_adjoint_residuals(rp_adjoint_t* h, double smooth[3], double unsmooth[3]) { ADJ_GUARD(A.residuals(smooth, unsmooth)); }
int rp_adjoint_exit(rp_adjoint_t* h, int* stop) {
  ADJ_GUARD({
    const bool e = A.exit();
    if (stop) *stop = e ? 1 : 0;
  });
}
int rp_adjoint_field(rp_adjoint_t* h, int which, rp_field_t** out) {
  ADJ_GUARD({
    need(out && which >= 0 && which < 8, RP_ERR_INVALID, "bad field index");
    *out = &h->views[which];
  });
}
int rp_adjoint_solver(rp_adjoint_t* h, int which, rp_solver_t** out) {
  ADJ_GUARD({
    need(out && which >= 0 && which < 4, RP_ERR_INVALID, "bad solver index");
    *out = &h->solvers[which];
  });
}

}  // extern "C"
#pragma GCC visibility pop
