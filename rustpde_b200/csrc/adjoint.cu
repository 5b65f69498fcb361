// adjoint.cu -- Navier2DAdjoint (src/navier/navier_adjoint.rs:128-1068) on the device: steady-state adjoint descent
// (Farazmand 2016).  Every step runs one update() of an inner Navier2D (the residual, 739-766), three full Hholtz
// fast-diagonalisation solves (the smoothers: six DMMA GEMMs per step in the confined case) and a pressure Poisson
// solve.  All arrays stay in HBM; the step is composed of the device Field2 operations (field.cu), the solver passes
// (solver.cu; real data on the specialised kernels + parity-split GEMMs) and the elementwise kernels of kernels.cu,
// in the reference's op order.
#include <cmath>

#include "model.h"

namespace rp {

static void copy_arr(Arr& dst, const Arr& src, cudaStream_t s) {
  if (dst.buf.bytes != src.buf.bytes) throw Error(RP_ERR_INTERNAL, "adjoint: shape mismatch in copy");
  rt::d2d(dst.buf.p, src.buf.p, src.buf.bytes, s);
}
// dst = (a + b * s1) * s0 on arrays of identical shape (complex arrays as interleaved doubles)
static void lin(Arr& dst, const Arr& a, const Arr& b, double s0, double s1, cudaStream_t s) {
  const int rc = dst.cplx ? 2 : 1;
  launch_combine(dst.d(), a.d(), b.d(), nullptr, dst.ld * rc, dst.rows, dst.cols * rc, s0, s1, s);
}
static void axpy(Arr& dst, double a, const Arr& src, cudaStream_t s) { lin(dst, dst, src, 1.0, a, s); }

Navier2DAdjoint::Navier2DAdjoint(int nx_, int ny_, double ra_, double pr_, double dt_, double aspect, bool adiabatic, bool periodic_)
    : nx(nx_), ny(ny_), periodic(periodic_), ra(ra_), pr(pr_), dt(dt_) {
  scale[0] = aspect, scale[1] = 1.0;
  dt_navier = 1e-2;  // navier_adjoint.rs:231
  navier.reset(new Navier2D(nx, ny, ra, pr, dt_navier, aspect, adiabatic, periodic, nullptr));
  nu = navier->nu, ka = navier->ka;
  auto B = [&](int kind, int n) { return get_base(kind, n); };
  const int kxu = periodic ? BASE_FOURIER_R2C : BASE_CHEB_DIRICHLET;
  const int kxt = periodic ? BASE_FOURIER_R2C : (adiabatic ? BASE_CHEB_NEUMANN : BASE_CHEB_DIRICHLET);
  const int kxo = periodic ? BASE_FOURIER_R2C : BASE_CHEBYSHEV;
  const int kxn = periodic ? BASE_FOURIER_R2C : BASE_CHEB_NEUMANN;
  for (int k = 0; k < 2; ++k) {  // [adjoint field, Navier-Stokes residual]  (:205-226)
    ux[k].reset(new Field2(Space2{B(kxu, nx), B(BASE_CHEB_DIRICHLET, ny)}));
    uy[k].reset(new Field2(Space2{B(kxu, nx), B(BASE_CHEB_DIRICHLET, ny)}));
    temp[k].reset(new Field2(Space2{B(kxt, nx), B(BASE_CHEB_DIRICHLET, ny)}));
  }
  pres[0].reset(new Field2(Space2{B(kxo, nx), B(BASE_CHEBYSHEV, ny)}));
  pres[1].reset(new Field2(Space2{B(kxn, nx), B(BASE_CHEB_NEUMANN, ny)}));
  field.reset(new Field2(Space2{B(kxo, nx), B(BASE_CHEBYSHEV, ny)}));
  const double sx2 = scale[0] * scale[0], sy2 = scale[1] * scale[1];
  solver_pres.reset(new Solver2(SOLVER_POISSON, pres[1]->sp, 1.0 / sx2, 1.0 / sy2, 0.0, nullptr));
  const double w = 1e0;  // weight_laplacian (:251): smoother = (1 - w D2)^-1
  smoother[0].reset(new Solver2(SOLVER_HHOLTZ, ux[1]->sp, w / sx2, w / sy2, 1.0, nullptr));
  smoother[1].reset(new Solver2(SOLVER_HHOLTZ, uy[1]->sp, w / sx2, w / sy2, 1.0, nullptr));
  smoother[2].reset(new Solver2(SOLVER_HHOLTZ, temp[1]->sp, w / sx2, w / sy2, 1.0, nullptr));
  const int ox = field->o0;
  rhs_.alloc(ox, ny, periodic);
  for (auto& u : unsm_) u.alloc(ox, ny, periodic);
  for (auto& p : phys_) p.alloc(nx, ny, false);
  conv_.alloc(nx, ny, false);
  bcv_.alloc(nx, ny, false);
  old_.alloc(ux[0]->m0, ux[0]->m1, periodic);
  red_ = DevBuf(sizeof(double) * RP_WSUM_DOUBLES);
  // physical boundary field fieldbc.v (navier.rs:314-332): backward of its ortho coefficients
  copy_arr(field->vhat, navier->tempbc_ortho(), stream);
  field->backward();
  copy_arr(bcv_, field->v, stream);
  rt::sync(stream);
}
Navier2DAdjoint::~Navier2DAdjoint() {}

void Navier2DAdjoint::set_velocity(double amp, double m, double n) {  // :994-999
  navier->apply_ic(*ux[0], amp, m, n, true);
  navier->apply_ic(*uy[0], -amp, m, n, false);
}
void Navier2DAdjoint::set_temperature(double amp, double m, double n) { navier->apply_ic(*temp[0], -amp, m, n, false); }  // :1001-1003

// conv += u * backward(gradient(f))   (conv_term.rs:22-42)
void Navier2DAdjoint::conv_term(Field2& f, const Arr& u, int d0, int d1) {
  f.gradient(d0, d1, scale);
  copy_arr(field->vhat, f.ortho, stream);
  field->backward();
  launch_combine(conv_.d(), conv_.d(), u.d(), field->v.d(), conv_.ld, nx, ny, 1.0, 1.0, stream);
}
// field.v = conv; forward; dealias (:573-579) -> field.vhat
void Navier2DAdjoint::finish_conv() {
  copy_arr(field->v, conv_, stream);
  field->forward();
  if (dealias) {
    Arr& a = field->vhat;
    launch_dealias(a.d(), a.ld, a.rows, a.cols, a.cplx ? 2 : 1, (a.rows * 2) / 3, (a.cols * 2) / 3, stream);  // navier.rs:1022-1032
  }
}
void Navier2DAdjoint::conv_u(int comp) {  // :547-611 (comp 0: ux, 1: uy)
  const Arr &uxp = phys_[0], &uyp = phys_[1], &tp = phys_[2];
  Field2& own = comp == 0 ? *ux[1] : *uy[1];
  const int d0 = comp == 0 ? 1 : 0, d1 = 1 - d0;
  conv_.zero(stream);
  conv_term(own, uxp, 1, 0);
  conv_term(own, uyp, 0, 1);
  conv_term(*ux[1], uxp, d0, d1);  // + adjoint contributions
  conv_term(*uy[1], uyp, d0, d1);
  conv_term(*temp[1], tp, d0, d1);
  conv_term(*temp[1], bcv_, d0, d1);
  finish_conv();
}
void Navier2DAdjoint::solve_u(int comp) {  // :632-674
  Field2& u0 = comp == 0 ? *ux[0] : *uy[0];
  Field2& u1 = comp == 0 ? *ux[1] : *uy[1];
  u0.to_ortho();
  copy_arr(rhs_, u0.ortho, stream);
  pres[0]->gradient(comp == 0 ? 1 : 0, comp == 0 ? 0 : 1, scale);
  axpy(rhs_, -dt, pres[0]->ortho, stream);
  conv_u(comp);
  axpy(rhs_, dt, field->vhat, stream);
  u1.gradient(2, 0, scale);
  axpy(rhs_, dt * nu, u1.ortho, stream);
  u1.gradient(0, 2, scale);
  axpy(rhs_, dt * nu, u1.ortho, stream);
  copy_arr(u0.ortho, rhs_, stream);
  u0.from_ortho();
}
void Navier2DAdjoint::solve_temp() {  // :676-697
  temp[0]->to_ortho();
  copy_arr(rhs_, temp[0]->ortho, stream);
  conv_.zero(stream);
  conv_term(*temp[1], phys_[0], 1, 0);
  conv_term(*temp[1], phys_[1], 0, 1);
  finish_conv();
  axpy(rhs_, dt, field->vhat, stream);
  uy[1]->to_ortho();  // + buoyancy (adjoint)
  axpy(rhs_, dt, uy[1]->ortho, stream);
  temp[1]->gradient(2, 0, scale);
  axpy(rhs_, dt * ka, temp[1]->ortho, stream);
  temp[1]->gradient(0, 2, scale);
  axpy(rhs_, dt * ka, temp[1]->ortho, stream);
  copy_arr(temp[0]->ortho, rhs_, stream);
  temp[0]->from_ortho();
}
void Navier2DAdjoint::divergence_to_rhs() {  // :714-719
  ux[0]->gradient(1, 0, scale);
  uy[0]->gradient(0, 1, scale);
  lin(rhs_, ux[0]->ortho, uy[0]->ortho, 1.0, 1.0, stream);
}
void Navier2DAdjoint::update_residual() {  // :739-766
  Field2* mine[3] = {ux[0].get(), uy[0].get(), temp[0].get()};
  Field2* theirs[3] = {navier->ux.get(), navier->uy.get(), navier->temp.get()};
  Field2* res[3] = {ux[1].get(), uy[1].get(), temp[1].get()};
  for (int i = 0; i < 3; ++i) {
    copy_arr(theirs[i]->vhat, mine[i]->vhat, stream);
    ++theirs[i]->vhat_version;
  }
  navier->update(1);
  for (int i = 0; i < 3; ++i) {
    lin(theirs[i]->vhat, theirs[i]->vhat, mine[i]->vhat, 1.0 / dt_navier, -1.0, stream);  // (new - old) / dt
    ++theirs[i]->vhat_version;
    theirs[i]->stream = stream;
    theirs[i]->to_ortho();
    copy_arr(unsm_[i], theirs[i]->ortho, stream);
    smoother[i]->solve_dev(unsm_[i], res[i]->vhat, periodic);
    lin(res[i]->vhat, res[i]->vhat, res[i]->vhat, -1.0, 0.0, stream);  // rescale by -1
  }
}

void Navier2DAdjoint::update(int nsteps) {  // :778-805
  for (int s = 0; s < nsteps; ++s) {
    ux[0]->backward();
    uy[0]->backward();
    temp[0]->backward();
    copy_arr(phys_[0], ux[0]->v, stream);
    copy_arr(phys_[1], uy[0]->v, stream);
    copy_arr(phys_[2], temp[0]->v, stream);
    update_residual();
    solve_u(0);
    solve_u(1);
    divergence_to_rhs();
    solver_pres->solve_dev(rhs_, pres[1]->vhat, periodic);  // :726-730
    launch_zero_elems(pres[1]->vhat.d(), periodic ? 2 : 1, stream);
    for (int comp = 0; comp < 2; ++comp) {  // project_velocity(1.0) (:699-712)
      Field2& u0 = comp == 0 ? *ux[0] : *uy[0];
      pres[1]->gradient(comp == 0 ? 1 : 0, comp == 0 ? 0 : 1, scale);
      copy_arr(old_, u0.vhat, stream);
      copy_arr(u0.ortho, pres[1]->ortho, stream);
      u0.from_ortho();
      lin(u0.vhat, old_, u0.vhat, 1.0, -1.0, stream);
    }
    pres[1]->to_ortho();  // update_pres (:732-737): the -nu div term is commented out in the reference
    axpy(pres[0]->vhat, 1.0 / dt, pres[1]->ortho, stream);
    solve_temp();
    time += dt;
  }
}

double Navier2DAdjoint::norm_l2(const Arr& a) {  // :916-926
  const int rc = a.cplx ? 2 : 1;
  launch_wsum(a.d(), nullptr, a.ld * rc, a.rows, a.cols * rc, nullptr, nullptr, 3, red_.as<double>(), stream);
  double r = 0.0;
  rt::d2h(&r, red_.p, 8, stream);
  rt::sync(stream);
  return std::sqrt(r);
}
void Navier2DAdjoint::residuals(double smooth[3], double unsmooth[3]) {
  Field2* res[3] = {ux[1].get(), uy[1].get(), temp[1].get()};
  for (int i = 0; i < 3; ++i) {
    if (smooth) smooth[i] = norm_l2(res[i]->vhat);
    if (unsmooth) unsmooth[i] = norm_l2(unsm_[i]);
  }
}
// eval_nu / eval_nuvol / eval_re (:952-992) and |div| of the adjoint fields: functions.rs through the inner solver's
// diagnostics (its ux / uy / temp are scratch between steps: update_residual overwrites them)
void Navier2DAdjoint::eval(double* o_nu, double* o_nuvol, double* o_re, double* o_div) {
  copy_arr(navier->ux->vhat, ux[0]->vhat, stream);
  copy_arr(navier->uy->vhat, uy[0]->vhat, stream);
  copy_arr(navier->temp->vhat, temp[0]->vhat, stream);
  navier->eval(o_nu, o_nuvol, o_re, o_div, nullptr);
}
bool Navier2DAdjoint::exit() {  // :892-910
  double d = 0.0;
  eval(nullptr, nullptr, nullptr, &d);
  if (std::isnan(d)) return true;
  double sm[3];
  residuals(sm, nullptr);
  return sm[0] + sm[1] + sm[2] < res_tol;
}
Field2* Navier2DAdjoint::field_by_index(int which) {
  switch (which) {
    case 0: return temp[0].get();
    case 1: return ux[0].get();
    case 2: return uy[0].get();
    case 3: return pres[0].get();
    case 4: return pres[1].get();
    case 5: return temp[1].get();
    case 6: return ux[1].get();
    case 7: return uy[1].get();
    default: throw Error(RP_ERR_INVALID, "field index out of range");
  }
}
Solver2* Navier2DAdjoint::solver_by_index(int which) {
  switch (which) {
    case 0: return smoother[0].get();
    case 1: return smoother[2].get();
    case 2: return solver_pres.get();
    case 3: return navier->solver[3].get();
    default: throw Error(RP_ERR_INVALID, "solver index out of range");
  }
}

}  // namespace rp
