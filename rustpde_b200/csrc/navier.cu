// navier.cu -- Navier2D::update (src/navier/navier.rs:737-765) on the device.
//
// One IMEX-Euler step of 2-D Rayleigh-Benard convection:
//   explicit: convection (conv_term.rs:22-42, navier.rs:538-616), buoyancy, grad p
//   implicit: diffusion, HholtzAdi (confined) / Hholtz (periodic)  (navier.rs:622-674)
//   projection: Poisson for the pseudo pressure, u -= grad phi, p update (683-721)
// All explicit terms depend on start-of-step state only, so the three
// convection terms are evaluated up-front in one batch:
//   x-backward (value and d/dx of ux, uy, T) -> y-backward / products /
//   y-forward -> x-forward fused with the x half of the implicit solves
//   -> y half of the solves -> divergence -> Poisson -> projection.
// Operators on different axes commute, which is what lets every pass be a
// single sweep of strided (x) or contiguous (y) lanes.
#include <cmath>

#include "model.h"

namespace rp {

static int half_up(int n) { return (n + 1) / 2; }

static double get_nu(double ra, double pr, double h) { return std::sqrt(pr / (ra / std::pow(h, 3.0))); }   // navier.rs:46-49
static double get_ka(double ra, double pr, double h) { return std::sqrt(1.0 / ((ra / std::pow(h, 3.0)) * pr)); }  // :52-55

Navier2D::Navier2D(int nx_, int ny_, double ra_, double pr_, double dt_, double aspect, bool adiabatic_, bool periodic_,
                   const EigData* eig)
    : nx(nx_), ny(ny_), periodic(periodic_), adiabatic(adiabatic_), ra(ra_), pr(pr_), dt(dt_), time(0.0) {
  scale[0] = aspect;
  scale[1] = 1.0;
  nu = get_nu(ra, pr, scale[1] * 2.0);
  ka = get_ka(ra, pr, scale[1] * 2.0);
  auto B = [&](int kind, int n) { return get_base(kind, n); };
  const int kx_u = periodic ? BASE_FOURIER_R2C : BASE_CHEB_DIRICHLET;
  const int kx_t = periodic ? BASE_FOURIER_R2C : (adiabatic ? BASE_CHEB_NEUMANN : BASE_CHEB_DIRICHLET);
  const int kx_o = periodic ? BASE_FOURIER_R2C : BASE_CHEBYSHEV;
  const int kx_n = periodic ? BASE_FOURIER_R2C : BASE_CHEB_NEUMANN;
  // navier.rs:234-248 / 398-408
  ux.reset(new Field2(Space2{B(kx_u, nx), B(BASE_CHEB_DIRICHLET, ny)}));
  uy.reset(new Field2(Space2{B(kx_u, nx), B(BASE_CHEB_DIRICHLET, ny)}));
  temp.reset(new Field2(Space2{B(kx_t, nx), B(BASE_CHEB_DIRICHLET, ny)}));
  pres0.reset(new Field2(Space2{B(kx_o, nx), B(BASE_CHEBYSHEV, ny)}));
  pres1.reset(new Field2(Space2{B(kx_n, nx), B(BASE_CHEB_NEUMANN, ny)}));
  field.reset(new Field2(Space2{B(kx_o, nx), B(BASE_CHEBYSHEV, ny)}));
  // navier.rs:250-266 / 410-426
  const double sx2 = std::pow(scale[0], 2.0), sy2 = std::pow(scale[1], 2.0);
  const int hk = periodic ? SOLVER_HHOLTZ : SOLVER_HHOLTZ_ADI;
  solver[0].reset(new Solver2(hk, ux->sp, dt * nu / sx2, dt * nu / sy2, 1.0, nullptr));
  solver[1].reset(new Solver2(hk, uy->sp, dt * nu / sx2, dt * nu / sy2, 1.0, nullptr));
  solver[2].reset(new Solver2(hk, temp->sp, dt * ka / sx2, dt * ka / sy2, 1.0, nullptr));
  solver[3].reset(new Solver2(SOLVER_POISSON, pres1->sp, 1.0 / sx2, 1.0 / sy2, 0.0, eig));
  // navier.rs:502-514 _scale
  for (Field2* f : {temp.get(), ux.get(), uy.get(), pres0.get()})
    for (int a = 0; a < 2; ++a) {
      for (auto& v : f->x[a]) v *= scale[a];
      for (auto& v : f->dx[a]) v *= scale[a];
    }
  // work arrays
  const int mx = periodic ? nx / 2 + 1 : nx - 2, my = ny - 2;
  const int ox = periodic ? mx : nx;
  for (int f = 0; f < 3; ++f) {
    ax_[f].alloc(nx, my, false);
    adx_[f].alloc(nx, my, false);
    bconv_[f].alloc(nx, ny, false);
    chat_[f].alloc(ox, ny, periodic);
    w_[f].alloc(mx, ny, periodic);
  }
  for (int i = 0; i < 8; ++i) phys_[i].alloc(nx, ny, false);
  vx_.alloc(mx, ny, periodic);
  ey_.alloc(mx, ny, periodic);
  div_.alloc(ox, ny, periodic);
  r1_.alloc(mx, ny, periodic);
  g_.alloc(mx, ny, periodic);
  h_.alloc(mx, my, periodic);
  a1_.alloc(mx, my, periodic);
  a2_.alloc(mx, my, periodic);
  dyp_.alloc(ox, ny, periodic);
  tbc_ortho_.alloc(ox, ny, periodic);
  dxtbc_.alloc(nx, ny, false);
  dytbc_.alloc(nx, ny, false);
  bcdiff_.alloc(ox, ny, periodic);
  if (!periodic) {
    dxp_.alloc(nx, ny, false);
    xs_p_.alloc(nx, my, false);
    xs_d_.alloc(nx, my, false);
  }
  red_ = DevBuf(sizeof(double) * RP_WSUM_DOUBLES);
  // navier.rs:301 / 461: Rayleigh-Benard boundary field, T = +0.5 at y=-1, -0.5 at y=+1
  // (bc_rbc 314-332, bc_rbc_periodic 474-492): in the ortho basis only T_1(y) is present.
  std::vector<double> tb((size_t)ox * ny * (periodic ? 2 : 1), 0.0);
  const double amp = periodic ? -0.5 * (double)nx : -0.5;  // r2c forward is unnormalised
  tb[(size_t)1 * (periodic ? 2 : 1)] = amp;
  set_tempbc_ortho(tb.data());
  // navier.rs:304 random_disturbance(0.1) uses an unseeded RNG -> fields start at
  // zero here; callers set deterministic ICs (set_velocity / set_temperature / upload).
}

Navier2D::~Navier2D() {
#ifndef RP_EMU
  if (graph_) cudaGraphExecDestroy(graph_);
  if (copy_stream_) {
    cudaStreamSynchronize(copy_stream_);
    cudaStreamDestroy(copy_stream_);
    cudaEventDestroy(ev_staged_);
    cudaEventDestroy(ev_consumed_);
  }
  if (fetch_stream_) {
    cudaStreamSynchronize(fetch_stream_);
    cudaStreamDestroy(fetch_stream_);
    cudaEventDestroy(ev_fetch_ready_);
    cudaEventDestroy(ev_fetched_);
  }
  if (div_host_) {
    cudaEventDestroy(ev_div_);
    cudaFreeHost(div_host_);
  }
#endif
}

void Navier2D::stage_state(const double* t, const double* u, const double* v, const double* p) {
  Field2* dst[4] = {temp.get(), ux.get(), uy.get(), pres0.get()};
  const double* src[4] = {t, u, v, p};
  cudaStream_t cs = stream;
#ifndef RP_EMU
  if (!copy_stream_) {
    RP_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    RP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_staged_, cudaEventDisableTiming));
    RP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_consumed_, cudaEventDisableTiming));
  } else {
    // the staging arrays may still be read by the previous commit
    RP_CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, ev_consumed_, 0));
  }
  cs = copy_stream_;
#endif
  for (int i = 0; i < 4; ++i) {
    if (!stage_[i].buf.p) stage_[i].alloc(dst[i]->vhat.rows, dst[i]->vhat.cols, dst[i]->vhat.cplx);
    stage_[i].upload(src[i], cs);
  }
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaEventRecord(ev_staged_, copy_stream_));
#endif
  staged_ = true;
}

void Navier2D::commit_staged() {
  if (!staged_) throw Error(RP_ERR_INVALID, "commit_staged: no staged state");
  Field2* dst[4] = {temp.get(), ux.get(), uy.get(), pres0.get()};
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_staged_, 0));
#endif
  for (int i = 0; i < 4; ++i) {
    rt::d2d(dst[i]->vhat.buf.p, stage_[i].buf.p, stage_[i].buf.bytes, stream);
    ++dst[i]->vhat_version;
  }
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaEventRecord(ev_consumed_, stream));
#endif
  staged_ = false;
}

// Asynchronous download of the state (the vhat arrays of temp, ux, uy, pres[0]) into caller-owned host buffers on a
// second copy stream: the snapshot is taken device-to-device on the compute stream (so later update()s may overwrite
// the live arrays), the device-to-host copies then overlap those updates.  fetch_wait() blocks until they landed.
void Navier2D::fetch_state(double* t, double* u, double* v, double* p) {
  Field2* src[4] = {temp.get(), ux.get(), uy.get(), pres0.get()};
  double* dst[4] = {t, u, v, p};
  cudaStream_t cs = stream;
#ifndef RP_EMU
  if (!fetch_stream_) {
    RP_CUDA_CHECK(cudaStreamCreateWithFlags(&fetch_stream_, cudaStreamNonBlocking));
    RP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fetch_ready_, cudaEventDisableTiming));
    RP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fetched_, cudaEventDisableTiming));
  } else {
    RP_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_fetched_, 0));  // the snapshot arrays may still be read by the last fetch
  }
#endif
  for (int i = 0; i < 4; ++i) {
    if (!snap_[i].buf.p) snap_[i].alloc(src[i]->vhat.rows, src[i]->vhat.cols, src[i]->vhat.cplx);
    rt::d2d(snap_[i].buf.p, src[i]->vhat.buf.p, snap_[i].buf.bytes, stream);
  }
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaEventRecord(ev_fetch_ready_, stream));
  RP_CUDA_CHECK(cudaStreamWaitEvent(fetch_stream_, ev_fetch_ready_, 0));
  cs = fetch_stream_;
#endif
  for (int i = 0; i < 4; ++i) snap_[i].download(dst[i], cs);
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaEventRecord(ev_fetched_, fetch_stream_));
#endif
}
void Navier2D::fetch_wait() {
#ifndef RP_EMU
  if (fetch_stream_) RP_CUDA_CHECK(cudaStreamSynchronize(fetch_stream_));
#endif
}

Field2* Navier2D::field_by_index(int which) {
  switch (which) {
    case 0: return temp.get();
    case 1: return ux.get();
    case 2: return uy.get();
    case 3: return pres0.get();
    case 4: return pres1.get();
    case 5: return field.get();
    default: throw Error(RP_ERR_INVALID, "field index out of range");
  }
}

void Navier2D::sync() { rt::sync(stream); }

static void copy_arr(Arr& dst, const Arr& src, cudaStream_t s) {
  if (dst.buf.bytes != src.buf.bytes) throw Error(RP_ERR_INTERNAL, "copy_arr: shape mismatch");
  rt::d2d(dst.buf.p, src.buf.p, src.buf.bytes, s);
}

void Navier2D::set_tempbc_ortho(const double* that_bc) {
  tbc_ortho_.upload(that_bc, stream);
  rebuild_bc();
}

// navier.solid = Some([mask, value]) (navier.rs:191; solid_masks.rs): must be set before the first update()
void Navier2D::set_solid(const double* mask, const double* value) {
  if (!ops_.empty()) throw Error(RP_ERR_INVALID, "set_solid must be called before the first update()");
  has_solid_ = mask != nullptr;
  if (!has_solid_) return;
  if (!solid_mask_.buf.p) {
    solid_mask_.alloc(nx, ny, false);
    solid_val_.alloc(nx, ny, false);
    phys_t_.alloc(nx, ny, false);
    tbc_phys_.alloc(nx, ny, false);
  }
  solid_mask_.upload(mask, stream);
  if (value)
    solid_val_.upload(value, stream);
  else
    solid_val_.zero(stream);
  // physical boundary field (fieldbc.v): backward of its ortho coefficients through the work field
  field->stream = stream;
  copy_bc_to_field();
  field->backward();
  copy_arr(tbc_phys_, field->v, stream);
  rt::sync(stream);
}
// penalisation term of field f (0 ux, 1 uy, 2 temp) in the fused product kernel
void Navier2D::set_solid_args(fk::YConvArgs& a, int f) {
  const fk::Mat none{nullptr, 0, 0, 0};
  a.mask = a.sval = a.w = a.wbc = none;
  a.ieta = 1.0 / 1e-2;  // eta = 1e-2 (navier.rs:553)
  if (!has_solid_) return;
  auto m = [](const Arr& x) { return fk::Mat{x.d(), x.ld, x.rows, x.cols}; };
  a.mask = m(solid_mask_);
  a.sval = f == 2 ? m(solid_val_) : m(zero_phys());
  a.w = f == 0 ? m(phys_[0]) : (f == 1 ? m(phys_[1]) : m(phys_t_));
  if (f == 2) a.wbc = m(tbc_phys_);
}
const Arr& Navier2D::zero_phys() {
  if (!zero_phys_.buf.p) {
    zero_phys_.alloc(nx, ny, false);
    zero_phys_.zero(stream);
  }
  return zero_phys_;
}

void Navier2D::copy_bc_to_field() { copy_arr(field->vhat, tbc_ortho_, stream); }

// Time-invariant pieces of the boundary field (navier.rs:547-550, 665-668)
void Navier2D::rebuild_bc() {
  Field2& f = *field;
  auto grad_to = [&](int ddx, int ddy) {
    copy_arr(f.vhat, tbc_ortho_, stream);
    f.gradient(ddx, ddy, scale);
  };
  grad_to(1, 0);
  copy_arr(f.vhat, f.ortho, stream);
  f.backward();
  copy_arr(dxtbc_, f.v, stream);
  grad_to(0, 1);
  copy_arr(f.vhat, f.ortho, stream);
  f.backward();
  copy_arr(dytbc_, f.v, stream);
  const int rc = f.cplx ? 2 : 1;
  grad_to(2, 0);
  launch_combine(bcdiff_.d(), f.ortho.d(), nullptr, nullptr, f.ortho.ld * rc, f.o0, f.o1 * rc, dt * ka, 0.0, stream);
  grad_to(0, 2);
  launch_combine(bcdiff_.d(), bcdiff_.d(), f.ortho.d(), nullptr, f.ortho.ld * rc, f.o0, f.o1 * rc, 1.0, dt * ka, stream);
  rt::sync(stream);
  // how much of the time-invariant arrays is exactly zero (bc_rbc: only T_1(y) of x-mode 0 is present, so tbc has one
  // non-zero row, its Laplacian none, and d/dx of the boundary field vanishes); the kernels skip what they need not read
  tbc_rows_ = bcdiff_rows_ = 1 << 30;
  dxtbc_zero_ = dytbc_zero_ = false;
  if (!periodic) {
    auto nz_rows = [&](const Arr& a) {
      std::vector<double> h((size_t)a.rows * a.cols);
      a.download(h.data(), stream);
      rt::sync(stream);
      int last = 0;
      for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j)
          if (h[(size_t)i * a.cols + j] != 0.0) last = i + 1;
      return last;
    };
    tbc_rows_ = nz_rows(tbc_ortho_);
    bcdiff_rows_ = nz_rows(bcdiff_);
    dxtbc_zero_ = nz_rows(dxtbc_) == 0;
    dytbc_zero_ = nz_rows(dytbc_) == 0;
  }
  graph_dirty_ = true;
}

// navier.rs:1035-1076: amp * sin(pi m x~) cos(pi n y~) (or cos/sin) on normalised coords, then forward()
void Navier2D::apply_ic(Field2& f, double amp, double m, double n, bool sin_cos) {
  const std::vector<double>&x = f.x[0], &y = f.x[1];
  const int nxp = f.n0, nyp = f.n1;
  std::vector<double> xs(x.size()), ys(y.size());
  for (size_t i = 0; i < x.size(); ++i) xs[i] = (x[i] - x[0]) / (x[x.size() - 1] - x[0]);
  for (size_t j = 0; j < y.size(); ++j) ys[j] = (y[j] - y[0]) / (y[y.size() - 1] - y[0]);
  const double ax = M_PI * m, ay = M_PI * n;
  std::vector<double> v((size_t)nxp * nyp);
  for (int i = 0; i < nxp; ++i)
    for (int j = 0; j < nyp; ++j)
      v[(size_t)i * nyp + j] = sin_cos ? amp * std::sin(ax * xs[i]) * std::cos(ay * ys[j])
                                       : amp * std::cos(ax * xs[i]) * std::sin(ay * ys[j]);
  f.stream = stream;
  f.v.upload(v.data(), stream);
  f.forward();
  rt::sync(stream);
}

void Navier2D::set_velocity(double amp, double m, double n) {  // navier.rs:927-930
  apply_ic(*ux, amp, m, n, true);
  apply_ic(*uy, -amp, m, n, false);
}
void Navier2D::set_temperature(double amp, double m, double n) {  // navier.rs:934-936
  apply_ic(*temp, -amp, m, n, false);
}

// --------------------------------------------------------------------------
void Navier2D::build_step() {
  if (!ops_.empty()) return;
  fk::apply_kflags();
  if (has_solid_ && !(fk::y_supported(ny) && (periodic ? fk::px_supported(nx) : fk::x_supported(nx))))
    throw Error(RP_ERR_INVALID, "solid masks are implemented on the specialised kernels only (ny = 2^k + 1, see rp_navier_kernel_path)");
  const char* nf = getenv("RUSTPDE_B200_NO_FAST");
  const bool fast_ok = !(nf && nf[0] == '1');
  if (periodic && fast_ok && fk::px_supported(nx) && fk::y_supported(ny))
    build_step_periodic_fast();
  else if (periodic)
    build_step_periodic();
  else if (fast_ok && fk::x_supported(nx) && fk::y_supported(ny))
    build_step_confined_fast();
  else
    build_step_confined();
  launches_per_step_ = (int)ops_.size();
  if (fast_ok && fast_ops_.empty() && (long long)nx * ny >= 256 * 256 && !getenv("RUSTPDE_B200_QUIET")) {
    // a performance cliff worth a line: the specialised kernels cover y lanes of 2^k + 1 points, confined x lanes of
    // 17 .. 2049 points (Bluestein, or the power-of-two DCT when nx = 2^k + 1) and periodic x lanes of 2^k points
    static bool warned = false;
    if (!warned) {
      warned = true;
      fprintf(stderr,
              "[rustpde_b200] Navier2D %dx%d %s runs on the generic lane programs (several times slower): the specialised kernels "
              "need ny = 2^k + 1 and %s\n",
              nx, ny, periodic ? "periodic" : "confined", periodic ? "nx = 2^k" : "17 <= nx <= 2049");
    }
  }
  if (getenv("RUSTPDE_B200_VERBOSE"))
    fprintf(stderr, "[rustpde_b200] Navier2D %dx%d %s: %s kernels, %d launches/step\n", nx, ny,
            periodic ? "periodic" : "confined", fast_ops_.empty() ? "lane-program" : "specialised", launches_per_step_);
}

void Navier2D::add_prog(ProgBuilder& pb, const char* name) {
  step_.push_back(pb.build());
  ops_.push_back(StepOp{0, (int)step_.size() - 1});
  opinfo_.push_back(OpInfo{name, step_.back().bytes, 0.0});
}

// y phase shared by both geometries: physical-space products (conv_term.rs:41)
// between y-backward and y-forward transforms.  Real arrays, rows = physical x.
void Navier2D::build_y_phase() {
  const Base& byu = *ux->sp.b1;
  const Base& byt = *temp->sp.b1;
  const Base& byo = *field->sp.b1;
  const int my = ny - 2;
  const Lay nat = lay_natural();
  const double isy = 1.0 / scale[1];
  Field2* flds[3] = {ux.get(), uy.get(), temp.get()};
  // phys_: 0 ux, 1 uy, 2 dxu, 3 dyu, 4 dxv, 5 dyv, 6 dxT, 7 dyT
  const int val_idx[3] = {0, 1, -1}, dx_idx[3] = {2, 4, 6}, dy_idx[3] = {3, 5, 7};
  for (int f = 0; f < 3; ++f) {
    const Base& by = (f == 2) ? byt : byu;
    (void)flds;
    ProgBuilder pb(AXIS_Y, half_up(nx));
    Lay l = lay_split(ny);
    pb.ld(0, ax_[f], my, l, 1.0, 0, 0, nullptr, ny);
    pb.toortho(0, by, l);
    Lay lv = l;
    if (val_idx[f] >= 0) {
      pb.copy(1, 0, ny, l, l);
      lv = pb.dct(1, byo, l, true);
      pb.st(1, phys_[val_idx[f]], ny, lv);
    }
    Lay ld = pb.diff(0, ny, l, 1, isy);
    ld = pb.dct(0, byo, ld, true);
    pb.st(0, phys_[dy_idx[f]], ny, ld);
    pb.ld(0, adx_[f], my, l, 1.0, 0, 0, nullptr, ny);
    pb.toortho(0, by, l);
    Lay lx = pb.dct(0, byo, l, true);
    pb.st(0, phys_[dx_idx[f]], ny, lx);
    add_prog(pb, "y_backward");
  }
  const int ny_cut = dealias ? (ny * 2) / 3 : -1;  // navier.rs:1029
  for (int f = 0; f < 3; ++f) {
    ProgBuilder pb(AXIS_Y, half_up(nx));
    pb.ld(0, phys_[0], ny, nat);
    pb.ld(1, phys_[dx_idx[f]], ny, nat);
    if (f == 2) pb.ld(1, dxtbc_, ny, nat, 1.0, LF_ACC);
    pb.mulpw(2, 0, 1, ny, nat, nat, false);
    pb.ld(0, phys_[1], ny, nat);
    pb.ld(1, phys_[dy_idx[f]], ny, nat);
    if (f == 2) pb.ld(1, dytbc_, ny, nat, 1.0, LF_ACC);
    pb.mulpw(2, 0, 1, ny, nat, nat, true);
    Lay l = pb.dct(2, byo, nat, false);
    pb.st(2, bconv_[f], ny, l, 1.0, 0, ny_cut, -1);
    add_prog(pb, "conv_y_forward");
  }
}

void Navier2D::build_step_confined() {
  const Base &bxu = *ux->sp.b0, &byu = *ux->sp.b1;
  const Base &bxt = *temp->sp.b0, &byt = *temp->sp.b1;
  const Base &bxn = *pres1->sp.b0, &byn = *pres1->sp.b1;
  const Base &bxo = *field->sp.b0, &byo = *field->sp.b1;
  const int mx = nx - 2, my = ny - 2;
  const double isx = 1.0 / scale[0], isy = 1.0 / scale[1];
  Field2* flds[3] = {ux.get(), uy.get(), temp.get()};
  const Base* bxs[3] = {&bxu, &bxu, &bxt};
  const Base* bys[3] = {&byu, &byu, &byt};
  // ---- 1. x-backward: value and d/dx of ux, uy, T ------------------------
  for (int f = 0; f < 3; ++f) {
    ProgBuilder pb(AXIS_X, half_up(my));
    Lay l = lay_split(nx);
    pb.ld(0, flds[f]->vhat, mx, l, 1.0, 0, 0, nullptr, nx);
    pb.toortho(0, *bxs[f], l);
    pb.copy(1, 0, nx, l, l);
    Lay l1 = pb.diff(1, nx, l, 1, isx);
    Lay l0 = pb.dct(0, bxo, l, true);
    l1 = pb.dct(1, bxo, l1, true);
    pb.st(0, ax_[f], nx, l0);
    pb.st(1, adx_[f], nx, l1);
    add_prog(pb, "x_backward_dct");
  }
  // ---- 2. y phase ----------------------------------------------------------
  build_y_phase();
  // ---- 3. x-forward + dealias + rhs assembly + x half of HholtzAdi ---------
  const int nx_cut = dealias ? (nx * 2) / 3 : -1;  // navier.rs:1028
  for (int f = 0; f < 3; ++f) {
    ProgBuilder pb(AXIS_X, half_up(ny));
    pb.ld(0, bconv_[f], nx, lay_natural());
    Lay l0 = pb.dct(0, bxo, lay_natural(), false);
    if (nx_cut >= 0) pb.cut(0, nx, nx_cut, l0);
    pb.scale(0, nx, -dt, l0);                                  // - dt * conv       (navier.rs:630, 651, 671)
    Lay l = lay_split(nx);
    // + to_ortho(field): S_y across lanes (coefficients d_j, l_{j-2}), S_x in the lane
    pb.ld(1, flds[f]->vhat, mx, l, 1.0, 0, 0, bys[f]->d_sd.as<double>(), nx);
    pb.ld(1, flds[f]->vhat, mx, l, 1.0, LF_ACC, -2, bys[f]->d_sl.as<double>());
    pb.toortho(1, *bxs[f], l);
    pb.axpy(0, 1, nx, 1.0, l0, l);
    if (f == 0) {  // - dt * d/dx pres / sx            (navier.rs:627)
      pb.ld(1, pres0->vhat, nx, l);
      Lay lp = pb.diff(1, nx, l, 1, -dt * isx);
      pb.axpy(0, 1, nx, 1.0, l0, lp);
    } else if (f == 1) {  // - dt * d/dy pres / sy, + dt * buoyancy  (navier.rs:646-648)
      pb.ld(0, dyp_, nx, l0, -dt, LF_ACC);
      pb.ld(1, temp->vhat, mx, l, 1.0, 0, 0, byt.d_sd.as<double>(), nx);
      pb.ld(1, temp->vhat, mx, l, 1.0, LF_ACC, -2, byt.d_sl.as<double>());
      pb.toortho(1, bxt, l);
      pb.axpy(0, 1, nx, dt, l0, l);
      pb.ld(0, tbc_ortho_, nx, l0, dt, LF_ACC);
    } else {  // + dt * ka * (dxx + dyy) fieldbc          (navier.rs:665-668)
      pb.ld(0, bcdiff_, nx, l0, 1.0, LF_ACC);
    }
    solver[f]->emit_x(pb, 0, l0);
    pb.st(0, w_[f], mx, l0);
    add_prog(pb, "x_forward_rhs_adi_x");
  }
  // ---- 4. y half of HholtzAdi (+ pieces of the divergence) -----------------
  for (int f = 0; f < 3; ++f) {
    ProgBuilder pb(AXIS_Y, half_up(mx));
    Lay l = lay_split(ny);
    pb.ld(0, w_[f], ny, l);
    solver[f]->emit_y(pb, 0, 1, l, false);
    pb.st(0, flds[f]->vhat, my, l);
    if (f == 0) {
      pb.toortho(0, byu, l);
      pb.st(0, vx_, ny, l);
    } else if (f == 1) {
      pb.toortho(0, byu, l);
      Lay ld = pb.diff(0, ny, l, 1, isy);
      pb.st(0, ey_, ny, ld);
    }
    add_prog(pb, "adi_y");
  }
  // ---- 5. divergence (navier.rs:698-703) + B2x of the Poisson rhs ----------
  {
    ProgBuilder pb(AXIS_X, half_up(ny));
    Lay l = lay_split(nx);
    pb.ld(0, vx_, mx, l, 1.0, 0, 0, nullptr, nx);
    pb.toortho(0, bxu, l);
    Lay ld = pb.diff(0, nx, l, 1, isx);
    pb.ld(1, ey_, mx, l, 1.0, 0, 0, nullptr, nx);
    pb.toortho(1, bxu, l);
    pb.axpy(0, 1, nx, 1.0, ld, l);
    pb.st(0, div_, nx, ld);
    solver[3]->emit_x(pb, 0, ld);
    pb.st(0, r1_, mx, ld);
    add_prog(pb, "divergence_b2x");
  }
  // ---- 6-8. fast diagonalisation: P., per-mode Fdma_y, Q. (poisson.rs:131-149)
  ops_.push_back(StepOp{1, 0});
  opinfo_.push_back(OpInfo{"poisson_gemm_fwd", 8.0 * ((double)mx * mx + 2.0 * mx * ny), 2.0 * mx * (double)mx * ny});
  {
    ProgBuilder pb(AXIS_Y, half_up(mx));
    Lay l = lay_split(ny);
    pb.ld(0, g_, ny, l);
    const FdmaModeDev& md = solver[3]->ts.mode;
    pb.ld(1, ArrRef(md.inv.p, md.inv_ld, md.nlanes, md.n, false), my, l);
    solver[3]->emit_y(pb, 0, 1, l, false);
    pb.st(0, h_, my, l);
    add_prog(pb, "poisson_mode_y");
  }
  ops_.push_back(StepOp{2, 0});
  opinfo_.push_back(OpInfo{"poisson_gemm_bwd", 8.0 * ((double)mx * mx + 2.0 * mx * my), 2.0 * mx * (double)mx * my});
  ops_.push_back(StepOp{3, 0});  // pres[1].vhat[[0,0]] = 0   (navier.rs:714)
  opinfo_.push_back(OpInfo{"zero_mode00", 8.0, 0.0});
  // ---- 9. projection (navier.rs:683-695): x part -----------------------------
  {
    ProgBuilder pb(AXIS_X, half_up(my));
    Lay l = lay_split(nx);
    pb.ld(0, pres1->vhat, mx, l, 1.0, 0, 0, nullptr, nx);
    pb.toortho(0, bxn, l);
    pb.copy(1, 0, nx, l, l);
    Lay ld = pb.diff(0, nx, l, 1, isx);
    pb.fromortho(0, bxu, ld);
    pb.fromortho(1, bxu, l);
    pb.st(0, a1_, mx, ld);
    pb.st(1, a2_, mx, l);
    add_prog(pb, "project_x");
  }
  // ---- 10. projection: y part; u -= from_ortho(grad phi) ---------------------
  {
    ProgBuilder pb(AXIS_Y, half_up(mx));
    Lay l = lay_split(ny);
    pb.ld(0, a1_, my, l, 1.0, 0, 0, nullptr, ny);
    pb.toortho(0, byn, l);
    pb.fromortho(0, byu, l);
    pb.st(0, ux->vhat, my, l, -1.0, LF_ACC);
    pb.ld(0, a2_, my, l, 1.0, 0, 0, nullptr, ny);
    pb.toortho(0, byn, l);
    Lay ld = pb.diff(0, ny, l, 1, isy);
    pb.fromortho(0, byu, ld);
    pb.st(0, uy->vhat, my, ld, -1.0, LF_ACC);
    add_prog(pb, "project_y");
  }
  // ---- 11. pressure update (navier.rs:717-721) + d/dy pres for the next step --
  {
    ProgBuilder pb(AXIS_Y, half_up(nx));
    Lay l = lay_split(ny);
    pb.ld(0, pres1->vhat, my, l, 1.0, 0, 0, bxn.d_sd.as<double>(), ny);
    pb.ld(0, pres1->vhat, my, l, 1.0, LF_ACC, -2, bxn.d_sl.as<double>());
    pb.toortho(0, byn, l);
    pb.scale(0, ny, 1.0 / dt, l);
    pb.ld(0, div_, ny, l, -nu, LF_ACC);
    pb.ld(0, pres0->vhat, ny, l, 1.0, LF_ACC);
    pb.st(0, pres0->vhat, ny, l);
    Lay ld = pb.diff(0, ny, l, 1, isy);
    pb.st(0, dyp_, ny, ld);
    add_prog(pb, "pressure_update");
  }
  (void)byo;
  (void)byt;
}


// --------------------------------------------------------------------------
// Confined step on the specialised kernels (fast_x.cu / fast_y.cu): same
// schedule and intermediate arrays as build_step_confined, one hand-written
// kernel per pass instead of a lane program.
// --------------------------------------------------------------------------
void Navier2D::add_fast(const char* name, double bytes, std::function<void()> fn) {
  fast_ops_.push_back(std::move(fn));
  ops_.push_back(StepOp{4, (int)fast_ops_.size() - 1});
  opinfo_.push_back(OpInfo{name, bytes, 0.0});
}

static fk::Mat mat_of(const Arr& a) { return fk::Mat{a.d(), a.ld, a.rows, a.cols}; }
static fk::DctTab dct_of(const Base& b) {
  fk::DctTab t;
  t.n = b.n;
  t.sc = b.d_sc.as<double2>();
  t.tw = b.fft.twc.as<double2>();
  t.chirp = b.fft.plan.chirp;
  t.bhat = b.fft.bhat_dr.as<double2>();
  return t;
}
static fk::B2Tabs b2_of(const Base& b) {
  return fk::B2Tabs{b.d_b2lo.as<double>(), b.d_b2di.as<double>(), b.d_b2up.as<double>()};
}
static fk::FdmaTabs fdma_of(const FdmaDev& f) {
  return fk::FdmaTabs{f.fp.as<double>(), f.bs.as<double>(), f.bp1.as<double>(), f.bp2.as<double>()};
}
// host copy of a device table of doubles
static std::vector<double> host_of(const DevBuf& b) {
  std::vector<double> v(b.bytes / sizeof(double));
  if (!v.empty()) {
    rt::d2h(v.data(), b.p, b.bytes, 0);
    rt::sync(0);
  }
  return v;
}
// from_ortho tables of base b for lanes of n points; ng = chunk groups of the kernel's first-order scans
fk::TdmaTabs Navier2D::tdma_of(const Base& b, int n, fk::ScanShape ng) {
  fk::TdmaTabs t{b.d_sd.as<double>(), b.d_sl.as<double>(), b.d_tfs.as<double>(), b.d_tfp.as<double>(), b.d_tbp.as<double>(),
                 nullptr, nullptr};
  const auto key = std::make_pair((const void*)&b, ng.ng * 1024 + ng.rows);
  auto it = perm_tdma_.find(key);
  if (it == perm_tdma_.end()) {
    const int m = n - 2;
    perm_.push_back(upload(fk::perm_table(m, true, ng, 4, {host_of(b.d_sd), host_of(b.d_sl), host_of(b.d_tfs), host_of(b.d_tfp)}, {0, 0, 0, 0})));
    const double* pf = perm_.back().as<double>();
    perm_.push_back(upload(fk::perm_table(m, false, ng, 1, {host_of(b.d_tbp)}, {0})));
    it = perm_tdma_.emplace(key, std::make_pair(pf, perm_.back().as<double>())).first;
  }
  t.pf = it->second.first, t.pb = it->second.second;
  return t;
}
void Navier2D::build_step_confined_fast() {
  const Base &bxu = *ux->sp.b0, &byu = *ux->sp.b1;
  const Base &bxt = *temp->sp.b0, &byt = *temp->sp.b1;
  const Base &bxn = *pres1->sp.b0, &byn = *pres1->sp.b1;
  const Base &bxo = *field->sp.b0, &byo = *field->sp.b1;
  const int mx = nx - 2, my = ny - 2;
  const double isx = 1.0 / scale[0], isy = 1.0 / scale[1];
  Field2* flds[3] = {ux.get(), uy.get(), temp.get()};
  const Base* bxs[3] = {&bxu, &bxu, &bxt};
  const Base* bys[3] = {&byu, &byu, &byt};
  const double fb = 8.0 * (double)nx * (double)ny;  // bytes of one field sweep
  // phys_: 0 ux, 1 uy, 2 dxu, 3 dyu, 4 dxv, 5 dyv, 6 dxT, 7 dyT
  const int val_idx[3] = {0, 1, -1}, dx_idx[3] = {2, 4, 6}, dy_idx[3] = {3, 5, 7};
  // Each group of three independent passes (ux, uy, T) is one launch (blockIdx.y = field).
  // ---- 1. x-backward: value and d/dx of ux, uy, T ------------------------
  {
    fk::XBackwardArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::XBackwardArgs& a = a3.a[f];
      a.src = mat_of(flds[f]->vhat);
      a.val = mat_of(ax_[f]);
      a.dx = mat_of(adx_[f]);
      a.sd = bxs[f]->d_sd.as<double>();
      a.sl = bxs[f]->d_sl.as<double>();
      a.isx = isx;
      a.t = dct_of(bxo);
    }
    add_fast("x_backward_dct", 9 * fb, [this, a3]() { fk::launch_x_backward(a3, 3, stream); });
  }
  // ---- 2. y-backward -> physical space ------------------------------------
  {
    fk::YBackwardArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::YBackwardArgs& a = a3.a[f];
      a.a = mat_of(ax_[f]);
      a.adx = mat_of(adx_[f]);
      a.val = val_idx[f] >= 0 ? mat_of(phys_[val_idx[f]]) : (has_solid_ ? mat_of(phys_t_) : fk::Mat{nullptr, 0, 0, 0});
      a.dy = mat_of(phys_[dy_idx[f]]);
      a.dx = mat_of(phys_[dx_idx[f]]);
      a.sd = bys[f]->d_sd.as<double>();
      a.sl = bys[f]->d_sl.as<double>();
      a.isy = isy;
      a.t = dct_of(byo);
    }
    add_fast("y_backward", 14 * fb, [this, a3]() { fk::launch_y_backward(a3, 3, stream); });
  }
  // ---- 3. products + forward DCT-y + dealias-y -----------------------------
  {
    fk::YConvArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::YConvArgs& a = a3.a[f];
      a.u = mat_of(phys_[0]);
      a.du = mat_of(phys_[dx_idx[f]]);
      a.v = mat_of(phys_[1]);
      a.dv = mat_of(phys_[dy_idx[f]]);
      a.bcx = f == 2 ? mat_of(dxtbc_) : fk::Mat{nullptr, 0, 0, 0};
      a.bcy = f == 2 ? mat_of(dytbc_) : fk::Mat{nullptr, 0, 0, 0};
      set_solid_args(a, f);
      a.out = mat_of(bconv_[f]);
      a.cut = dealias ? (ny * 2) / 3 : ny;  // navier.rs:1029
      a.t = dct_of(byo);
    }
    add_fast("conv_y_forward", 17 * fb, [this, a3]() {
      fk::YConvArgs3 b3 = a3;  // identically zero boundary gradients are not read
      if (dxtbc_zero_) b3.a[2].bcx = fk::Mat{nullptr, 0, 0, 0};
      if (dytbc_zero_) b3.a[2].bcy = fk::Mat{nullptr, 0, 0, 0};
      fk::launch_y_conv(b3, 3, stream);
    });
  }
  // ---- 4. x-forward + dealias + rhs assembly + x half of HholtzAdi ---------
  // RUSTPDE_B200_XS=1: rhs assembly / x sweeps / divergence / projection as streaming column scans (fast_xs.cu) instead
  // of the tile kernels -- measured equal within 2 % at 2048 x 2049, so the tile kernels stay the default
  const char* nxs = getenv("RUSTPDE_B200_XS");
  const bool use_xs = fk::xs_supported(nx) && nxs && nxs[0] == '1';
  // RUSTPDE_B200_XW=0: keep the banded x sweeps inside the tile kernels (round-1 schedule); default: warp-serial
  // column sweeps (fast_xw.cu)
  const char* nxw = getenv("RUSTPDE_B200_XW");
  const bool use_xw = !use_xs && fk::xw_supported(nx) && !(nxw && nxw[0] == '0');
  // RUSTPDE_B200_XW=2: divergence and projection as warp-serial sweeps too.  Parity-green but slower (one warp per
  // strip and ~40 instructions per chain step: 0.190 / 0.233 ms against 0.083 / 0.109 ms of the tile kernels), so opt-in
  const bool use_xw2 = use_xw && nxw && nxw[0] == '2';
  if (use_xs) {
    // forward DCT-x (tile kernel) -> chat = -dt * cut(F_x conv); then rhs assembly + B2_x + Fdma_x as streaming column scans
    fk::XFdctArgs3 d3;
    fk::XsRhsAdiArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::XFdctArgs& d = d3.a[f];
      d.conv = mat_of(bconv_[f]);
      d.out = mat_of(chat_[f]);
      d.cut = dealias ? (nx * 2) / 3 : nx;  // navier.rs:1028
      d.scale = -dt;
      d.t = dct_of(bxo);
      fk::XsRhsAdiArgs& a = a3.a[f];
      a.chat = mat_of(chat_[f]);
      a.rhs = mat_of(chat_[f]);  // in place
      a.out = mat_of(w_[f]);
      a.fld = mat_of(flds[f]->vhat);
      a.fxsd = bxs[f]->d_sd.as<double>(), a.fxsl = bxs[f]->d_sl.as<double>();
      a.fysd = bys[f]->d_sd.as<double>(), a.fysl = bys[f]->d_sl.as<double>();
      a.mode = f;
      a.dxp = mat_of(dxp_);
      a.dyp = mat_of(dyp_);
      a.tmp = mat_of(temp->vhat);
      a.tbc = mat_of(tbc_ortho_);
      a.bcdiff = mat_of(bcdiff_);
      a.txsd = bxt.d_sd.as<double>(), a.txsl = bxt.d_sl.as<double>();
      a.tysd = byt.d_sd.as<double>(), a.tysl = byt.d_sl.as<double>();
      a.dt = dt;
      a.nx = nx;
      const FdmaDev& fd = solver[f]->adi[0].fdma;
      const int m = nx - 2;
      const fk::ScanShape sh = fk::xs_scan_shape();
      perm_.push_back(upload(fk::perm_table(m, true, sh, 4, {host_of(bxo.d_b2lo), host_of(bxo.d_b2di), host_of(bxo.d_b2up), host_of(fd.fp)}, {0, 0, 0, 0})));
      a.pt1 = perm_.back().as<double>();
      perm_.push_back(upload(fk::perm_table(m, false, sh, 4, {host_of(fd.bs), host_of(fd.bp1), host_of(fd.bp2)}, {0, 0, 0})));
      a.pt2 = perm_.back().as<double>();
    }
    add_fast("x_forward_dct", 6 * fb, [this, d3]() { fk::launch_x_fdct(d3, 3, stream); });
    add_fast("rhs_adi_x", 14 * fb, [this, a3]() { fk::launch_xs_rhs_adi(a3, 3, stream); });
  } else {
    fk::XForwardArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::XForwardArgs& a = a3.a[f];
      a.conv = mat_of(bconv_[f]);
      a.out = mat_of(w_[f]);
      a.cut = dealias ? (nx * 2) / 3 : nx;  // navier.rs:1028
      a.dt = dt;
      a.fld = mat_of(flds[f]->vhat);
      a.fxsd = bxs[f]->d_sd.as<double>(), a.fxsl = bxs[f]->d_sl.as<double>();
      a.fysd = bys[f]->d_sd.as<double>(), a.fysl = bys[f]->d_sl.as<double>();
      a.mode = f;
      a.pres = mat_of(pres0->vhat);
      a.dyp = mat_of(dyp_);
      a.tmp = mat_of(temp->vhat);
      a.tbc = mat_of(tbc_ortho_);
      a.bcdiff = mat_of(bcdiff_);
      a.txsd = bxt.d_sd.as<double>(), a.txsl = bxt.d_sl.as<double>();
      a.tysd = byt.d_sd.as<double>(), a.tysl = byt.d_sl.as<double>();
      a.isx = isx;
      a.b2 = b2_of(bxo);
      a.f = fdma_of(solver[f]->adi[0].fdma);
      {  // chunk-major packed copies of the sweep coefficients (fast.h perm_table)
        const FdmaDev& fd = solver[f]->adi[0].fdma;
        const int m = nx - 2;
        perm_.push_back(upload(fk::perm_table(m, true, fk::x_scan_shape(nx), 4,
                                              {host_of(bxo.d_b2lo), host_of(bxo.d_b2di), host_of(bxo.d_b2up), host_of(fd.fp)}, {0, 0, 0, 0})));
        a.pt1 = perm_.back().as<double>();
        perm_.push_back(upload(fk::perm_table(m, false, fk::x_scan2_shape(nx), 4, {host_of(fd.bs), host_of(fd.bp1), host_of(fd.bp2)}, {0, 0, 0})));
        a.pt2 = perm_.back().as<double>();
      }
      a.t = dct_of(bxo);
      a.rhs = fk::Mat{nullptr, 0, 0, 0};
    }
    if (use_xw) {
      // the tile kernel stops after the rhs assembly; B2_x + Fdma_x run as warp-serial column sweeps (fast_xw.cu).
      // Scratch: chat_ holds the [nx, ny] rhs, bconv_ (consumed by the tile kernel) the forward-sweep result
      fk::XwAdiArgs3 w3;
      for (int f = 0; f < 3; ++f) {
        a3.a[f].rhs = mat_of(chat_[f]);
        fk::XwAdiArgs& w = w3.a[f];
        const FdmaDev& fd = solver[f]->adi[0].fdma;
        const int m = nx - 2;
        w.in = mat_of(chat_[f]);
        w.tmp = mat_of(bconv_[f]);
        w.tmp.rows = m;
        w.out = mat_of(w_[f]);
        perm_.push_back(upload(fk::pack_rows(m, 4, {host_of(bxo.d_b2lo), host_of(bxo.d_b2di), host_of(bxo.d_b2up), host_of(fd.fp)}, {0, 0, 0, 0})));
        w.cf = perm_.back().as<double>();
        perm_.push_back(upload(fk::pack_rows(m, 4, {host_of(fd.bs), host_of(fd.bp1), host_of(fd.bp2)}, {0, 0, 0})));
        w.cb = perm_.back().as<double>();
        w.nx = nx;
      }
      add_fast("x_forward_rhs", 13 * fb, [this, a3]() {
        fk::XForwardArgs3 b3 = a3;  // the zero structure of the boundary arrays is looked up at launch (graph capture) time
        for (int f = 0; f < 3; ++f) b3.a[f].tbc_rows = tbc_rows_, b3.a[f].bcdiff_rows = bcdiff_rows_;
        fk::launch_x_forward(b3, 3, stream);
      });
      add_fast("adi_x", 12 * fb, [this, w3]() { fk::launch_xw_adi(w3, 3, stream); });
    } else {
      add_fast("x_forward_rhs_adi_x", 14 * fb, [this, a3]() {
        fk::XForwardArgs3 b3 = a3;
        for (int f = 0; f < 3; ++f) b3.a[f].tbc_rows = tbc_rows_, b3.a[f].bcdiff_rows = bcdiff_rows_;
        fk::launch_x_forward(b3, 3, stream);
      });
    }
  }
  // ---- 5. y half of HholtzAdi (+ pieces of the divergence) -----------------
  {
    fk::YAdiArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::YAdiArgs& a = a3.a[f];
      a.w = mat_of(w_[f]);
      a.out = mat_of(flds[f]->vhat);
      a.aux = f == 0 ? mat_of(vx_) : (f == 1 ? mat_of(ey_) : fk::Mat{nullptr, 0, 0, 0});
      a.mode = f == 0 ? 1 : (f == 1 ? 2 : 0);
      a.sd = byu.d_sd.as<double>(), a.sl = byu.d_sl.as<double>();
      a.isy = isy;
      a.b2 = b2_of(byo);
      a.f = fdma_of(solver[f]->adi[1].fdma);
      {  // chunk-major packed copies of the sweep coefficients (fast.h perm_table)
        const FdmaDev& fd = solver[f]->adi[1].fdma;
        const fk::ScanShape ng = fk::y_scan_shape(ny);
        const int m = ny - 2;
        perm_.push_back(upload(fk::perm_table(m, true, ng, 4, {host_of(byo.d_b2lo), host_of(byo.d_b2di), host_of(byo.d_b2up), host_of(fd.fp)},
                                              {0, 0, 0, 0})));
        a.pt1 = perm_.back().as<double>();
        perm_.push_back(upload(fk::perm_table(m, false, ng, 4, {host_of(fd.bs), host_of(fd.bp1), host_of(fd.bp2)}, {0, 0, 0})));
        a.pt2 = perm_.back().as<double>();
      }
      a.ny = ny;
    }
    add_fast("adi_y", 8 * fb, [this, a3]() { fk::launch_y_adi(a3, 3, stream); });
  }
  // ---- 6. divergence (navier.rs:698-703) + B2x of the Poisson rhs ----------
  {
    fk::XDivArgs a;
    a.vx = mat_of(vx_), a.ey = mat_of(ey_), a.div = mat_of(div_), a.r1 = mat_of(r1_);
    a.sd = bxu.d_sd.as<double>(), a.sl = bxu.d_sl.as<double>();
    a.isx = isx;
    a.b2 = b2_of(bxo);
    a.nx = nx;
    if (use_xw2) {
      perm_.push_back(upload(fk::xw_div_table(nx, isx, host_of(bxu.d_sd), host_of(bxu.d_sl), host_of(bxo.d_b2lo), host_of(bxo.d_b2di),
                                              host_of(bxo.d_b2up))));
      xw_div_.vx = a.vx, xw_div_.ey = a.ey, xw_div_.div = a.div, xw_div_.r1 = a.r1;
      xw_div_.tab = perm_.back().as<double>();
      xw_div_.nx = nx;
      const fk::XwDivArgs w = xw_div_;
      add_fast("divergence_b2x", 4 * fb, [this, w]() { fk::launch_xw_div(w, stream); });
    } else if (use_xs)
      add_fast("divergence_b2x", 4 * fb, [this, a]() { fk::launch_xs_div(a, stream); });
    else
      add_fast("divergence_b2x", 4 * fb, [this, a]() { fk::launch_x_div(a, stream); });
  }
  // ---- 7-9. fast diagonalisation: P., per-mode Fdma_y, Q. (poisson.rs:131-149)
  ops_.push_back(StepOp{1, 0});
  opinfo_.push_back(OpInfo{"poisson_gemm_fwd", 8.0 * ((double)mx * mx + 2.0 * mx * ny), 2.0 * mx * (double)mx * ny});
  {
    fk::YModeArgs a;
    a.g = mat_of(g_), a.h = mat_of(h_);
    a.b2 = b2_of(byo);
    a.m = solver[3]->mode_tabs();
    a.ny = ny;
    add_fast("poisson_mode_y", 3 * fb, [this, a]() { fk::launch_y_mode(a, stream); });
  }
  ops_.push_back(StepOp{2, 0});
  opinfo_.push_back(OpInfo{"poisson_gemm_bwd", 8.0 * ((double)mx * mx + 2.0 * mx * my), 2.0 * mx * (double)mx * my});
  ops_.push_back(StepOp{3, 0});  // pres[1].vhat[[0,0]] = 0   (navier.rs:714)
  opinfo_.push_back(OpInfo{"zero_mode00", 8.0, 0.0});
  // ---- 10-11. projection (navier.rs:683-695) --------------------------------
  if (use_xw2) {
    fk::XwProjectArgs a;
    a.phi = mat_of(pres1->vhat), a.a1 = mat_of(a1_), a.a2 = mat_of(a2_);
    std::vector<double> t1, t2;
    fk::xw_project_tables(nx, isx, host_of(bxn.d_sd), host_of(bxn.d_sl), host_of(bxu.d_sd), host_of(bxu.d_sl), t1, t2);
    perm_.push_back(upload(t1));
    a.t1 = perm_.back().as<double>();
    perm_.push_back(upload(t2));
    a.t2 = perm_.back().as<double>();
    a.nx = nx;
    add_fast("project_x", 3 * fb, [this, a]() { fk::launch_xw_project(a, stream); });
  } else if (use_xs) {
    fk::XsProjectArgs a;
    a.phi = mat_of(pres1->vhat), a.a1 = mat_of(a1_), a.a2 = mat_of(a2_);
    a.p = mat_of(xs_p_), a.d = mat_of(xs_d_);
    a.nsd = bxn.d_sd.as<double>(), a.nsl = bxn.d_sl.as<double>();
    a.t = tdma_of(bxu, nx, fk::xs_scan_shape());
    a.isx = isx;
    a.nx = nx;
    add_fast("project_x", 3 * fb, [this, a]() { fk::launch_xs_project(a, stream); });
  } else {
    fk::XProjectArgs a;
    a.phi = mat_of(pres1->vhat), a.a1 = mat_of(a1_), a.a2 = mat_of(a2_);
    a.nsd = bxn.d_sd.as<double>(), a.nsl = bxn.d_sl.as<double>();
    a.t = tdma_of(bxu, nx, fk::x_scan_shape(nx));
    a.isx = isx;
    a.nx = nx;
    add_fast("project_x", 3 * fb, [this, a]() { fk::launch_x_project(a, stream); });
  }
  {
    fk::YProjectArgs a;
    a.a1 = mat_of(a1_), a.a2 = mat_of(a2_), a.ux = mat_of(ux->vhat), a.uy = mat_of(uy->vhat);
    a.nsd = byn.d_sd.as<double>(), a.nsl = byn.d_sl.as<double>();
    a.t = tdma_of(byu, ny, fk::y_scan_shape(ny));
    a.isy = isy;
    a.ny = ny;
    add_fast("project_y", 6 * fb, [this, a]() { fk::launch_y_project(a, stream); });
  }
  // ---- 12. pressure update (navier.rs:717-721) + d/dy pres for the next step --
  {
    fk::YPresArgs a;
    a.phi = mat_of(pres1->vhat), a.div = mat_of(div_), a.pres = mat_of(pres0->vhat), a.dyp = mat_of(dyp_);
    a.xsd = bxn.d_sd.as<double>(), a.xsl = bxn.d_sl.as<double>();
    a.ysd = byn.d_sd.as<double>(), a.ysl = byn.d_sl.as<double>();
    a.inv_dt = 1.0 / dt, a.nu = nu, a.isy = isy;
    a.ny = ny;
    a.only_dyp = 0;
    add_fast("pressure_update", 5 * fb, [this, a]() { fk::launch_y_pres(a, stream); });
    // - dt/sx d/dx pres of the next step's ux rhs (navier.rs:627), as a column scan
    fk::XsDiffArgs dx;
    dx.src = mat_of(pres0->vhat), dx.dst = mat_of(dxp_);
    dx.sc = -dt * isx;
    dx.nx = nx;
    if (use_xs) add_fast("pressure_dx", 2 * fb, [this, dx]() { fk::launch_xs_dxp(dx, stream); });
    // the same kernels refresh the pressure gradients after the pressure was rewritten from outside (update())
    fk::YPresArgs r = a;
    r.only_dyp = 1;
    fast_dyp_ = [this, r, dx, use_xs]() {
      fk::launch_y_pres(r, stream);
      if (use_xs) fk::launch_xs_dxp(dx, stream);
    };
    // |div u| of the current velocity for exit() (navier.rs:855-879): y parts, then the divergence kernel of the step
    fk::YDivPrepArgs dp;
    dp.ux = mat_of(ux->vhat), dp.uy = mat_of(uy->vhat), dp.vx = mat_of(vx_), dp.ey = mat_of(ey_);
    dp.sd = byu.d_sd.as<double>(), dp.sl = byu.d_sl.as<double>();
    dp.isy = isy, dp.ny = ny;
    fk::XDivArgs xd;
    xd.vx = mat_of(vx_), xd.ey = mat_of(ey_), xd.div = mat_of(div_), xd.r1 = mat_of(r1_);
    xd.sd = bxu.d_sd.as<double>(), xd.sl = bxu.d_sl.as<double>();
    xd.isx = isx;
    xd.b2 = b2_of(bxo);
    xd.nx = nx;
    const fk::XwDivArgs xwd = xw_div_;
    fast_div_ = [this, dp, xd, xwd, use_xs, use_xw2]() {
      fk::launch_y_divprep(dp, stream);
      if (use_xw2)
        fk::launch_xw_div(xwd, stream);
      else if (use_xs)
        fk::launch_xs_div(xd, stream);
      else
        fk::launch_x_div(xd, stream);
    };
  }
  (void)my;
}

// --------------------------------------------------------------------------
// Periodic step on the specialised kernels (fast_p.cu + the physical-space y
// kernels of fast_y.cu): same schedule and arrays as build_step_periodic.
// --------------------------------------------------------------------------
// Stencil tables of the warp-serial per-mode passes (fast_pw.cu): per column j {sd_j, sl_{j-2}, tsd_j, tsl_{j-2}} of the
// field's own y base and of the temperature's; for the divergence {sd_j, sl_{j-2}, 2 j / sy, 0}
void Navier2D::build_pw_tables() {
  if (pw_rs_[0]) return;
  {  // projection + pressure update as row sweeps: Neumann stencil of phi, Dirichlet stencil of the velocity along y
    const Base &bn = *pres1->sp.b1, &bu = *ux->sp.b1;
    std::vector<double> t1, t2;
    fk::pw_project_tables(ny, 1.0 / scale[1], host_of(bn.d_sd), host_of(bn.d_sl), host_of(bu.d_sd), host_of(bu.d_sl), host_of(bu.d_tfs),
                          host_of(bu.d_tfp), host_of(bu.d_tbp), t1, t2);
    perm_.push_back(upload(t1));
    pw_prj_[0] = perm_.back().as<double>();
    perm_.push_back(upload(t2));
    pw_prj_[1] = perm_.back().as<double>();
  }
  const Base &byu = *ux->sp.b1, &byt = *temp->sp.b1;
  const Base* bys[3] = {&byu, &byu, &byt};
  const std::vector<double> tsd = host_of(byt.d_sd), tsl = host_of(byt.d_sl);
  for (int f = 0; f < 3; ++f) {
    perm_.push_back(upload(fk::pack_rows(ny, 4, {host_of(bys[f]->d_sd), host_of(bys[f]->d_sl), tsd, tsl}, {0, -2, 0, -2})));
    pw_rs_[f] = perm_.back().as<double>();
  }
  std::vector<double> two_j(ny);
  for (int j = 0; j < ny; ++j) two_j[j] = 2.0 * (double)j * (1.0 / scale[1]);
  perm_.push_back(upload(fk::pack_rows(ny, 4, {host_of(byu.d_sd), host_of(byu.d_sl), two_j}, {0, -2, 0})));
  pw_rs_[3] = perm_.back().as<double>();
}

void Navier2D::build_step_periodic_fast() {
  const Base &bx = *ux->sp.b0, &byu = *ux->sp.b1, &byt = *temp->sp.b1, &byn = *pres1->sp.b1, &byo = *field->sp.b1;
  const int mk = nx / 2 + 1;
  const double isx = 1.0 / scale[0], isy = 1.0 / scale[1];
  Field2* flds[3] = {ux.get(), uy.get(), temp.get()};
  const Base* bys[3] = {&byu, &byu, &byt};
  const double fb = 8.0 * (double)nx * (double)ny;
  const int val_idx[3] = {0, 1, -1}, dx_idx[3] = {2, 4, 6}, dy_idx[3] = {3, 5, 7};
  const fk::Mat none{nullptr, 0, 0, 0};
  {  // ---- 1. x-backward (c2r): value and (ik/sx) derivative -------------------
    fk::PC2rArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::PC2rArgs& a = a3.a[f];
      a.src = mat_of(flds[f]->vhat), a.val = mat_of(ax_[f]), a.dx = mat_of(adx_[f]);
      a.isx = isx;
      a.tw = bx.fft.twc.as<double2>();
      a.n = nx;
    }
    add_fast("x_backward_c2r", 9 * fb, [this, a3]() { fk::launch_p_c2r(a3, 3, stream); });
  }
  {  // ---- 2. y-backward -> physical space ------------------------------------
    fk::YBackwardArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::YBackwardArgs& a = a3.a[f];
      a.a = mat_of(ax_[f]), a.adx = mat_of(adx_[f]);
      a.val = val_idx[f] >= 0 ? mat_of(phys_[val_idx[f]]) : (has_solid_ ? mat_of(phys_t_) : none);
      a.dy = mat_of(phys_[dy_idx[f]]), a.dx = mat_of(phys_[dx_idx[f]]);
      a.sd = bys[f]->d_sd.as<double>(), a.sl = bys[f]->d_sl.as<double>();
      a.isy = isy;
      a.t = dct_of(byo);
    }
    add_fast("y_backward", 14 * fb, [this, a3]() { fk::launch_y_backward(a3, 3, stream); });
  }
  {  // ---- 3. products + forward DCT-y + dealias-y -----------------------------
    fk::YConvArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::YConvArgs& a = a3.a[f];
      a.u = mat_of(phys_[0]), a.du = mat_of(phys_[dx_idx[f]]), a.v = mat_of(phys_[1]), a.dv = mat_of(phys_[dy_idx[f]]);
      a.bcx = f == 2 ? mat_of(dxtbc_) : none;
      a.bcy = f == 2 ? mat_of(dytbc_) : none;
      set_solid_args(a, f);
      a.out = mat_of(bconv_[f]);
      a.cut = dealias ? (ny * 2) / 3 : ny;
      a.t = dct_of(byo);
    }
    add_fast("conv_y_forward", 17 * fb, [this, a3]() { fk::launch_y_conv(a3, 3, stream); });
  }
  {  // ---- 4. x-forward (r2c) + dealias ------------------------------------------
    fk::PR2cArgs3 a3;
    for (int f = 0; f < 3; ++f) {
      fk::PR2cArgs& a = a3.a[f];
      a.src = mat_of(bconv_[f]), a.dst = mat_of(chat_[f]);
      a.u = a.du = a.v = a.dv = a.bcx = a.bcy = none;
      a.sdst.nparts = 0, a.j0 = 0, a.ny = ny;
      a.cut = dealias ? (mk * 2) / 3 : mk;  // navier.rs:1028 with shape[0] = nx/2+1
      a.tw = bx.fft.twc.as<double2>();
      a.n = nx;
    }
    add_fast("x_forward_r2c", 6 * fb, [this, a3]() { fk::launch_p_r2c(a3, 3, stream); });
  }
  // ---- 5. rhs assembly + per-mode Helmholtz solves (hholtz.rs:156-197).  The buoyancy term of uy
  //         reads the old temperature, so the temperature solve is a launch of its own.
  fk::PHholtzArgs3 h3;
  for (int f = 0; f < 3; ++f) {
    fk::PHholtzArgs& a = h3.a[f];
    a.chat = mat_of(chat_[f]);
    a.fld = mat_of(flds[f]->vhat), a.out = mat_of(flds[f]->vhat);
    a.pres = mat_of(pres0->vhat), a.tmp = mat_of(temp->vhat), a.tbc = mat_of(tbc_ortho_), a.bcdiff = mat_of(bcdiff_);
    a.sd = bys[f]->d_sd.as<double>(), a.sl = bys[f]->d_sl.as<double>();
    a.tsd = byt.d_sd.as<double>(), a.tsl = byt.d_sl.as<double>();
    a.mode = f;
    a.dt = dt, a.isx = isx, a.isy = isy;
    a.b2 = b2_of(byo);
    a.m = solver[f]->mode_tabs();
    a.ny = ny;
    a.k0 = 0;
    a.dyp = mat_of(dyp_);
    build_pw_tables();
    a.rs = pw_rs_[f];
  }
  if (fk::pw_enabled(mk, false)) {
    // row sweeps (fast_pw.cu): the three fields in ONE launch -- the chains of a field are independent of the row
    // count, so a launch of its own for the temperature would only add its latency.  The temperature solve then
    // writes to scratch (w_[2], viewed with the pitch of temp.vhat) and is copied back, because uy reads the old field.
    fk::PHholtzArgs3 a3 = h3;
    a3.a[2].out = fk::Mat{w_[2].d(), temp->vhat.ld, temp->vhat.rows, temp->vhat.cols};
    const size_t tbytes = (size_t)temp->vhat.ld * temp->vhat.rows * 16;
    if (w_[2].buf.bytes < tbytes) throw Error(RP_ERR_INTERNAL, "periodic step: scratch smaller than temp.vhat");
    add_fast("rhs_hholtz_mode_y", 16 * fb, [this, a3, tbytes]() {
      fk::launch_p_hholtz(a3, 3, stream);
      rt::d2d(temp->vhat.buf.p, w_[2].buf.p, tbytes, stream);
    });
  } else {
    add_fast("rhs_hholtz_mode_y", 10 * fb, [this, h3]() { fk::launch_p_hholtz(h3, 2, stream); });
    fk::PHholtzArgs3 t3 = h3;
    t3.a[0] = h3.a[2];
    add_fast("rhs_hholtz_mode_y", 4 * fb, [this, t3]() { fk::launch_p_hholtz(t3, 1, stream); });
  }
  {  // ---- 6. divergence + Poisson (per-mode) -------------------------------------
    fk::PDivPoisArgs a;
    a.ux = mat_of(ux->vhat), a.uy = mat_of(uy->vhat), a.div = mat_of(div_), a.phi = mat_of(pres1->vhat);
    a.sd = byu.d_sd.as<double>(), a.sl = byu.d_sl.as<double>();
    a.isx = isx, a.isy = isy;
    a.b2 = b2_of(byo);
    a.m = solver[3]->mode_tabs();
    a.ny = ny;
    a.k0 = 0;
    build_pw_tables();
    a.rs = pw_rs_[3];
    add_fast("divergence_poisson_mode_y", 5 * fb, [this, a]() { fk::launch_p_divpois(a, stream); });
  }
  {  // ---- 7. projection + pressure update -----------------------------------------
    fk::PProjectArgs a;
    a.phi = mat_of(pres1->vhat), a.ux = mat_of(ux->vhat), a.uy = mat_of(uy->vhat), a.div = mat_of(div_), a.pres = mat_of(pres0->vhat);
    a.nsd = byn.d_sd.as<double>(), a.nsl = byn.d_sl.as<double>();
    a.t = tdma_of(byu, ny, fk::y_scan_shape(ny));
    a.isx = isx, a.isy = isy, a.nu = nu, a.inv_dt = 1.0 / dt;
    a.ny = ny;
    a.k0 = 0;
    build_pw_tables();
    a.w1 = pw_prj_[0], a.w2 = pw_prj_[1];
    // scratch of the row-sweep form: chat_[0], chat_[1] (consumed by the Helmholtz pass), viewed as [mk, my]
    a.z1 = mat_of(chat_[0]), a.z2 = mat_of(chat_[1]);
    a.z1.cols = a.z2.cols = ny - 2;
    add_fast("project_pressure_update", 8 * fb, [this, a]() { fk::launch_p_project(a, stream); });
  }
}

// --------------------------------------------------------------------------
// Slab decomposition of the periodic step over the Fourier modes kx.
// Every per-mode operation (y transforms, stencils, rhs assembly, Helmholtz /
// Poisson solves, projection) is local to a row slab [k0, k0 + mkl) of the
// spectral arrays; the x FFT and the physical-space products are local to a
// column slab [j0, j0 + nyl) of physical y.  One step = phase 1 (kx slab) ->
// all-to-all of 6 arrays -> phase 2 (y slab) -> all-to-all of 3 arrays ->
// phase 3 (kx slab).  The y transform is applied before the x transform in
// the backward direction (the reference does x first, space2.rs:346-356; the
// operators commute).  State arrays are the row slabs of this object's own
// arrays; the exchange buffers belong to the caller.
// --------------------------------------------------------------------------
static fk::Mat row_slab(const Arr& a, int k0, int rows) {
  return fk::Mat{a.d() + (size_t)k0 * a.ld * (a.cplx ? 2 : 1), a.ld, rows, a.cols};
}
static fk::Mat dense(double* p, int rows, int cols) { return fk::Mat{p, cols, rows, cols}; }

static fk::Scatter scatter_of(int world, const int* beg, double* const* peers) {
  fk::Scatter sc;
  sc.nparts = 0;
  if (world <= 0 || !beg || !peers) return sc;
  if (world > 8) throw Error(RP_ERR_INVALID, "at most 8 peers");
  sc.nparts = world;
  for (int q = 0; q < world; ++q) sc.ptr[q] = peers[q], sc.beg[q] = beg[q];
  sc.beg[world] = beg[world];
  return sc;
}

void Navier2D::slab_phase1(int k0, int mkl, double* const out[6], int world, const int* joff, double* const* peers) {
  if (!periodic || !fk::px_supported(nx) || !fk::y_supported(ny)) throw Error(RP_ERR_INVALID, "slab phases need the specialised periodic kernels");
  const Base &byu = *ux->sp.b1, &byt = *temp->sp.b1, &byo = *field->sp.b1;
  Field2* flds[3] = {ux.get(), uy.get(), temp.get()};
  const Base* bys[3] = {&byu, &byu, &byt};
  fk::PYBackArgs3 a3;
  for (int f = 0; f < 3; ++f) {
    fk::PYBackArgs& a = a3.a[f];
    a.src = row_slab(flds[f]->vhat, k0, mkl);
    a.val = dense(out ? out[f] : nullptr, mkl, ny);
    a.dy = dense(out ? out[3 + f] : nullptr, mkl, ny);
    if (!out) a.val.p = a.dy.p = ux->vhat.d();  // non-null marker: the outputs go to the peers
    a.sval = scatter_of(world, joff, peers ? peers + (size_t)f * world : nullptr);
    a.sdy = scatter_of(world, joff, peers ? peers + (size_t)(3 + f) * world : nullptr);
    a.k0 = k0;
    a.sd = bys[f]->d_sd.as<double>(), a.sl = bys[f]->d_sl.as<double>();
    a.isy = 1.0 / scale[1];
    a.t = dct_of(byo);
  }
  fk::launch_p_ybackward(a3, 3, stream);
}

void Navier2D::slab_phase2(int j0, int nyl, const double* const in[6], double* work, double* const out[3], int world,
                           const int* koff, double* const* peers) {
  if (!periodic || !fk::px_supported(nx)) throw Error(RP_ERR_INVALID, "slab phases need the specialised periodic kernels");
  if (has_solid_) throw Error(RP_ERR_INVALID, "solid masks are not supported by the slab-decomposed step");
  const Base& bx = *ux->sp.b0;
  const int mk = nx / 2 + 1;
  const int dx_idx[3] = {2, 4, 6}, dy_idx[3] = {3, 5, 7};
  const fk::Mat none{nullptr, 0, 0, 0};
  auto phys = [&](int i) { return dense(work + (size_t)i * nx * nyl, nx, nyl); };
  fk::PC2rArgs3 c3;
  for (int f = 0; f < 3; ++f) {  // value and (ik/sx) derivative of ux, uy, T
    fk::PC2rArgs& a = c3.a[f];
    a.src = dense(const_cast<double*>(in[f]), mk, nyl);
    a.val = f < 2 ? phys(f) : none;
    a.dx = phys(dx_idx[f]);
    a.isx = 1.0 / scale[0];
    a.tw = bx.fft.twc.as<double2>();
    a.n = nx;
  }
  fk::launch_p_c2r(c3, 3, stream);
  for (int f = 0; f < 3; ++f) {  // d/dy of ux, uy, T
    fk::PC2rArgs& a = c3.a[f];
    a.src = dense(const_cast<double*>(in[3 + f]), mk, nyl);
    a.val = phys(dy_idx[f]);
    a.dx = none;
  }
  fk::launch_p_c2r(c3, 3, stream);
  fk::PR2cArgs3 r3;
  for (int f = 0; f < 3; ++f) {  // products + r2c + dealias in kx
    fk::PR2cArgs& a = r3.a[f];
    a.src = none;
    a.u = phys(0), a.du = phys(dx_idx[f]), a.v = phys(1), a.dv = phys(dy_idx[f]);
    a.bcx = f == 2 ? fk::Mat{dxtbc_.d() + j0, dxtbc_.ld, nx, nyl} : none;
    a.bcy = f == 2 ? fk::Mat{dytbc_.d() + j0, dytbc_.ld, nx, nyl} : none;
    a.dst = dense(out ? out[f] : nullptr, mk, nyl);
    a.cut = dealias ? (mk * 2) / 3 : mk;
    a.tw = bx.fft.twc.as<double2>();
    a.n = nx;
    a.sdst = scatter_of(world, koff, peers ? peers + (size_t)f * world : nullptr);
    a.j0 = j0, a.ny = ny;
  }
  fk::launch_p_r2c(r3, 3, stream);
}

void Navier2D::slab_phase3(int k0, int mkl, const double* const in[3]) {
  if (!periodic || !fk::px_supported(nx) || !fk::y_supported(ny)) throw Error(RP_ERR_INVALID, "slab phases need the specialised periodic kernels");
  const Base &byu = *ux->sp.b1, &byt = *temp->sp.b1, &byn = *pres1->sp.b1, &byo = *field->sp.b1;
  const double isx = 1.0 / scale[0], isy = 1.0 / scale[1];
  Field2* flds[3] = {ux.get(), uy.get(), temp.get()};
  const Base* bys[3] = {&byu, &byu, &byt};
  fk::PYFwdArgs3 y3;
  for (int f = 0; f < 3; ++f) {  // forward DCT-y + dealias in y -> chat_f rows
    fk::PYFwdArgs& a = y3.a[f];
    a.src = dense(const_cast<double*>(in[f]), mkl, ny);
    a.dst = row_slab(chat_[f], k0, mkl);
    a.cut = dealias ? (ny * 2) / 3 : ny;
    a.t = dct_of(byo);
  }
  fk::launch_p_yforward(y3, 3, stream);
  auto slab_mode = [&](Solver2& sv) {
    fk::ModeTabs t = sv.mode_tabs();
    t.lam += k0;
    t.inv += (size_t)k0 * sv.ts.mode.inv_ld;
    return t;
  };
  fk::PHholtzArgs3 h3;
  for (int f = 0; f < 3; ++f) {
    fk::PHholtzArgs& a = h3.a[f];
    a.chat = row_slab(chat_[f], k0, mkl);
    a.fld = row_slab(flds[f]->vhat, k0, mkl), a.out = a.fld;
    a.pres = row_slab(pres0->vhat, k0, mkl), a.tmp = row_slab(temp->vhat, k0, mkl);
    a.tbc = row_slab(tbc_ortho_, k0, mkl), a.bcdiff = row_slab(bcdiff_, k0, mkl);
    a.sd = bys[f]->d_sd.as<double>(), a.sl = bys[f]->d_sl.as<double>();
    a.tsd = byt.d_sd.as<double>(), a.tsl = byt.d_sl.as<double>();
    a.mode = f;
    a.dt = dt, a.isx = isx, a.isy = isy;
    a.b2 = b2_of(byo);
    a.m = slab_mode(*solver[f]);
    a.ny = ny;
    a.k0 = k0;
    a.dyp = row_slab(dyp_, k0, mkl);
    build_pw_tables();
    a.rs = pw_rs_[f];
  }
  if (fk::pw_enabled(mkl, false)) {  // row sweeps: three fields in one launch, temperature through scratch (see build_step_periodic_fast)
    const long long tld = temp->vhat.ld;
    h3.a[2].out = fk::Mat{w_[2].d() + (size_t)k0 * tld * 2, tld, mkl, temp->vhat.cols};
    fk::launch_p_hholtz(h3, 3, stream);
    rt::d2d((char*)temp->vhat.buf.p + (size_t)k0 * tld * 16, (char*)w_[2].buf.p + (size_t)k0 * tld * 16, (size_t)mkl * tld * 16, stream);
  } else {
    fk::launch_p_hholtz(h3, 2, stream);
    fk::PHholtzArgs3 t3 = h3;
    t3.a[0] = h3.a[2];
    fk::launch_p_hholtz(t3, 1, stream);
  }
  {
    fk::PDivPoisArgs a;
    a.ux = row_slab(ux->vhat, k0, mkl), a.uy = row_slab(uy->vhat, k0, mkl);
    a.div = row_slab(div_, k0, mkl), a.phi = row_slab(pres1->vhat, k0, mkl);
    a.sd = byu.d_sd.as<double>(), a.sl = byu.d_sl.as<double>();
    a.isx = isx, a.isy = isy;
    a.b2 = b2_of(byo);
    a.m = slab_mode(*solver[3]);
    a.ny = ny;
    a.k0 = k0;
    build_pw_tables();
    a.rs = pw_rs_[3];
    fk::launch_p_divpois(a, stream);
  }
  {
    fk::PProjectArgs a;
    a.phi = row_slab(pres1->vhat, k0, mkl), a.ux = row_slab(ux->vhat, k0, mkl), a.uy = row_slab(uy->vhat, k0, mkl);
    a.div = row_slab(div_, k0, mkl), a.pres = row_slab(pres0->vhat, k0, mkl);
    a.nsd = byn.d_sd.as<double>(), a.nsl = byn.d_sl.as<double>();
    a.t = tdma_of(byu, ny, fk::y_scan_shape(ny));
    a.isx = isx, a.isy = isy, a.nu = nu, a.inv_dt = 1.0 / dt;
    a.ny = ny;
    a.k0 = k0;
    build_pw_tables();
    a.w1 = pw_prj_[0], a.w2 = pw_prj_[1];
    a.z1 = row_slab(chat_[0], k0, mkl), a.z2 = row_slab(chat_[1], k0, mkl);
    a.z1.cols = a.z2.cols = ny - 2;
    fk::launch_p_project(a, stream);
  }
  time += dt;
}

void Navier2D::build_step_periodic() {
  const Base &bx = *ux->sp.b0, &byu = *ux->sp.b1, &byt = *temp->sp.b1, &byn = *pres1->sp.b1;
  const int mk = nx / 2 + 1, my = ny - 2;
  const double isx = 1.0 / scale[0], isy = 1.0 / scale[1];
  Field2* flds[3] = {ux.get(), uy.get(), temp.get()};
  const Base* bys[3] = {&byu, &byu, &byt};
  const Lay nat = lay_natural();
  // ---- 1. x-backward (c2r): value and (ik/sx) derivative -------------------
  for (int f = 0; f < 3; ++f) {
    ProgBuilder pb(AXIS_X, half_up(my));
    pb.irfft_ld(0, flds[f]->vhat, bx);
    pb.st(0, ax_[f], nx, nat);
    pb.irfft_ld(0, flds[f]->vhat, bx, isx, true);
    pb.st(0, adx_[f], nx, nat);
    add_prog(pb, "x_backward_c2r");
  }
  // ---- 2. y phase ------------------------------------------------------------
  build_y_phase();
  // ---- 3. x-forward (r2c) + dealias ------------------------------------------
  const int kx_cut = dealias ? (mk * 2) / 3 : -1;
  for (int f = 0; f < 3; ++f) {
    ProgBuilder pb(AXIS_X, half_up(ny));
    pb.ld(0, bconv_[f], nx, nat);
    pb.rfft_st(0, chat_[f], bx, 1.0, kx_cut);
    add_prog(pb, "x_forward_r2c");
  }
  // ---- 4. rhs assembly + per-mode Helmholtz solves (hholtz.rs:156-197) ------
  auto inv_of = [&](int s) {
    const FdmaModeDev& md = solver[s]->ts.mode;
    return ArrRef(md.inv.p, md.inv_ld, md.nlanes, md.n, false);
  };
  for (int f = 0; f < 3; ++f) {
    ProgBuilder pb(AXIS_Y, mk);
    Lay l = lay_split(ny);
    pb.ld(0, chat_[f], ny, l, -dt);
    if (f == 0) {
      pb.ld(0, pres0->vhat, ny, l, -dt * isx, LF_ACC | LF_MULIK);
    } else if (f == 1) {
      pb.ld(1, pres0->vhat, ny, l);
      Lay lp = pb.diff(1, ny, l, 1, -dt * isy);
      pb.axpy(0, 1, ny, 1.0, l, lp);
      pb.ld(1, temp->vhat, my, l, 1.0, 0, 0, nullptr, ny);
      pb.toortho(1, byt, l);
      pb.axpy(0, 1, ny, dt, l, l);
      pb.ld(0, tbc_ortho_, ny, l, dt, LF_ACC);
    } else {
      pb.ld(0, bcdiff_, ny, l, 1.0, LF_ACC);
    }
    pb.ld(1, flds[f]->vhat, my, l, 1.0, 0, 0, nullptr, ny);
    pb.toortho(1, *bys[f], l);
    pb.axpy(0, 1, ny, 1.0, l, l);
    pb.ld(1, inv_of(f), my, l, 1.0, LF_BCAST);
    solver[f]->emit_y(pb, 0, 1, l, true);
    pb.st(0, flds[f]->vhat, my, l);
    add_prog(pb, "rhs_hholtz_mode_y");
  }
  // ---- 5. divergence + Poisson (per-mode) -------------------------------------
  {
    ProgBuilder pb(AXIS_Y, mk);
    Lay l = lay_split(ny);
    pb.ld(0, ux->vhat, my, l, isx, LF_MULIK, 0, nullptr, ny);
    pb.toortho(0, byu, l);
    pb.ld(1, uy->vhat, my, l, 1.0, 0, 0, nullptr, ny);
    pb.toortho(1, byu, l);
    Lay ld = pb.diff(1, ny, l, 1, isy);
    pb.axpy(0, 1, ny, 1.0, l, ld);
    pb.st(0, div_, ny, l);
    pb.ld(1, inv_of(3), my, l, 1.0, LF_BCAST);
    solver[3]->emit_y(pb, 0, 1, l, true);
    pb.setzero00(0, l, true);
    pb.st(0, pres1->vhat, my, l);
    add_prog(pb, "divergence_poisson_mode_y");
  }
  // ---- 6. projection + pressure update ------------------------------------------
  {
    ProgBuilder pb(AXIS_Y, mk);
    Lay l = lay_split(ny);
    pb.ld(0, pres1->vhat, my, l, isx, LF_MULIK, 0, nullptr, ny);
    pb.toortho(0, byn, l);
    pb.fromortho(0, byu, l);
    pb.st(0, ux->vhat, my, l, -1.0, LF_ACC);
    pb.ld(0, pres1->vhat, my, l, 1.0, 0, 0, nullptr, ny);
    pb.toortho(0, byn, l);
    pb.copy(1, 0, ny, l, l);
    Lay ld = pb.diff(0, ny, l, 1, isy);
    pb.fromortho(0, byu, ld);
    pb.st(0, uy->vhat, my, ld, -1.0, LF_ACC);
    pb.scale(1, ny, 1.0 / dt, l);
    pb.ld(1, div_, ny, l, -nu, LF_ACC);
    pb.ld(1, pres0->vhat, ny, l, 1.0, LF_ACC);
    pb.st(1, pres0->vhat, ny, l);
    add_prog(pb, "project_pressure_update");
  }
}

void Navier2D::run_op(const StepOp& op) {
  switch (op.kind) {
    case 0: step_[op.idx].launch(stream); break;
    case 1: solver[3]->gemm_fwd(r1_, g_, ny); break;
    case 2: solver[3]->gemm_bwd(h_, pres1->vhat, ny - 2); break;
    case 3: launch_zero_elems(pres1->vhat.d(), 1, stream); break;
    case 4: fast_ops_[op.idx](); break;
  }
}

void Navier2D::run_step() {
  for (const StepOp& op : ops_) run_op(op);
}

// Per-launch device times of one update(), averaged over `reps` eager steps
// (CUDA events on the launching stream).  Advances the solution by `reps` steps.
void Navier2D::profile(int reps, std::vector<double>& ms) {
  prepare_step();
  ms.assign(ops_.size(), 0.0);
#ifndef RP_EMU
  std::vector<cudaEvent_t> ev(ops_.size() + 1);
  for (auto& e : ev) RP_CUDA_CHECK(cudaEventCreate(&e));
  for (int r = 0; r < reps; ++r) {
    RP_CUDA_CHECK(cudaEventRecord(ev[0], stream));
    for (size_t i = 0; i < ops_.size(); ++i) {
      run_op(ops_[i]);
      RP_CUDA_CHECK(cudaEventRecord(ev[i + 1], stream));
    }
    RP_CUDA_CHECK(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < ops_.size(); ++i) {
      float t = 0.f;
      RP_CUDA_CHECK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
      ms[i] += (double)t / reps;
    }
    time += dt;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (size_t i = 0; i < ops_.size(); ++i)
    if (ops_[i].kind == 0) step_[ops_[i].idx].dump_prof(opinfo_[i].name.c_str());
#else
  for (int r = 0; r < reps; ++r) {
    run_step();
    time += dt;
  }
#endif
}

// Everything a step needs before its first launch: the launch list, the streams of the member objects, and
// d/dy pres of the *current* pressure (every step leaves it behind for the next one, so the refresh only runs at
// the first step and after the pressure was rewritten from outside: upload / commit_staged / forward / from_ortho).
void Navier2D::prepare_step() {
  build_step();
  for (Field2* f : {temp.get(), ux.get(), uy.get(), pres0.get(), pres1.get(), field.get()}) f->stream = stream;
  for (auto& s : solver) s->stream = stream;
  if (!periodic && dyp_version_ != pres0->vhat_version) {
    if (fast_dyp_) {
      fast_dyp_();
    } else {
      pres0->gradient(0, 1, scale);
      copy_arr(dyp_, pres0->ortho, stream);
    }
    dyp_version_ = pres0->vhat_version;
  }
}

#ifndef RP_EMU
// Captures one step into graph_.  On any failure the capture is ended, the capture stream destroyed and the
// object's streams restored before the error is rethrown (graph_dirty_ stays set), so the object stays usable.
void Navier2D::capture_graph() {
  if (graph_) {
    cudaGraphExecDestroy(graph_);
    graph_ = nullptr;
  }
  cudaStream_t cs = nullptr;
  RP_CUDA_CHECK(cudaStreamCreate(&cs));
  const cudaStream_t saved = stream;
  auto restore = [&]() {
    stream = saved;
    for (auto& s : solver) s->stream = saved;
    for (Field2* f : {temp.get(), ux.get(), uy.get(), pres0.get(), pres1.get(), field.get()}) f->stream = saved;
  };
  cudaGraph_t g = nullptr;
  bool capturing = false;
  try {
    stream = cs;
    for (auto& s : solver) s->stream = cs;
    RP_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    capturing = true;
    run_step();
    capturing = false;
    RP_CUDA_CHECK(cudaStreamEndCapture(cs, &g));
    RP_CUDA_CHECK(cudaGraphInstantiate(&graph_, g, 0));
  } catch (...) {
    if (capturing) {
      cudaGraph_t dead = nullptr;
      cudaStreamEndCapture(cs, &dead);  // invalidates the capture; `dead` is null on failure
      if (dead) cudaGraphDestroy(dead);
    }
    if (g) cudaGraphDestroy(g);
    if (graph_) {
      cudaGraphExecDestroy(graph_);
      graph_ = nullptr;
    }
    cudaGetLastError();
    restore();
    cudaStreamDestroy(cs);
    throw;
  }
  cudaGraphDestroy(g);
  restore();
  cudaStreamDestroy(cs);
  graph_dirty_ = false;
}
#endif

void Navier2D::update(int nsteps) {
  prepare_step();
#ifndef RP_EMU
  if (use_graph_) {
    if (!graph_ || graph_dirty_) capture_graph();
    for (int i = 0; i < nsteps; ++i) {
      RP_CUDA_CHECK(cudaGraphLaunch(graph_, stream));
      time += dt;
    }
    return;
  }
#endif
  for (int i = 0; i < nsteps; ++i) {
    run_step();
    time += dt;  // navier.rs:764
  }
}

// --------------------------------------------------------------------------
// Diagnostics (src/navier/functions.rs, navier.rs:855-879)
// --------------------------------------------------------------------------
// queues |div u|^2 of the current velocity into red_[0] on `stream` (no host sync)
void Navier2D::enqueue_div2() {
  build_step();
  if (fast_div_) {  // specialised kernels: vx_, ey_, div_, r1_ are scratch between steps
    fast_div_();
    launch_wsum(div_.d(), nullptr, div_.ld, div_.rows, div_.cols, nullptr, nullptr, 3, red_.as<double>(), stream);
    return;
  }
  for (Field2* f : {ux.get(), uy.get()}) f->stream = stream;
  ux->gradient(1, 0, scale);
  uy->gradient(0, 1, scale);
  const int rc = ux->cplx ? 2 : 1;
  Arr& o = ux->ortho;
  launch_combine(o.d(), o.d(), uy->ortho.d(), nullptr, o.ld * rc, o.rows, o.cols * rc, 1.0, 1.0, stream);
  launch_wsum(o.d(), nullptr, o.ld * rc, o.rows, o.cols * rc, nullptr, nullptr, 3, red_.as<double>(), stream);
}

double Navier2D::div_norm() {
  enqueue_div2();
  double r = 0.0;
  rt::d2h(&r, red_.p, 8, stream);
  rt::sync(stream);
  return std::sqrt(r);
}

// exit() without a host sync per step (navier.rs:855-862): div_async() queues |div u|^2 of the current state and its
// copy into a page-locked slot; div_poll() reports the most recent value that has arrived.  integrate() therefore
// sees a NaN one check late instead of stalling the stream every step.
void Navier2D::div_async() {
#ifndef RP_EMU
  if (!div_host_) {
    RP_CUDA_CHECK(cudaMallocHost((void**)&div_host_, 2 * sizeof(double)));
    div_host_[0] = div_host_[1] = 0.0;
    RP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_div_, cudaEventDisableTiming));
  } else if (div_pending_) {
    RP_CUDA_CHECK(cudaEventSynchronize(ev_div_));  // the slot is still in flight: at most one outstanding request
    div_last_ = std::sqrt(div_host_[0]);
    div_have_ = true;
  }
  enqueue_div2();
  rt::d2h(div_host_, red_.p, 8, stream);
  RP_CUDA_CHECK(cudaEventRecord(ev_div_, stream));
  div_pending_ = true;
#else
  div_last_ = div_norm();
  div_have_ = true;
#endif
}
// returns true when a value is available (out = |div u|_2 of the most recently completed request)
bool Navier2D::div_poll(double* out, bool wait) {
#ifndef RP_EMU
  if (div_pending_) {
    cudaError_t e = wait ? cudaEventSynchronize(ev_div_) : cudaEventQuery(ev_div_);
    if (e == cudaSuccess) {
      div_last_ = std::sqrt(div_host_[0]);
      div_have_ = true;
      div_pending_ = false;
    } else if (e != cudaErrorNotReady) {
      RP_CUDA_CHECK(e);
    }
  }
#else
  (void)wait;
#endif
  if (out && div_have_) *out = div_last_;
  return div_have_;
}

void Navier2D::eval(double* o_nu, double* o_nuvol, double* o_re, double* o_div, double* o_ekin) {
  for (Field2* f : {temp.get(), ux.get(), uy.get(), pres0.get(), pres1.get(), field.get()}) f->stream = stream;
  Field2& F = *field;
  const int rc = F.cplx ? 2 : 1;
  auto set_that = [&]() {  // field.vhat = temp.to_ortho() + fieldbc.to_ortho()
    temp->to_ortho();
    launch_combine(F.vhat.d(), temp->ortho.d(), tbc_ortho_.d(), nullptr, F.vhat.ld * rc, F.o0, F.o1 * rc, 1.0, 1.0, stream);
  };
  if (o_div) *o_div = div_norm();
  if (o_nu) {  // functions.rs:12-36
    set_that();
    F.gradient(0, 1, nullptr);
    launch_combine(F.vhat.d(), F.ortho.d(), nullptr, nullptr, F.vhat.ld * rc, F.o0, F.o1 * rc,
                   -1.0 * (1.0 / (scale[1] / 2.0)), 0.0, stream);
    F.backward();
    std::vector<double> xa;
    F.average_axis0(xa);
    *o_nu = (xa[xa.size() - 1] + xa[0]) / 2.0;
  }
  if (o_nuvol) {  // functions.rs:42-75
    set_that();
    F.backward();
    uy->backward();
    Arr& tmp = phys_[7];
    launch_combine(tmp.d(), F.v.d(), nullptr, nullptr, tmp.ld, nx, ny, 1.0, 0.0, stream);  // T physical
    F.gradient(0, 1, nullptr);
    launch_combine(F.vhat.d(), F.ortho.d(), nullptr, nullptr, F.vhat.ld * rc, F.o0, F.o1 * rc, 1.0 / (scale[1] * -1.0), 0.0,
                   stream);
    F.backward();
    // field.v = (dtdz + uy*T/kappa) * 2 * scale[1]
    launch_combine(F.v.d(), F.v.d(), tmp.d(), uy->v.d(), F.v.ld, nx, ny, 2.0 * scale[1], 1.0 / ka, stream);
    *o_nuvol = F.average();
  }
  if (o_re || o_ekin) {  // functions.rs:82-101
    ux->backward();
    uy->backward();
    for (int mode = 1; mode <= 2; ++mode) {
      if ((mode == 1 && !o_re) || (mode == 2 && !o_ekin)) continue;
      // averaging weights of `field` (unscaled coords; the ratio dx/L is scale invariant)
      launch_wsum(ux->v.d(), uy->v.d(), ux->v.ld, nx, ny, F.weights_x(), F.weights_y(), mode, red_.as<double>(), stream);
      double r = 0.0;
      rt::d2h(&r, red_.p, 8, stream);
      rt::sync(stream);
      if (mode == 1)
        *o_re = r * (2.0 * scale[1] / nu);
      else
        *o_ekin = r;
    }
  }
}

}  // namespace rp
