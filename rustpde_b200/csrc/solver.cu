// solver.cu -- HholtzAdi, Hholtz and Poisson (src/solver/hholtz_adi.rs,
// hholtz.rs, poisson.rs, fdma_tensor.rs) on the device.
//
//   HholtzAdi : rhs -> B2x -> B2y -> Fdma_x -> Fdma_y           (hholtz_adi.rs:98-130)
//   Hholtz    : rhs -> B2x -> B2y -> P. -> per-mode Fdma_y -> Q. (hholtz.rs:156-197)
//   Poisson   : same with A = +c I2 S, alpha = 0, lam shift      (poisson.rs:50-149)
// Operators acting on different axes commute, so each solve is regrouped
// into one x pass (strided lanes), the dense contraction(s) and one y pass.
#include <cmath>

#include "model.h"

namespace rp {

static int half_up(int n) { return (n + 1) / 2; }

static Diags combine(const Diags& a, double sa, const Diags& b, double sb) {
  Diags r;
  const int m = (int)a.dia.size();
  r.resize(m);
  for (size_t i = 0; i < r.low.size(); ++i) r.low[i] = a.low[i] * sa + b.low[i] * sb;
  for (size_t i = 0; i < r.dia.size(); ++i) r.dia[i] = a.dia[i] * sa + b.dia[i] * sb;
  for (size_t i = 0; i < r.up1.size(); ++i) r.up1[i] = a.up1[i] * sa + b.up1[i] * sb;
  for (size_t i = 0; i < r.up2.size(); ++i) r.up2[i] = a.up2[i] * sa + b.up2[i] * sb;
  return r;
}

static std::vector<double> dense_from(const Diags& d, double s) {
  const int m = (int)d.dia.size();
  std::vector<double> a((size_t)m * m, 0.0);
  for (int r = 0; r < m; ++r) {
    a[(size_t)r * m + r] = d.dia[r] * s;
    if (r + 2 < m) {
      a[(size_t)r * m + r + 2] = d.up1[r] * s;
      a[(size_t)(r + 2) * m + r] = d.low[r] * s;
    }
    if (r + 4 < m) a[(size_t)r * m + r + 4] = d.up2[r] * s;
  }
  return a;
}

Solver2::Solver2(int kind_, const Space2& sp_, double cx, double cy, double alpha, const EigData* eig)
    : kind(kind_), sp(sp_) {
  const Base &b0 = *sp.b0, &b1 = *sp.b1;
  if (!b1.is_cheb() || b1.is_bc() || b0.is_bc()) throw Error(RP_ERR_INVALID, "solver: unsupported base on an axis");
  x_fourier = !b0.is_cheb();
  n0 = x_fourier ? b0.m : b0.n;
  n1 = b1.n;
  m0 = x_fourier ? b0.m : b0.n - 2;
  m1 = b1.n - 2;
  if (kind == SOLVER_HHOLTZ_ADI) {
    if (x_fourier) throw Error(RP_ERR_INVALID, "HholtzAdi with a Fourier axis is not on the Navier2D path (use Hholtz)");
    // mat = mat_a - c * mat_b = C - c A, pre-swept (hholtz_adi.rs:54-55, fdma.rs:33-37)
    Diags dx = combine(b0.C, 1.0, b0.A, -cx), dy = combine(b1.C, 1.0, b1.A, -cy);
    fdma_sweep(dx);
    fdma_sweep(dy);
    build_fdma_dev(dx, adi[0].fdma);
    build_fdma_dev(dy, adi[1].fdma);
    return;
  }
  // FdmaTensor: a = sign * c * mat_b (laplacian), c = mat_a (mass)  (hholtz.rs:52-70, poisson.rs:60-78)
  const double sign = (kind == SOLVER_POISSON) ? 1.0 : -1.0;
  const double al = (kind == SOLVER_POISSON) ? 0.0 : alpha;
  ts.x_diag = x_fourier;
  if (x_fourier) {
    ts.lam.resize(m0);
    for (int k = 0; k < m0; ++k) ts.lam[k] = sign * (-(double)k * (double)k) * cx;  // lap = diag(-k^2), r2c.rs:372-379
  } else {
    std::vector<double> Q, P;
    if (eig && eig->lam && eig->q && eig->p) {
      ts.lam.assign(eig->lam, eig->lam + m0);
      Q.assign(eig->q, eig->q + (size_t)m0 * m0);
      P.assign(eig->p, eig->p + (size_t)m0 * m0);
    } else {
      std::vector<double> Cx = dense_from(b0.C, 1.0), Ax = dense_from(b0.A, sign * cx);
      const char* gm = getenv("RUSTPDE_B200_EIG");
      if (gm && std::string(gm) == "full")
        lapack_eig_setup(m0, Cx, Ax, ts.lam, Q, P);  // dgeev on the full matrix, like utils.rs:66-98
      else
        lapack_eig_setup_parity(m0, Cx, Ax, ts.lam, Q, P);
    }
    hq_ = Q;
    hp_ = P;
    ts.P.alloc(m0, m0, false);
    ts.Q.alloc(m0, m0, false);
    ts.P.upload(P.data(), 0);
    ts.Q.upload(Q.data(), 0);
    // parity-split mode: every eigenvector is purely even or purely odd (exact zeros elsewhere)
    std::vector<int> par(m0, -1);
    bool cb = true;
    for (int k = 0; k < m0 && cb; ++k) {
      bool has[2] = {false, false};
      for (int i = 0; i < m0; ++i) {
        if (Q[(size_t)i * m0 + k] != 0.0) has[i & 1] = true;
        if (P[(size_t)k * m0 + i] != 0.0) has[i & 1] = true;
      }
      if (has[0] && has[1]) cb = false;
      par[k] = has[1] ? 1 : 0;
    }
    std::vector<int> ke, ko;
    for (int k = 0; k < m0; ++k) (par[k] ? ko : ke).push_back(k);
    const int me = (m0 + 1) / 2, mo = m0 / 2;
    if (cb && (int)ke.size() == me && (int)ko.size() == mo && m0 >= 4) {
      ts.split = true;
      ts.me = me, ts.mo = mo;
      std::vector<double> pe((size_t)me * me), po((size_t)mo * mo), qe((size_t)me * me), qo((size_t)mo * mo);
      for (int a = 0; a < me; ++a)
        for (int b = 0; b < me; ++b) {
          pe[(size_t)a * me + b] = P[(size_t)ke[a] * m0 + 2 * b];
          qe[(size_t)b * me + a] = Q[(size_t)(2 * b) * m0 + ke[a]];
        }
      for (int a = 0; a < mo; ++a)
        for (int b = 0; b < mo; ++b) {
          po[(size_t)a * mo + b] = P[(size_t)ko[a] * m0 + 2 * b + 1];
          qo[(size_t)b * mo + a] = Q[(size_t)(2 * b + 1) * m0 + ko[a]];
        }
      ts.Pe.alloc(me, me, false), ts.Po.alloc(mo, mo, false), ts.Qe.alloc(me, me, false), ts.Qo.alloc(mo, mo, false);
      ts.Pe.upload(pe.data(), 0), ts.Po.upload(po.data(), 0), ts.Qe.upload(qe.data(), 0), ts.Qo.upload(qo.data(), 0);
      // internal mode order: even-parity modes first
      std::vector<double> l2;
      for (int k : ke) l2.push_back(ts.lam[k]);
      for (int k : ko) l2.push_back(ts.lam[k]);
      lam_export_ = ts.lam;
      ts.lam = l2;
    }
    rt::sync(0);
  }
  {
    // lam_export_: reference order (descending), before the singularity shift -- what export_eig() returns and
    // what *_create_with_eig expects, so the round trip never shifts twice
    if (lam_export_.empty()) lam_export_ = ts.lam;
    if (kind == SOLVER_POISSON && std::fabs(lam_export_[0]) < 1e-10)  // poisson.rs:80-83
      for (auto& l : ts.lam) l -= 1e-10;
  }
  Diags Ay = combine(b1.A, sign * cy, b1.A, 0.0);
  build_fdma_mode_dev(Ay, b1.C, ts.lam, al, ts.mode);
}

void Solver2::export_eig(double* lam, double* q, double* p) const {
  if (kind == SOLVER_HHOLTZ_ADI) throw Error(RP_ERR_INVALID, "HholtzAdi has no eigen set-up data");
  if (lam) std::copy(lam_export_.begin(), lam_export_.end(), lam);
  if (!ts.x_diag) {
    if (q) std::copy(hq_.begin(), hq_.end(), q);
    if (p) std::copy(hp_.begin(), hp_.end(), p);
  }
}

void Solver2::emit_x(ProgBuilder& pb, int r, Lay lay) const {
  if (x_fourier) return;
  pb.bandmv(r, *sp.b0, lay);
  if (kind == SOLVER_HHOLTZ_ADI) pb.fdma(r, adi[0].fdma, lay);
}

void Solver2::emit_y(ProgBuilder& pb, int r, int rinv, Lay lay, bool complex_lanes) const {
  pb.bandmv(r, *sp.b1, lay);
  if (kind == SOLVER_HHOLTZ_ADI)
    pb.fdma(r, adi[1].fdma, lay);
  else
    pb.fdmamode(r, rinv, ts.mode, lay, complex_lanes);
}

// C = P . B over `ncols_real` real columns (complex data = interleaved real columns)
static GemmArgs gemm_args(const Arr& A, int M, int K, const Arr& in, int b_r0, int b_rs, Arr& out, int c_r0, int c_rs,
                          int ncols_real) {
  GemmArgs g;
  g.A = A.d();
  g.B = in.d();
  g.C = out.d();
  g.M = M, g.N = ncols_real, g.K = K;
  g.lda = A.ld;
  g.ldb = in.cplx ? in.ld * 2 : in.ld;
  g.ldc = out.cplx ? out.ld * 2 : out.ld;
  g.b_r0 = b_r0, g.b_rs = b_rs, g.c_r0 = c_r0, g.c_rs = c_rs;
  return g;
}
void Solver2::gemm_fwd(const Arr& in, Arr& out, int ncols_real) const {
  if (ts.split) {  // even rows of `in` -> modes [0, me), odd rows -> modes [me, m0)
    launch_dgemm2(gemm_args(ts.Pe, ts.me, ts.me, in, 0, 2, out, 0, 1, ncols_real),
                  gemm_args(ts.Po, ts.mo, ts.mo, in, 1, 2, out, ts.me, 1, ncols_real), stream);
    return;
  }
  launch_dgemm(gemm_args(ts.P, m0, m0, in, 0, 1, out, 0, 1, ncols_real), stream);
}
void Solver2::gemm_bwd(const Arr& in, Arr& out, int ncols_real) const {
  if (ts.split) {
    launch_dgemm2(gemm_args(ts.Qe, ts.me, ts.me, in, 0, 1, out, 0, 2, ncols_real),
                  gemm_args(ts.Qo, ts.mo, ts.mo, in, ts.me, 1, out, 1, 2, ncols_real), stream);
    return;
  }
  launch_dgemm(gemm_args(ts.Q, m0, m0, in, 0, 1, out, 0, 1, ncols_real), stream);
}

void Solver2::build_programs(bool cd) {
  const int q = cd ? 1 : 0;
  if (px_[q].valid || py_[q].valid) return;
  const bool lanes_c = cd || x_fourier;  // complex lanes in the y pass
  Arr& in = lanes_c ? in_c : in_r;
  Arr& out = lanes_c ? out_c : out_r;
  if (!in.buf.p) {
    in.alloc(n0, n1, lanes_c);
    out.alloc(m0, m1, lanes_c);
  }
  Arr &t1 = t1_[q], &t2 = t2_[q], &t3 = t3_[q];
  t1.alloc(m0, n1, lanes_c);
  t2.alloc(m0, n1, lanes_c);
  t3.alloc(m0, m1, lanes_c);
  const bool use_gemm = (kind != SOLVER_HHOLTZ_ADI) && !x_fourier;
  if (!x_fourier) {  // x pass over the n1 columns
    ProgBuilder pb(AXIS_X, lanes_c ? n1 : half_up(n1));
    Lay l = lay_split(n0);
    pb.ld(0, in, n0, l);
    emit_x(pb, 0, l);
    pb.st(0, t1, m0, l);
    px_[q] = pb.build();
  }
  {
    const Arr& src = x_fourier ? in : (use_gemm ? t2 : t1);
    Arr& dst = use_gemm ? t3 : out;
    ProgBuilder pb(AXIS_Y, lanes_c ? m0 : half_up(m0));
    Lay l = lay_split(n1);
    pb.ld(0, src, n1, l);
    if (kind != SOLVER_HHOLTZ_ADI) {
      // reciprocal pivots of this lane's swept system (set-up data)
      pb.ld(1, ArrRef(ts.mode.inv.p, ts.mode.inv_ld, ts.mode.nlanes, ts.mode.n, false), m1, l, 1.0,
            lanes_c ? LF_BCAST : 0);
    }
    emit_y(pb, 0, 1, l, lanes_c);
    pb.st(0, dst, m1, l);
    py_[q] = pb.build();
  }
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

fk::ModeTabs Solver2::mode_tabs() {
  if (have_mode_tabs_) return mode_tabs_;
  const FdmaModeDev& m = ts.mode;
  fk::ModeTabs t;
  t.a_low = m.a_low.as<double>(), t.a_up1 = m.a_up1.as<double>(), t.a_up2 = m.a_up2.as<double>();
  t.c_low = m.c_low.as<double>(), t.c_up1 = m.c_up1.as<double>(), t.c_up2 = m.c_up2.as<double>();
  t.lam = m.lam.as<double>();
  t.alpha = m.alpha;
  t.inv = m.inv.as<double>();
  t.inv_ld = m.inv_ld;
  t.pf = t.pb = t.rf = t.rb = nullptr;
  const fk::ScanShape ng = fk::y_scan_shape(n1);
  if (ng.ng > 0) {  // chunk-major packed copies of the raw bands (fast.h perm_table)
    const Base& by = *sp.b1;  // the B2 rows depend on n only
    const int mm = n1 - 2;
    const std::vector<double> al = host_of(m.a_low), cl = host_of(m.c_low), au1 = host_of(m.a_up1), cu1 = host_of(m.c_up1),
                              au2 = host_of(m.a_up2), cu2 = host_of(m.c_up2);
    perm_.push_back(upload(fk::perm_table(mm, true, ng, 6, {host_of(by.d_b2lo), host_of(by.d_b2di), host_of(by.d_b2up), al, cl},
                                          {0, 0, 0, -2, -2})));
    t.pf = perm_.back().as<double>();
    perm_.push_back(upload(fk::perm_table(mm, false, ng, 8, {au1, cu1, au2, cu2, au2, cu2, al, cl}, {0, 0, 0, 0, -2, -2, -2, -2})));
    t.pb = perm_.back().as<double>();
    // plain per-column copies for the warp-serial sweeps (fast_pw.cu)
    perm_.push_back(upload(fk::pack_rows(mm, 6, {host_of(by.d_b2lo), host_of(by.d_b2di), host_of(by.d_b2up), al, cl}, {0, 0, 0, -2, -2})));
    t.rf = perm_.back().as<double>();
    perm_.push_back(upload(fk::pack_rows(mm, 8, {au1, cu1, au2, cu2, au2, cu2, al, cl}, {0, 0, 0, 0, -2, -2, -2, -2})));
    t.rb = perm_.back().as<double>();
  }
  mode_tabs_ = t;
  have_mode_tabs_ = true;
  return t;
}

bool Solver2::fast_path() const {
  static const char* nf = getenv("RUSTPDE_B200_NO_FAST");
  if (nf && nf[0] == '1') return false;
  if (x_fourier || !fk::y_supported(n1)) return false;
  if (kind == SOLVER_HHOLTZ_ADI) return fk::x_supported(n0);
  return true;  // Chebyshev x: the x pass is the elementwise B2 matvec, any n0
}

int Solver2::launches_per_solve(bool cd) const {
  if (!cd && fast_path()) return kind == SOLVER_HHOLTZ_ADI ? 2 : 4;
  const bool use_gemm = (kind != SOLVER_HHOLTZ_ADI) && !x_fourier;
  return (x_fourier ? 0 : 1) + 1 + (use_gemm ? 2 : 0);
}

static fk::Mat mat_of(const Arr& a) { return fk::Mat{a.d(), a.ld, a.rows, a.cols}; }

// Real data on the specialised kernels (same arithmetic as the fused Navier2D passes): the stand-alone
// HholtzAdi / Hholtz / Poisson::solve of config 2 (examples/hholtz_2d.rs) no longer goes through the lane programs.
void Solver2::solve_fast() {
  fk::apply_kflags();
  if (!in_r.buf.p) {
    in_r.alloc(n0, n1, false);
    out_r.alloc(m0, m1, false);
  }
  Arr &t1 = t1_[0], &t2 = t2_[0], &t3 = t3_[0];
  if (!t1.buf.p) {
    t1.alloc(m0, n1, false);
    t2.alloc(m0, n1, false);
    t3.alloc(m0, m1, false);
  }
  const Base &b0 = *sp.b0, &b1 = *sp.b1;
  if (kind == SOLVER_HHOLTZ_ADI) {
    if (!have_adi_tabs_) {
      const fk::ScanShape sx1 = fk::x_scan_shape(n0), sx2 = fk::x_scan2_shape(n0), sy = fk::y_scan_shape(n1);
      const Base* bs[2] = {&b0, &b1};
      for (int ax = 0; ax < 2; ++ax) {
        const FdmaDev& fd = adi[ax].fdma;
        const Base& b = *bs[ax];
        const int m = (ax == 0 ? n0 : n1) - 2;
        perm_.push_back(upload(fk::perm_table(m, true, ax == 0 ? sx1 : sy, 4,
                                              {host_of(b.d_b2lo), host_of(b.d_b2di), host_of(b.d_b2up), host_of(fd.fp)}, {0, 0, 0, 0})));
        adi_pt_[ax][0] = perm_.back().as<double>();
        perm_.push_back(upload(fk::perm_table(m, false, ax == 0 ? sx2 : sy, 4, {host_of(fd.bs), host_of(fd.bp1), host_of(fd.bp2)}, {0, 0, 0})));
        adi_pt_[ax][1] = perm_.back().as<double>();
      }
      have_adi_tabs_ = true;
    }
    fk::XAdiArgs xa;
    xa.in = mat_of(in_r), xa.out = mat_of(t1);
    xa.pt1 = adi_pt_[0][0], xa.pt2 = adi_pt_[0][1];
    xa.nx = n0;
    fk::launch_x_adi(xa, stream);
    fk::YAdiArgs3 y3;
    fk::YAdiArgs& ya = y3.a[0];
    ya.w = mat_of(t1), ya.out = mat_of(out_r), ya.aux = fk::Mat{nullptr, 0, 0, 0};
    ya.mode = 0;
    ya.sd = ya.sl = nullptr;
    ya.isy = 1.0;
    ya.b2 = fk::B2Tabs{b1.d_b2lo.as<double>(), b1.d_b2di.as<double>(), b1.d_b2up.as<double>()};
    ya.f = fk::FdmaTabs{adi[1].fdma.fp.as<double>(), adi[1].fdma.bs.as<double>(), adi[1].fdma.bp1.as<double>(), adi[1].fdma.bp2.as<double>()};
    ya.pt1 = adi_pt_[1][0], ya.pt2 = adi_pt_[1][1];
    ya.ny = n1;
    y3.a[1] = y3.a[2] = ya;
    fk::launch_y_adi(y3, 1, stream);
    return;
  }
  launch_b2x(in_r.d(), in_r.ld, t1.d(), t1.ld, n0, n1, b0.d_b2lo.as<double>(), b0.d_b2di.as<double>(), b0.d_b2up.as<double>(), stream);
  gemm_fwd(t1, t2, n1);
  fk::YModeArgs ma;
  ma.g = mat_of(t2), ma.h = mat_of(t3);
  ma.b2 = fk::B2Tabs{b1.d_b2lo.as<double>(), b1.d_b2di.as<double>(), b1.d_b2up.as<double>()};
  ma.m = mode_tabs();
  ma.ny = n1;
  fk::launch_y_mode(ma, stream);
  gemm_bwd(t3, out_r, m1);
}

void Solver2::solve_dev(const Arr& in, Arr& out, bool cd) {
  const bool lanes_c = cd || x_fourier;
  Arr& ain = lanes_c ? in_c : in_r;
  Arr& aout = lanes_c ? out_c : out_r;
  if (!ain.buf.p) {
    ain.alloc(n0, n1, lanes_c);
    aout.alloc(m0, m1, lanes_c);
  }
  if (in.buf.bytes != ain.buf.bytes || out.buf.bytes != aout.buf.bytes) throw Error(RP_ERR_SHAPE, "Dimension mismatch in solver input");
  rt::d2d(ain.buf.p, in.buf.p, ain.buf.bytes, stream);
  solve(cd);
  rt::d2d(out.buf.p, aout.buf.p, aout.buf.bytes, stream);
}

void Solver2::solve(bool cd) {
  if (!cd && fast_path()) {
    solve_fast();
    return;
  }
  build_programs(cd);
  const int q = cd ? 1 : 0;
  const bool lanes_c = cd || x_fourier;
  Arr& out = lanes_c ? out_c : out_r;
  const bool use_gemm = (kind != SOLVER_HHOLTZ_ADI) && !x_fourier;
  if (px_[q].valid) px_[q].launch(stream);
  if (use_gemm) gemm_fwd(t1_[q], t2_[q], lanes_c ? 2 * n1 : n1);
  py_[q].launch(stream);
  if (use_gemm) gemm_bwd(t3_[q], out, lanes_c ? 2 * m1 : m1);
}

}  // namespace rp
