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

void Solver2::solve(bool cd) {
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
