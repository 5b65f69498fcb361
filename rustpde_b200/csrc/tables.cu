// tables.cu -- host-side plan/table construction (see tables.h).
#include "tables.h"

#include <cmath>
#include <complex>
#include <mutex>

namespace rp {

typedef long double ld_t;
typedef std::complex<long double> lc_t;
static const ld_t kPi = 3.14159265358979323846264338327950288L;

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// exp(-2 pi i k / L) with exact octant reduction
static lc_t unit_root(long long k, long long L) {
  k %= L;
  if (k < 0) k += L;
  ld_t a = -2.0L * kPi * (ld_t)k / (ld_t)L;
  return lc_t(cosl(a), sinl(a));
}

static void host_fft(std::vector<lc_t>& a) {  // iterative radix-2, forward
  const int n = (int)a.size();
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (int len = 2; len <= n; len <<= 1) {
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        lc_t w = unit_root(k, len);
        lc_t u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

void build_fft_tables(int L, FftTables& out) {
  FftPlan& P = out.plan;
  P.L = L;
  P.pow2 = is_pow2(L) ? 1 : 0;
  P.Lb = P.pow2 ? L : next_pow2(2 * L - 1);
  const int Lt = P.pow2 ? L : P.Lb;
  std::vector<double2> tw(Lt);
  for (int k = 0; k < Lt; ++k) {
    lc_t w = unit_root(k, Lt);
    tw[k] = make_double2((double)w.real(), (double)w.imag());
  }
  out.tw = upload(tw);
  P.tw = out.tw.as<double2>();
  {  // per-span compact copy for the specialised kernels: twc[S/2 - 1 + q] = exp(-2 pi i q / S), q < S/2, S = 2, 4, .., Lt
     // (a stage of span S reads consecutive entries instead of every (Lt/S)-th one of the full table)
    std::vector<double2> twc(Lt > 1 ? Lt - 1 : 1);
    for (int S = 2; S <= Lt; S <<= 1)
      for (int q = 0; q < S / 2; ++q) {
        lc_t w = unit_root(q, S);
        twc[S / 2 - 1 + q] = make_double2((double)w.real(), (double)w.imag());
      }
    out.twc = upload(twc);
  }
  P.chirp = nullptr;
  P.bhat = nullptr;
  if (!P.pow2) {
    // chirp c_j = exp(-i pi j^2 / L) = exp(-2 pi i (j^2 mod 2L) / (2L))
    std::vector<double2> chirp(L);
    std::vector<lc_t> h(P.Lb, lc_t(0, 0));
    for (long long j = 0; j < L; ++j) {
      long long q = (j * j) % (2LL * L);
      lc_t c = unit_root(q, 2LL * L);
      chirp[j] = make_double2((double)c.real(), (double)c.imag());
      lc_t hc = std::conj(c);  // exp(+i pi j^2 / L)
      h[j] = hc;
      if (j > 0) h[P.Lb - j] = hc;
    }
    host_fft(h);
    std::vector<double2> bhat(P.Lb);
    for (int k = 0; k < P.Lb; ++k) {
      lc_t v = h[k] / (ld_t)P.Lb;
      bhat[k] = make_double2((double)v.real(), (double)v.imag());
    }
    out.chirp = upload(chirp);
    out.bhat = upload(bhat);
    // position of X[k] after the in-place DIF stages (radix 2^(log2 Lb % 3) first, then 8)
    int lg = 0;
    while ((1 << lg) < P.Lb) ++lg;
    std::vector<double2> bdr(P.Lb);
    for (int k = 0; k < P.Lb; ++k) {
      int kk = k, span = P.Lb, pos = 0, ls = lg;
      while (span > 1) {
        const int lr = (ls % 3) ? (ls % 3) : 3, R = 1 << lr;
        pos += (kk % R) * (span / R);
        kk /= R;
        span /= R;
        ls -= lr;
      }
      bdr[pos] = bhat[k];
    }
    // the fused middle stage of the convolution (fast.cuh dif_dit_mid) reads rows 8u .. 8u+7 per thread: stored
    // transposed ([r][u]) the 16-byte loads of a warp are contiguous
    std::vector<double2> bdt(P.Lb);
    for (int u = 0; u < P.Lb / 8; ++u)
      for (int r = 0; r < 8; ++r) bdt[(size_t)r * (P.Lb / 8) + u] = bdr[(size_t)u * 8 + r];
    out.bhat_dr = upload(bdt);
    P.chirp = out.chirp.as<double2>();
    P.bhat = out.bhat.as<double2>();
  }
}

// --------------------------------------------------------------------------
static void b2_rows(int n, std::vector<double>& lo, std::vector<double>& di, std::vector<double>& up) {
  // B2 = _pinv(n, 2), ortho.rs:160-171; indexed by row i
  lo.assign(n, 0.0);
  di.assign(n, 0.0);
  up.assign(n, 0.0);
  if (n > 2) lo[2] = 0.25;
  for (int i = 3; i < n; ++i) lo[i] = 1.0 / (double)(4LL * i * (i - 1));
  for (int i = 2; i < n - 2; ++i) di[i] = -1.0 / (double)(2LL * ((long long)i * i - 1));
  for (int i = 2; i < n - 4; ++i) up[i] = 1.0 / (double)(4LL * i * (i + 1));
}

static void build_cheb_family(Base& b) {
  const int n = b.n, m = n - 2, N = n - 1;
  // nodes, ortho.rs:71-80
  b.x.resize(n);
  for (int k = 0; k < n; ++k) b.x[k] = -std::sin(M_PI * ((double)N - 2.0 * k) / (2.0 * (double)N));
  // DCT plan
  build_fft_tables(N, b.fft);
  std::vector<double2> sc(N / 2 + 1);
  for (int j = 0; j <= N / 2; ++j) {
    ld_t a = kPi * (ld_t)j / (ld_t)N;
    sc[j] = make_double2((double)sinl(a), (double)cosl(a));
  }
  b.d_sc = upload(sc);
  DctPlan dp;
  dp.n = n;
  dp.fft = b.fft.plan;
  dp.sc = b.d_sc.as<double2>();
  b.d_dct = upload_struct(dp);
  // stencil (composite_stencil.rs:117-157)
  if (b.is_composite()) {
    b.sd.assign(m, 1.0);
    b.sl.assign(m, -1.0);
    if (b.kind == BASE_CHEB_NEUMANN)
      for (int k = 0; k < m; ++k) {
        double k2 = (double)((long long)k * k), kk2 = (double)((long long)(k + 2) * (k + 2));
        b.sl[k] = -1.0 * k2 / kk2;
      }
    b.d_sd = upload(b.sd);
    b.d_sl = upload(b.sl);
    // (S^T S) tridiagonal, composite_stencil.rs:160-171, pre-factored like linalg.rs:14-57
    std::vector<double> mainv(m), off(m > 2 ? m - 2 : 0);
    for (int i = 0; i < m; ++i) mainv[i] = b.sd[i] * b.sd[i] + b.sl[i] * b.sl[i];
    for (int i = 0; i + 2 < m; ++i) off[i] = b.sd[i + 2] * b.sl[i];
    std::vector<double> w(m > 2 ? m - 2 : 0), fs(m), fp(m, 0.0), bp(m, 0.0);
    for (int i = 0; i < m; ++i) {
      double den = mainv[i];
      if (i >= 2) den = mainv[i] - off[i - 2] * w[i - 2];
      if (i < m - 2) w[i] = off[i] / den;
      fs[i] = 1.0 / den;
      if (i >= 2) fp[i] = -off[i - 2] / den;
    }
    for (int i = 0; i + 2 < m; ++i) bp[i] = -w[i];
    b.d_tfs = upload(fs);
    b.d_tfp = upload(fp);
    b.d_tbp = upload(bp);
    TdmaTab tt;
    tt.fs = b.d_tfs.as<double>();
    tt.fp = b.d_tfp.as<double>();
    tt.bp = b.d_tbp.as<double>();
    b.d_tdma = upload_struct(tt);
  }
  if (b.kind == BASE_CHEB_DIRICHLET_BC) {
    b.t0[0] = 0.5, b.t0[1] = 0.5, b.t1[0] = -0.5, b.t1[1] = 0.5;
    return;
  }
  if (b.kind == BASE_CHEB_NEUMANN_BC) {
    b.t0[0] = 0.5, b.t0[1] = 0.5, b.t1[0] = -0.125, b.t1[1] = 0.125;
    return;
  }
  // preconditioner rows and operator bands (SURVEY 8a'')
  std::vector<double> lo, di, up;
  b2_rows(n, lo, di, up);
  b.b2lo.assign(m, 0.0);
  b.b2di.assign(m, 0.0);
  b.b2up.assign(m, 0.0);
  for (int r = 0; r < m; ++r) {
    const int i = r + 2;
    b.b2lo[r] = lo[i];
    if (r < m - 2) b.b2di[r] = di[i];
    if (r < m - 4) b.b2up[r] = up[i];
  }
  b.d_b2lo = upload(b.b2lo);
  b.d_b2di = upload(b.b2di);
  b.d_b2up = upload(b.b2up);
  auto s_entry = [&](int row, int col) -> double {
    if (col < 0 || col >= m) return 0.0;
    if (b.kind == BASE_CHEBYSHEV) return row == col + 2 ? 1.0 : 0.0;
    if (row == col) return b.sd[col];
    if (row == col + 2) return b.sl[col];
    return 0.0;
  };
  auto a_entry = [&](int r, int c) { return s_entry(r + 2, c); };
  auto c_entry = [&](int r, int c) {
    const int i = r + 2;
    double tot = 0.0;
    const int ks[3] = {i - 2, i, i + 2};
    const double bs[3] = {lo[i], di[i], (i + 2 < n) ? up[i] : 0.0};
    for (int q = 0; q < 3; ++q)
      if (ks[q] >= 0 && ks[q] < n && bs[q] != 0.0) tot += bs[q] * s_entry(ks[q], c);
    return tot;
  };
  b.A.resize(m);
  b.C.resize(m);
  for (int r = 0; r < m; ++r) {
    b.A.dia[r] = a_entry(r, r);
    b.C.dia[r] = c_entry(r, r);
    if (r + 2 < m) {
      b.A.up1[r] = a_entry(r, r + 2);
      b.C.up1[r] = c_entry(r, r + 2);
      b.A.low[r] = a_entry(r + 2, r);
      b.C.low[r] = c_entry(r + 2, r);
    }
    if (r + 4 < m) {
      b.A.up2[r] = a_entry(r, r + 4);
      b.C.up2[r] = c_entry(r, r + 4);
    }
  }
}

static void build_fourier(Base& b) {
  const int n = b.n;
  // c2c.rs:63-66: Array1::range(0, 2pi, 2pi/n)
  const double step = 2.0 * M_PI / (double)n;
  const int cnt = (int)std::ceil((2.0 * M_PI - 0.0) / step);
  b.x.resize(cnt);
  for (int i = 0; i < cnt; ++i) b.x[i] = 0.0 + step * (double)i;
  build_fft_tables(n, b.fft);
  b.d_fft = upload_struct(b.fft.plan);
}

std::shared_ptr<Base> get_base(int kind, int n) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, std::shared_ptr<Base>> cache;
  std::lock_guard<std::mutex> g(mu);
  auto key = std::make_pair(kind, n);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  if (kind < 0 || kind > BASE_FOURIER_R2C) throw Error(RP_ERR_INVALID, "unknown base kind");
  if (n < 6) throw Error(RP_ERR_INVALID, "base size must be >= 6");
  auto b = std::make_shared<Base>();
  b->kind = kind;
  b->n = n;
  if (kind == BASE_FOURIER_R2C) {
    b->m = n / 2 + 1;
    build_fourier(*b);
  } else {
    b->m = (kind == BASE_CHEBYSHEV) ? n : (b->is_bc() ? 2 : n - 2);
    build_cheb_family(*b);
  }
  cache[key] = b;
  return b;
}

// --------------------------------------------------------------------------
void fdma_sweep(Diags& d) {  // fdma.rs:73-82
  const int n = (int)d.dia.size();
  for (int i = 2; i < n; ++i) {
    d.low[i - 2] /= d.dia[i - 2];
    d.dia[i] -= d.low[i - 2] * d.up1[i - 2];
    if (i < n - 2) d.up1[i] -= d.low[i - 2] * d.up2[i - 2];
  }
}

void build_fdma_dev(const Diags& s, FdmaDev& out) {
  const int n = (int)s.dia.size();
  std::vector<double> fp(n, 0.0), bs(n), bp1(n, 0.0), bp2(n, 0.0);
  for (int i = 0; i < n; ++i) {
    if (i >= 2) fp[i] = -s.low[i - 2];
    bs[i] = 1.0 / s.dia[i];
    if (i < n - 2) bp1[i] = -s.up1[i] * bs[i];
    if (i < n - 4) bp2[i] = -s.up2[i] * bs[i];
  }
  out.n = n;
  out.fp = upload(fp);
  out.bs = upload(bs);
  out.bp1 = upload(bp1);
  out.bp2 = upload(bp2);
  FdmaTab t;
  t.fp = out.fp.as<double>();
  t.bs = out.bs.as<double>();
  t.bp1 = out.bp1.as<double>();
  t.bp2 = out.bp2.as<double>();
  out.tab = upload_struct(t);
}

void build_fdma_mode_dev(const Diags& A, const Diags& C, const std::vector<double>& lam, double alpha, FdmaModeDev& out) {
  const int n = (int)A.dia.size();
  const int nl = (int)lam.size();
  out.n = n;
  out.nlanes = nl;
  out.alpha = alpha;
  out.inv_ld = (n + 7) & ~7;
  std::vector<double> inv((size_t)nl * out.inv_ld, 0.0);
  std::vector<double> dia(n), up1(n);
  for (int l = 0; l < nl; ++l) {  // per-lane sweep, hholtz.rs:185-190 + fdma.rs:73-82
    const double mu = lam[l] + alpha;
    for (int i = 0; i < n; ++i) dia[i] = A.dia[i] + C.dia[i] * mu;
    for (int i = 0; i + 2 < n; ++i) up1[i] = A.up1[i] + C.up1[i] * mu;
    for (int i = 2; i < n; ++i) {
      const double low = (A.low[i - 2] + C.low[i - 2] * mu) / dia[i - 2];
      dia[i] -= low * up1[i - 2];
      if (i < n - 2) up1[i] -= low * (A.up2[i - 2] + C.up2[i - 2] * mu);
    }
    double* row = &inv[(size_t)l * out.inv_ld];
    for (int i = 0; i < n; ++i) row[i] = 1.0 / dia[i];
  }
  out.inv = upload(inv);
  auto pad = [&](const std::vector<double>& v) {
    std::vector<double> w(n, 0.0);
    for (size_t i = 0; i < v.size(); ++i) w[i] = v[i];
    return w;
  };
  out.a_low = upload(pad(A.low));
  out.a_dia = upload(pad(A.dia));
  out.a_up1 = upload(pad(A.up1));
  out.a_up2 = upload(pad(A.up2));
  out.c_low = upload(pad(C.low));
  out.c_dia = upload(pad(C.dia));
  out.c_up1 = upload(pad(C.up1));
  out.c_up2 = upload(pad(C.up2));
  out.lam = upload(lam);
  FdmaModeTab t;
  t.a_low = out.a_low.as<double>();
  t.a_dia = out.a_dia.as<double>();
  t.a_up1 = out.a_up1.as<double>();
  t.a_up2 = out.a_up2.as<double>();
  t.c_low = out.c_low.as<double>();
  t.c_dia = out.c_dia.as<double>();
  t.c_up1 = out.c_up1.as<double>();
  t.c_up2 = out.c_up2.as<double>();
  t.lam = out.lam.as<double>();
  t.alpha = alpha;
  out.tab = upload_struct(t);
}

}  // namespace rp
